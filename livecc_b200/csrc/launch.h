// Launch helper: ordinary launch or programmatic dependent launch (PDL) of a kernel.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>

namespace lcc {

// Process-wide count of kernel launches issued by this library (every launch site calls count_launch()). Launches
// recorded into a CUDA graph are counted once at capture; the host adds (nodes x replays) itself. Read through
// lcc_launch_count() (bench.py's `gpu_launches` is this counter, not arithmetic).
inline std::atomic<uint64_t> g_launches{0};
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remember per (kernel
// instantiation, device) that it was raised, so a second engine on another device of the same process works.
struct SmemAttrOnce {
    bool set[64] = {};
};
template <typename Kern>
inline int ensure_dyn_smem(SmemAttrOnce& once, Kern kern, int bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!once.set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -1;
        once.set[dev] = true;
    }
    return 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace lcc
