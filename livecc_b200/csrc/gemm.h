// Internal interface of the tcgen05 GEMM core (see gemm_tcgen05.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcc {

// Fused epilogues; values are part of the C ABI (include/livecc_b200.h, LCC_EPI_*).
enum GemmEpilogue {
    EPI_NONE = 0,            // C = bf16(acc)
    EPI_BIAS = 1,            // C = bf16(acc + bias)
    EPI_BIAS_QUICKGELU = 2,  // C = quick_gelu(bf16(acc + bias))          (ViT fc1, mq2vl.py:333-337)
    EPI_BIAS_GELU = 3,       // C = gelu_erf(bf16(acc + bias))            (merger, mq2vl.py:317-326)
    EPI_RESIDUAL = 4,        // C = bf16(bf16(acc) + residual)            (o_proj/down_proj, mq2vl.py:645-660)
    EPI_BIAS_RESIDUAL = 5,   // C = bf16(bf16(acc + bias) + residual)     (ViT proj/fc2, mq2vl.py:479-487)
    EPI_SWIGLU = 6,          // C[:, j] = bf16(silu(bf16 gate_j) * bf16 up_j), gate/up rows interleaved by 16
    EPI_PARTIAL_F32 = 7,     // internal (split-K): fp32 partial tiles, reduced by splitk_reduce_kernel
};

struct GemmArgs {
    const void* A;  // [M, K] bf16, row stride lda
    const void* B;  // [N, K] bf16, row stride ldb   (nn.Linear weight layout)
    void* C;        // [M, N] bf16 (EPI_SWIGLU: [M, N/2]), row stride ldc
    int M, N, K;
    int lda, ldb, ldc;
    const void* bias;      // [N] bf16 or null
    const void* residual;  // [M, N] bf16, row stride ldr, or null
    int ldr;
    int epi;
    int block_n;  // N tile: 0 = cost model, else a multiple of 16 in [32, 256] (EPI_SWIGLU: multiple of 32)
    // optional split-K scratch (fp32, >= splits*M*N*4 bytes); null = never split. `splits` is set internally.
    void* splitk_ws = nullptr;
    size_t splitk_ws_bytes = 0;
    int splits = 1;
};

int gemm_bf16_tn(const GemmArgs& a, int num_sms, cudaStream_t stream);

// TMA descriptor of a row-major bf16 matrix [rows, cols] (row stride ld elements) with box [box_rows, box_cols];
// box_cols*2 bytes must equal the swizzle span: 128 (SWIZZLE_128B) or 32 (swizzle32 = SWIZZLE_32B).
// Out-of-range rows/cols are zero-filled.
// 3-D variant (SWIZZLE_128B, inner box 64 elements): dims/box innermost first, strides in elements for dims 1 and 2.
int make_tmap_bf16_3d_sw128(CUtensorMap* tm, const void* ptr, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                            int64_t stride2, int box1, int box2);
int make_tmap_bf16_2d_box(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                          int box_rows, bool swizzle32);

}  // namespace lcc
