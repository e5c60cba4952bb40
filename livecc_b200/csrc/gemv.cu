// Decode-step (M = 1) projections: weight-streaming GEMV kernels, HBM-bound by construction
// (14.1 GB of bf16 weights per generated token at 7B). CUDA cores only — a 1-row
// operand cannot feed a tensor-core tile — with 16-byte non-allocating loads, >= 8 independent loads
// in flight per lane, the activation vector staged once per CTA in shared memory, and the
// RMSNorm / bias / SwiGLU / residual work fused around the dot products.
//
// Launch-gap hiding: every kernel is launched with programmatic dependent launch (PDL). A CTA first
// issues its first batch of weight loads (weights are never written during a step), then executes
// griddepcontrol.wait, and only then reads the activation vector / finished flag produced by the
// previous kernel. The weight stream therefore starts while the predecessor is still draining.
//
//   gemv_rows_kernel  : K <= 8192.  Each warp owns ROWS consecutive weight rows (full K).
//   gemv_splitk_kernel: large K (down_proj, K = 18944). A CTA owns ROWS rows; its 8 warps split K.
// (A persistent one-CTA-per-SM variant was measured slower — 3.87 vs 3.29 ms/step — because it gives up the
//  thread-level parallelism that hides the load->use latency; see DESIGN.md.)
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"
#include "ops.h"

namespace lcc {

enum { GV_BIAS = 0, GV_RESIDUAL = 1, GV_SWIGLU = 2, GV_LOGITS = 3 };

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x) {
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(ww[j]), b = unpack_bf16x2(xw[j]);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
    }
    return acc;
}

// Stage x[K] (bf16) into shared memory; with NORM, apply Qwen2VLRMSNorm (mq2vl.py:126-131):
// xs = w_norm * bf16(x * rsqrt(mean(x^2) + eps)).
template <bool NORM>
__device__ __forceinline__ void stage_x(bf16* xs, const bf16* __restrict__ x, const bf16* __restrict__ nw,
                                        float eps, int K, float* red) {
    float sq = 0.f;
    for (int c = threadIdx.x * 8; c < K; c += blockDim.x * 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + c);
        *reinterpret_cast<uint4*>(xs + c) = u;
        if (NORM) {
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(uw[j]); sq += f.x * f.x + f.y * f.y; }
        }
    }
    if (NORM) {
        const float tot = block_sum(sq, red);
        const float rs = rsqrtf(tot / (float)K + eps);
        for (int c = threadIdx.x * 8; c < K; c += blockDim.x * 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(xs + c);
            const uint4 wv = *reinterpret_cast<const uint4*>(nw + c);
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
            const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(uw[j]), g = unpack_bf16x2(ww[j]);
                o[j] = pack_bf16x2(g.x * rbf(f.x * rs), g.y * rbf(f.y * rs));
            }
            *reinterpret_cast<uint4*>(xs + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();
}

struct GemvParams {
    const bf16* W; int ldw;
    const bf16* x;          // [K]
    const bf16* norm_w;     // [K] or null
    float eps;
    int N, K;
    const bf16* bias;       // GV_BIAS
    bf16* out;              // GV_BIAS: [N]; GV_SWIGLU: [N/2]; GV_RESIDUAL: in/out residual stream [N]
    float* out_f32;         // GV_LOGITS: raw logits [N]
    float* out_f32_b;       // GV_LOGITS: second copy (processed-logits buffer)
    const int* finished;
    const uint8_t* pf_ptr;  // L2 prefetch of the NEXT kernel's weights, issued by the last CTAs of this grid
    unsigned pf_bytes;
};

// The decode step is a strict chain of weight-streaming kernels. While the tail of one kernel drains (SM
// imbalance, last-wave effects) HBM would idle; the last CTAs of every GEMV therefore prefetch the head of
// the next kernel's weight matrix into the 126 MB L2 (fire-and-forget `prefetch.global.L2`), so the next
// kernel's first megabytes come from L2 and the small matrices (qkv 33 MB, o_proj 26 MB) almost entirely.
__device__ __forceinline__ void l2_prefetch_tail(const GemvParams& p) {
    if (!p.pf_ptr || p.pf_bytes == 0) return;
    const unsigned ntail = min(gridDim.x, 296u);
    if (blockIdx.x + ntail < gridDim.x) return;
    const unsigned part = blockIdx.x - (gridDim.x - ntail);
    const unsigned lines = (p.pf_bytes + 127u) >> 7;
    const unsigned per = (lines + ntail - 1) / ntail;
    const unsigned l0 = part * per, l1 = min(lines, l0 + per);
    for (unsigned l = l0 + threadIdx.x; l < l1; l += blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pf_ptr + (size_t)l * 128));
}

template <int ROWS, bool NORM, int EPI, int UNROLL>
__global__ void __launch_bounds__(256) gemv_rows_kernel(const GemvParams p) {
    extern __shared__ __align__(16) uint8_t smem_gemv[];
    bf16* xs = reinterpret_cast<bf16*>(smem_gemv);
    __shared__ float red[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * ROWS;
    const bool active = row0 < p.N;
    // GV_SWIGLU: the warp's two rows are gate row and its matching up row (16 apart in a 32-row group)
    int rows[ROWS];
    if (EPI == GV_SWIGLU) {
        const int j = min(row0, p.N - 2) / 2;  // output index
        rows[0] = (j >> 4) * 32 + (j & 15);
        if (ROWS > 1) rows[ROWS - 1] = rows[0] + 16;
    } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) rows[r] = min(row0 + r, p.N - 1);
    }
    const int K = p.K;
    // ---- weight prefetch: weights are never written during a step, so the first batch of loads
    //      (ROWS x UNROLL 16-byte loads per lane) is issued before the activation vector of the previous
    //      kernel is even looked at. Loads past the end of the row are predicated off. ----
    uint4 w[ROWS][UNROLL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int cc = lane * 8 + u * 256;
            w[r][u] = (active && cc < K) ? ld_stream16(p.W + (size_t)rows[r] * p.ldw + cc) : make_uint4(0, 0, 0, 0);
        }
    pdl_wait();
    if (p.finished && *p.finished) return;
    // NORM: x is normalised once per CTA into shared memory. Otherwise x is read straight from global memory
    // through the read-only path (L1-resident after the first touch on an SM): no staging, no block barrier.
    const bf16* xsrc = xs;
    if (NORM) stage_x<true>(xs, p.x, p.norm_w, p.eps, K, red);
    else xsrc = p.x;
    if (!active) { pdl_launch_dependents(); l2_prefetch_tail(p); return; }
    auto ldx = [&](int off) -> uint4 {
        return NORM ? *reinterpret_cast<const uint4*>(xsrc + off) : __ldg(reinterpret_cast<const uint4*>(xsrc + off));
    };

    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    // first batch: already in registers (loads past K were predicated to zero; x reads are clamped)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int cc = min(lane * 8 + u * 256, K - 8);
        const uint4 xv = ldx(cc);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] += dot8(w[r][u], xv);
    }
    int c = lane * 8 + UNROLL * 256;
    // full batches: plain loads, no predication
    for (; (c - lane * 8) + UNROLL * 256 <= K; c += UNROLL * 256) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) w[r][u] = ld_stream16(p.W + (size_t)rows[r] * p.ldw + c + u * 256);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint4 xv = ldx(c + u * 256);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += dot8(w[r][u], xv);
        }
    }
    // remainder (< UNROLL chunks per lane): one predicated batch, all loads in flight together
    if ((c - lane * 8) < K) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = c + u * 256;
                w[r][u] = (cc < K) ? ld_stream16(p.W + (size_t)rows[r] * p.ldw + cc) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int cc = min(c + u * 256, K - 8);
            const uint4 xv = ldx(cc);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += dot8(w[r][u], xv);
        }
    }
    pdl_launch_dependents();  // this warp's weight stream is done: let the next kernel start its prefetch
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
    l2_prefetch_tail(p);
    if (lane != 0) return;
    if (EPI == GV_SWIGLU) {
        const float gt = rbf(acc[0]), up = rbf(acc[ROWS - 1]);
        const float sl = rbf(gt / (1.0f + expf(-gt)));
        p.out[row0 / 2] = f2bf(sl * up);
    } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int n = row0 + r;
            if (n >= p.N) break;
            if (EPI == GV_BIAS) p.out[n] = f2bf(acc[r] + bf2f(p.bias[n]));
            else if (EPI == GV_RESIDUAL) p.out[n] = f2bf(rbf(acc[r]) + bf2f(p.out[n]));
            else if (EPI == GV_LOGITS) {
                const float v = rbf(acc[r]);  // lm_head output is bf16, then .float() (gen/utils.py:2762)
                p.out_f32[n] = v;
                if (p.out_f32_b) p.out_f32_b[n] = v;
            }
        }
    }
}

// Large-K residual GEMV (down_proj). A CTA owns ROWS rows; its 8 warps split K. The activation vector is NOT
// staged in shared memory: a 16-byte x chunk is loaded once per lane (read-only path, L1-resident after the
// first touch on an SM) and reused for all ROWS rows, so a CTA starts streaming weights immediately and the
// 717 CTAs do not each copy 37 KB of x through L2.
template <int ROWS, int UNROLL>
__global__ void __launch_bounds__(256) gemv_splitk_kernel(const GemvParams p) {
    static_assert(ROWS <= 8, "part[] is sized for <= 8 rows");
    __shared__ float part[8][ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * ROWS;
    const int K = p.K;
    int c = (warp * 32 + lane) * 8;
    uint4 w[ROWS][UNROLL];
    const bool full_first = (c + (UNROLL - 1) * 2048) < K;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            w[r][u] = full_first ? ld_stream16(p.W + (size_t)min(row0 + r, p.N - 1) * p.ldw + c + u * 2048)
                                 : make_uint4(0, 0, 0, 0);
    pdl_wait();
    if (p.finished && *p.finished) return;
    const bf16* __restrict__ xg = p.x;
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    if (full_first) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xg + c + u * 2048));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += dot8(w[r][u], xv);
        }
        c += UNROLL * 2048;
    }
    for (; c + (UNROLL - 1) * 2048 < K; c += UNROLL * 2048) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                w[r][u] = ld_stream16(p.W + (size_t)min(row0 + r, p.N - 1) * p.ldw + c + u * 2048);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xg + c + u * 2048));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += dot8(w[r][u], xv);
        }
    }
    for (; c < K; c += 2048) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xg + c));
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            acc[r] += dot8(ld_stream16(p.W + (size_t)min(row0 + r, p.N - 1) * p.ldw + c), xv);
    }
    pdl_launch_dependents();
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        acc[r] = warp_sum(acc[r]);
        if (lane == 0) part[warp][r] = acc[r];
    }
    l2_prefetch_tail(p);
    __syncthreads();
    if (threadIdx.x < ROWS) {
        const int n = row0 + threadIdx.x;
        if (n < p.N) {
            float t = 0.f;
#pragma unroll
            for (int wi = 0; wi < 8; ++wi) t += part[wi][threadIdx.x];
            p.out[n] = f2bf(rbf(t) + bf2f(p.out[n]));  // down_proj + residual (mq2vl.py:660)
        }
    }
}

template <typename Kern>
static int set_smem(Kern kern, int bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : -1;
}

static int check_common(const GemvParams& p) {
    if (p.N <= 0 || p.K <= 0 || (p.K % 8) || (p.ldw % 8)) return -1;
    if (p.K * 2 > 200 * 1024) return -2;
    return 0;
}

#define LCC_LAUNCH_B(kern, grid, block, smem)                                                          \
    do {                                                                                               \
        static SmemAttrOnce once_;                                                                     \
        if (ensure_dyn_smem(once_, kern, 200 * 1024)) return -3;                                       \
        count_launch();                                                                                \
        if (launch_kernel(kern, dim3(grid), dim3(block), (size_t)(smem), s, pdl, p) != cudaSuccess) return -4; \
    } while (0)
#define LCC_LAUNCH(kern, grid, smem) LCC_LAUNCH_B(kern, grid, 256, smem)

// qkv = W_qkv * rmsnorm(h) + b          (mq2vl.py:631, 559-565)
int gemv_norm_bias(const bf16* W, int ldw, const bf16* x, const bf16* norm_w, float eps, const bf16* bias,
                   bf16* out, int N, int K, const int* finished, const void* pf_ptr, size_t pf_bytes, int num_sms, bool pdl,
                   cudaStream_t s) {
    GemvParams p{}; p.W = W; p.ldw = ldw; p.x = x; p.norm_w = norm_w; p.eps = eps; p.N = N; p.K = K;
    p.bias = bias; p.out = out; p.finished = finished; p.pf_ptr = (const uint8_t*)pf_ptr; p.pf_bytes = (unsigned)pf_bytes;
    if (int r = check_common(p)) return r;
    // measured on B200 (tools/bench_kernel.py sweep): 2 rows/warp x 4 loads x 4 warps = 12.8 us at N=4608,K=3584;
    // every shape tried lands in 12.8-17 us: the kernel is bound by launch + first-load + epilogue latency
    int rpw = 2, unroll = 4, warps = 4;  // tuning hooks: rows per warp, loads in flight per row, warps per CTA
    if (const char* e = getenv("LIVECC_SMALL_RPW")) rpw = atoi(e);
    if (const char* e = getenv("LIVECC_SMALL_UNROLL")) unroll = atoi(e);
    if (const char* e = getenv("LIVECC_SMALL_WARPS")) warps = atoi(e);
    const int rows_cta = rpw * warps;
    if (rpw == 1 && unroll == 4) LCC_LAUNCH_B((gemv_rows_kernel<1, true, GV_BIAS, 4>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
    else if (rpw == 1) LCC_LAUNCH_B((gemv_rows_kernel<1, true, GV_BIAS, 8>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
    else if (unroll == 4) LCC_LAUNCH_B((gemv_rows_kernel<2, true, GV_BIAS, 4>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
    else LCC_LAUNCH_B((gemv_rows_kernel<2, true, GV_BIAS, 8>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
    return 0;
}

// h += W * x                                (o_proj mq2vl.py:593,645; down_proj :504,660)
int gemv_residual(const bf16* W, int ldw, const bf16* x, bf16* h_inout, int N, int K, const int* finished,
                  const void* pf_ptr, size_t pf_bytes, int num_sms, bool pdl, cudaStream_t s) {
    GemvParams p{}; p.W = W; p.ldw = ldw; p.x = x; p.N = N; p.K = K; p.out = h_inout; p.finished = finished;
    p.pf_ptr = (const uint8_t*)pf_ptr; p.pf_bytes = (unsigned)pf_bytes;
    if (int r = check_common(p)) return r;
    if (K > 8192) {
        // 4 rows per CTA measured best on B200 (25.9 us vs 26.7-27.1 for 5/6/8 rows at N=3584, K=18944)
        int unroll = 2;
        if (const char* e = getenv("LIVECC_SPLITK_UNROLL")) unroll = atoi(e);  // tuning hook
        if (unroll == 4) LCC_LAUNCH((gemv_splitk_kernel<4, 4>), (N + 3) / 4, 0);
        else if (unroll == 3) LCC_LAUNCH((gemv_splitk_kernel<4, 3>), (N + 3) / 4, 0);
        else LCC_LAUNCH((gemv_splitk_kernel<4, 2>), (N + 3) / 4, 0);
    } else {
        int rpw = 0, unroll = 8, warps = 8;
        if (const char* e = getenv("LIVECC_SMALL_RPW")) rpw = atoi(e);
        if (const char* e = getenv("LIVECC_SMALL_UNROLL")) unroll = atoi(e);
        if (const char* e = getenv("LIVECC_SMALL_WARPS")) warps = atoi(e);
        if (rpw == 0) { rpw = 1; unroll = 8; warps = 8; }  // measured best (11.1 us at N=K=3584; sweep range 11-15 us)
        const int rows_cta = rpw * warps;
        if (rpw == 1 && unroll == 4) LCC_LAUNCH_B((gemv_rows_kernel<1, false, GV_RESIDUAL, 4>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
        else if (rpw == 1) LCC_LAUNCH_B((gemv_rows_kernel<1, false, GV_RESIDUAL, 8>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
        else if (unroll == 4) LCC_LAUNCH_B((gemv_rows_kernel<2, false, GV_RESIDUAL, 4>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
        else LCC_LAUNCH_B((gemv_rows_kernel<2, false, GV_RESIDUAL, 8>), (N + rows_cta - 1) / rows_cta, warps * 32, K * 2);
    }
    return 0;
}

// act = silu(Wg * rmsnorm(h)) * (Wu * rmsnorm(h)), gate/up rows interleaved by 16   (mq2vl.py:502-504, 657-659)
int gemv_norm_swiglu(const bf16* W_gu, int ldw, const bf16* x, const bf16* norm_w, float eps, bf16* act,
                     int N2 /* = 2*I */, int K, const int* finished, const void* pf_ptr, size_t pf_bytes, int num_sms,
                     bool pdl, cudaStream_t s) {
    GemvParams p{}; p.W = W_gu; p.ldw = ldw; p.x = x; p.norm_w = norm_w; p.eps = eps; p.N = N2; p.K = K;
    p.out = act; p.finished = finished; p.pf_ptr = (const uint8_t*)pf_ptr; p.pf_bytes = (unsigned)pf_bytes;
    if (int r = check_common(p)) return r;
    if (N2 % 32) return -5;
    LCC_LAUNCH((gemv_rows_kernel<2, true, GV_SWIGLU, 4>), (N2 + 15) / 16, K * 2);
    return 0;
}

// logits = float(bf16(W_lm * rmsnorm(h)))   (mq2vl.py:905, 1437-1438; gen/utils.py:2762)
int gemv_norm_logits(const bf16* W, int ldw, const bf16* x, const bf16* norm_w, float eps, float* logits,
                     float* logits_copy, int N, int K, const int* finished, int num_sms, bool pdl, cudaStream_t s) {
    GemvParams p{}; p.W = W; p.ldw = ldw; p.x = x; p.norm_w = norm_w; p.eps = eps; p.N = N; p.K = K;
    p.out_f32 = logits; p.out_f32_b = logits_copy; p.finished = finished;
    if (int r = check_common(p)) return r;
    LCC_LAUNCH((gemv_rows_kernel<4, true, GV_LOGITS, 4>), (N + 31) / 32, K * 2);
    return 0;
}

}  // namespace lcc
