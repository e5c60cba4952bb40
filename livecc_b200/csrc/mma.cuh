// mma.sync / ldmatrix / cp.async helpers and the warp-level online-softmax step shared by the mma.sync attention
// kernels (attention.cu) and the persistent decode-step kernel (decode_mega.cu).
#pragma once
#include "common.cuh"

namespace lcc {

// ---------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = smem_u32(smem);
    const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                            const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Online softmax update for one warp tile. s: raw scores (fp32) of 16 rows x 16*NT16 columns,
// already masked with -inf. Each thread owns rows (lane/4) and (lane/4 + 8). Converts P to bf16
// A-fragments and rescales O.
template <int D, int NT16>
__device__ __forceinline__ void softmax_step(float (&s)[NT16 * 2][4], float scale_log2, float (&m)[2],
                                             float (&l)[2], float (&o)[D / 8][4], uint32_t (&pf)[NT16][4]) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < NT16 * 2; ++n) {
        mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float mnew[2], corr[2], msub[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mnew[r] = fmaxf(m[r], mx[r]);
        msub[r] = (mnew[r] == -INFINITY) ? 0.f : mnew[r] * scale_log2;
        corr[r] = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] * scale_log2 - msub[r]);
        m[r] = mnew[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < NT16 * 2; ++n) {
        s[n][0] = exp2f(s[n][0] * scale_log2 - msub[0]);
        s[n][1] = exp2f(s[n][1] * scale_log2 - msub[0]);
        s[n][2] = exp2f(s[n][2] * scale_log2 - msub[1]);
        s[n][3] = exp2f(s[n][3] * scale_log2 - msub[1]);
        rs[0] += s[n][0] + s[n][1];
        rs[1] += s[n][2] + s[n][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l[r] = l[r] * corr[r] + rs[r];  // per-thread partial row sums
#pragma unroll
    for (int d = 0; d < D / 8; ++d) {
        o[d][0] *= corr[0]; o[d][1] *= corr[0];
        o[d][2] *= corr[1]; o[d][3] *= corr[1];
    }
#pragma unroll
    for (int kt = 0; kt < NT16; ++kt) {
        pf[kt][0] = pack_bf16x2(s[kt * 2][0], s[kt * 2][1]);
        pf[kt][1] = pack_bf16x2(s[kt * 2][2], s[kt * 2][3]);
        pf[kt][2] = pack_bf16x2(s[kt * 2 + 1][0], s[kt * 2 + 1][1]);
        pf[kt][3] = pack_bf16x2(s[kt * 2 + 1][2], s[kt * 2 + 1][3]);
    }
}

}  // namespace lcc
