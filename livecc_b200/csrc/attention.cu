// Attention kernels of the LiveCC path, mma.sync.m16n8k16 bf16 flash-style with cp.async-staged K/V tiles.
// The ViT and prefill kernels here were the round's first implementation and are now the fallback / cross-check of
// the tcgen05 + TMEM kernels (attention_tc.cu, attention_prefill_tc.cu), which the launchers below dispatch to by
// default; the one-token decode attention (HBM-bound) lives only here.
//
//  * flash_fwd_kernel<D, CAUSAL, PAGED>: multi-row attention.
//      - ViT (D=80, non-causal, K/V read from the fused qkv buffer, one cu_seqlens segment per
//        blockIdx.z):  VisionAttention, mq2vl.py:392-454.
//      - decoder prefill (D=128, causal over past+new, K/V read from the paged cache, the G query
//        heads of a KV group packed into the row dimension so a K/V tile is loaded once per group):
//        Qwen2VLAttention, mq2vl.py:572-594 + eager_attention_forward :353-375.
//  * attn_decode_kernel: one new token, split-KV over the paged cache; the 7 query heads of a KV
//    group form one 16-row MMA tile; fuses the 1-D RoPE of q, the RoPE + append of the new k/v.
//    The last CTA of a KV group to finish merges the split-KV partials (no separate combine launch).
// Softmax statistics are fp32; P is rounded to bf16 before P·V (as the reference does, :370).
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"
#include "mma.cuh"
#include "ops.h"

namespace lcc {

// One warp: S[16 x 16*NT16] = Q_frag (16 x D) * K_tile^T.  K_tile rows (tokens) start at `ks`
// (row stride LDS elements); NT16 groups of 16 tokens.
template <int D, int LDS, int NT16>
__device__ __forceinline__ void warp_qk(const uint32_t (&qf)[D / 16][4], const bf16* ks, int lane,
                                        float (&s)[NT16 * 2][4]) {
#pragma unroll
    for (int n = 0; n < NT16 * 2; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[n][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
        for (int nt = 0; nt < NT16; ++nt) {
            // x4: matrices (tok 0-7,k 0-7) (tok 0-7,k 8-15) (tok 8-15,k 0-7) (tok 8-15,k 8-15)
            const int mat = lane >> 3, r = lane & 7;
            const bf16* p = ks + (size_t)(nt * 16 + (mat >> 1) * 8 + r) * LDS + kk * 16 + (mat & 1) * 8;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, p);
            mma_bf16_16816(s[nt * 2], qf[kk], b0, b1);
            mma_bf16_16816(s[nt * 2 + 1], qf[kk], b2, b3);
        }
    }
}

// One warp: O[16 x D] += P (16 x 16*NT16, bf16 A fragments) * V_tile (tokens x D).
template <int D, int LDS, int NT16>
__device__ __forceinline__ void warp_pv(const uint32_t (&pf)[NT16][4], const bf16* vs, int lane,
                                        float (&o)[D / 8][4]) {
#pragma unroll
    for (int kt = 0; kt < NT16; ++kt) {
#pragma unroll
        for (int dn = 0; dn < D / 16; ++dn) {
            // trans x4: matrices (tok 0-7,d 0-7) (tok 8-15,d 0-7) (tok 0-7,d 8-15) (tok 8-15,d 8-15)
            const int mat = lane >> 3, r = lane & 7;
            const bf16* p = vs + (size_t)(kt * 16 + (mat & 1) * 8 + r) * LDS + dn * 16 + (mat >> 1) * 8;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(b0, b1, b2, b3, p);
            mma_bf16_16816(o[dn * 2], pf[kt], b0, b1);
            mma_bf16_16816(o[dn * 2 + 1], pf[kt], b2, b3);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Multi-row flash attention. CTA = 4 warps = 64 packed rows; KV tiles of 64 tokens.
// Packed row r of KV group g maps to (query position r / G, query head g*G + r % G).
// ---------------------------------------------------------------------------------------------
struct FlashParams {
    const bf16* q;      // [S, *] row stride q_ld; head h at column h*D
    int q_ld;
    const bf16* k;      // contiguous mode: row stride kv_ld, head g at column g*D
    const bf16* v;
    int kv_ld;
    const bf16* kc;     // paged mode: [pages, Hkv, P=64, D]
    const bf16* vc;
    const int* page_table;
    int Hkv;
    bf16* out;          // [S, *] row stride o_ld; head h at column h*D
    int o_ld;
    const int* cu_seqlens;  // contiguous mode: segment boundaries (blockIdx.z); null => one segment [0,S)
    int S;              // number of query positions (paged mode / single segment)
    int past;           // paged+causal: number of cached tokens before the S new ones
    int G;              // query heads per KV head
    float scale_log2;   // softmax scale * log2(e)
    // split-KV (paged mode only): blockIdx.z = split; partials are merged by flash_merge_kernel
    int nsplit;
    float* part_o;      // [nsplit, Hkv, S*G, D] unnormalised fp32
    float* part_ml;     // [nsplit, Hkv, S*G, 2]
};

template <int D, bool CAUSAL, bool PAGED>
__global__ void __launch_bounds__(128) flash_fwd_kernel(const FlashParams p) {
    constexpr int LDS = D + 8;
    constexpr int BM = 64, BN = 64;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    bf16* sq = reinterpret_cast<bf16*>(smem_attn);           // [64][LDS]
    bf16* sk = sq + BM * LDS;                                 // [2][64][LDS]
    bf16* sv = sk + 2 * BN * LDS;                             // [2][64][LDS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;  // KV head / group
    int seg_start = 0, seg_len = p.S;
    if (!PAGED && p.cu_seqlens) {
        seg_start = p.cu_seqlens[blockIdx.z];
        seg_len = p.cu_seqlens[blockIdx.z + 1] - seg_start;
    }
    const int rows_total = seg_len * p.G;
    const int row0 = blockIdx.x * BM;
    if (row0 >= rows_total) return;

    // ---- stage Q tile (zero-filled beyond rows_total) ----
    for (int i = threadIdx.x; i < BM * (D / 8); i += 128) {
        const int r = i / (D / 8), c = i % (D / 8);
        const int rr = row0 + r;
        const bool ok = rr < rows_total;
        const int pos = ok ? rr / p.G : 0, hq = g * p.G + (ok ? rr % p.G : 0);
        cp_async16(sq + r * LDS + c * 8, p.q + (size_t)(seg_start + pos) * p.q_ld + (size_t)hq * D + c * 8, ok);
    }
    cp_async_commit();

    // KV range for this CTA
    int kv_len = PAGED ? (p.past + p.S) : seg_len;
    if (CAUSAL) {
        const int last_row = min(row0 + BM, rows_total) - 1;
        kv_len = min(kv_len, p.past + last_row / p.G + 1);
    }
    const int n_tiles = (kv_len + BN - 1) / BN;
    int tile_lo = 0, tile_hi = n_tiles;
    const bool split = PAGED && p.nsplit > 1;
    if (split) {
        const int tiles_total = (p.past + p.S + BN - 1) / BN;
        const int tps = (tiles_total + p.nsplit - 1) / p.nsplit;
        tile_lo = blockIdx.z * tps;
        tile_hi = min(n_tiles, tile_lo + tps);
        if (tile_lo >= tile_hi) {  // nothing to do for this split: neutral partial (m = -inf, l = 0)
            cp_async_commit();
            cp_async_wait<0>();
            float* ml = p.part_ml + (((size_t)blockIdx.z * p.Hkv + g) * rows_total + row0) * 2;
            for (int i = threadIdx.x; i < min(BM, rows_total - row0); i += 128) { ml[i * 2] = -INFINITY; ml[i * 2 + 1] = 0.f; }
            return;
        }
    }

    auto load_kv = [&](int tile, int buf) {
        bf16* dk = sk + buf * BN * LDS;
        bf16* dv = sv + buf * BN * LDS;
        const int t0 = tile * BN;
        const bf16 *gk, *gv;
        size_t stride;
        if (PAGED) {
            const int page = p.page_table[tile];
            gk = p.kc + ((size_t)page * p.Hkv + g) * BN * D;
            gv = p.vc + ((size_t)page * p.Hkv + g) * BN * D;
            stride = D;
        } else {
            gk = p.k + (size_t)(seg_start + t0) * p.kv_ld + (size_t)g * D;
            gv = p.v + (size_t)(seg_start + t0) * p.kv_ld + (size_t)g * D;
            stride = p.kv_ld;
        }
        for (int i = threadIdx.x; i < BN * (D / 8); i += 128) {
            const int r = i / (D / 8), c = i % (D / 8);
            const bool ok = (t0 + r) < kv_len;
            cp_async16(dk + r * LDS + c * 8, gk + (size_t)r * stride + c * 8, ok);
            cp_async16(dv + r * LDS + c * 8, gv + (size_t)r * stride + c * 8, ok);
        }
    };

    if (tile_lo < tile_hi) load_kv(tile_lo, 0);
    cp_async_commit();

    cp_async_wait<1>();  // Q landed
    __syncthreads();
    uint32_t qf[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
        const int mat = lane >> 3, r = lane & 7;
        const bf16* ptr = sq + (size_t)(warp * 16 + (mat & 1) * 8 + r) * LDS + kk * 16 + (mat >> 1) * 8;
        ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], ptr);
    }

    float o[D / 8][4];
#pragma unroll
    for (int d = 0; d < D / 8; ++d)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[d][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

    // causal limit (inclusive key index) of this thread's two rows
    int lim[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = row0 + warp * 16 + (lane >> 2) + r * 8;
        lim[r] = CAUSAL ? (p.past + rr / p.G) : (kv_len - 1);
    }

    for (int t = tile_lo; t < tile_hi; ++t) {
        const int buf = (t - tile_lo) & 1;
        if (t + 1 < tile_hi) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const bf16* ks = sk + buf * BN * LDS;
        const bf16* vs = sv + buf * BN * LDS;
        float s[8][4];
        warp_qk<D, LDS, 4>(qf, ks, lane, s);
        // mask: key index beyond the causal limit or beyond kv_len
        const int kbase = t * BN + 2 * (lane & 3);
        const bool need_mask = CAUSAL ? true : ((t + 1) * BN > kv_len);
        if (need_mask) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int kidx = kbase + n * 8;
                if (kidx > lim[0] || kidx >= kv_len) s[n][0] = -INFINITY;
                if (kidx + 1 > lim[0] || kidx + 1 >= kv_len) s[n][1] = -INFINITY;
                if (kidx > lim[1] || kidx >= kv_len) s[n][2] = -INFINITY;
                if (kidx + 1 > lim[1] || kidx + 1 >= kv_len) s[n][3] = -INFINITY;
            }
        }
        uint32_t pf[4][4];
        softmax_step<D, 4>(s, p.scale_log2, m, l, o, pf);
        warp_pv<D, LDS, 4>(pf, vs, lane, o);
        __syncthreads();  // all warps done with this buffer before it is refilled
    }

    // finalize: full row sums across the quad, normalise, store bf16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    if (split) {  // unnormalised partial + (m, l) for the merge kernel
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int rr = row0 + warp * 16 + (lane >> 2) + r * 8;
            if (rr >= rows_total) continue;
            const size_t prow = ((size_t)blockIdx.z * p.Hkv + g) * rows_total + rr;
            float* dst = p.part_o + prow * D + 2 * (lane & 3);
#pragma unroll
            for (int d = 0; d < D / 8; ++d)
                *reinterpret_cast<float2*>(dst + d * 8) = make_float2(o[d][2 * r], o[d][2 * r + 1]);
            if ((lane & 3) == 0) { p.part_ml[prow * 2] = m[r]; p.part_ml[prow * 2 + 1] = l[r]; }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = row0 + warp * 16 + (lane >> 2) + r * 8;
        if (rr >= rows_total) continue;
        const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
        const int pos = rr / p.G, hq = g * p.G + rr % p.G;
        bf16* dst = p.out + (size_t)(seg_start + pos) * p.o_ld + (size_t)hq * D + 2 * (lane & 3);
#pragma unroll
        for (int d = 0; d < D / 8; ++d) {
            *reinterpret_cast<uint32_t*>(dst + d * 8) = pack_bf16x2(o[d][2 * r] * inv, o[d][2 * r + 1] * inv);
        }
    }
}

// Merge of the split-KV partials of the prefill attention: one CTA per packed row, one thread per dim.
__global__ void __launch_bounds__(128) flash_merge_kernel(const float* __restrict__ part_o,
                                                          const float* __restrict__ part_ml, int nsplit, int Hkv,
                                                          int rows_total, int G, float scale_log2,
                                                          bf16* __restrict__ out, int o_ld) {
    const int rr = blockIdx.x, g = blockIdx.y, d = threadIdx.x;
    float ms[8], ls[8];
    float M = -INFINITY;
#pragma unroll
    for (int sp = 0; sp < 8; ++sp) {
        if (sp < nsplit) {
            const size_t prow = ((size_t)sp * Hkv + g) * rows_total + rr;
            ms[sp] = part_ml[prow * 2];
            ls[sp] = part_ml[prow * 2 + 1];
            M = fmaxf(M, ms[sp]);
        }
    }
    float acc = 0.f, L = 0.f;
#pragma unroll
    for (int sp = 0; sp < 8; ++sp) {
        if (sp < nsplit && ms[sp] != -INFINITY) {
            const size_t prow = ((size_t)sp * Hkv + g) * rows_total + rr;
            const float f = exp2f((ms[sp] - M) * scale_log2);
            acc += f * part_o[prow * 128 + d];
            L += f * ls[sp];
        }
    }
    const int pos = rr / G, hq = g * G + rr % G;
    out[(size_t)pos * o_ld + (size_t)hq * 128 + d] = f2bf(L > 0.f ? acc / L : 0.f);
}

template <int D, bool CAUSAL, bool PAGED>
static int launch_flash(const FlashParams& p, dim3 grid, cudaStream_t s) {
    constexpr int LDS = D + 8;
    constexpr int smem = (64 + 4 * 64) * LDS * 2;
    auto kern = flash_fwd_kernel<D, CAUSAL, PAGED>;
    static SmemAttrOnce once;  // per instantiation
    if (ensure_dyn_smem(once, kern, smem)) return -1;
    { lcc::count_launch(); kern<<<grid, 128, smem, s>>>(p); }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ViT attention with MT (= 2) row tiles per warp: CTA = 4 warps = 128 query rows. Every K/V fragment loaded
// with ldmatrix feeds MT MMAs, which halves the shared-memory traffic per FLOP of the 64-row kernel above
// (the limiter of the mma.sync path at head_dim 80). Non-causal, K/V read from the fused qkv buffer, one
// cu_seqlens segment per blockIdx.z (VisionAttention, mq2vl.py:392-454).
// ---------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(128) vit_flash_kernel(const FlashParams p) {
    constexpr int D = 80, LDS = D + 8, BN = 64, BM = 64 * MT;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    bf16* sq = reinterpret_cast<bf16*>(smem_attn);  // [BM][LDS]
    bf16* sk = sq + BM * LDS;                        // [2][64][LDS]
    bf16* sv = sk + 2 * BN * LDS;                    // [2][64][LDS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int seg_start = p.cu_seqlens[blockIdx.z];
    const int seg_len = p.cu_seqlens[blockIdx.z + 1] - seg_start;
    const int row0 = blockIdx.x * BM;
    if (row0 >= seg_len) return;

    for (int i = threadIdx.x; i < BM * (D / 8); i += 128) {
        const int r = i / (D / 8), c = i % (D / 8);
        const bool ok = (row0 + r) < seg_len;
        cp_async16(sq + r * LDS + c * 8, p.q + (size_t)(seg_start + (ok ? row0 + r : 0)) * p.q_ld + (size_t)head * D + c * 8, ok);
    }
    cp_async_commit();
    const int n_tiles = (seg_len + BN - 1) / BN;
    auto load_kv = [&](int tile, int buf) {
        bf16* dk = sk + buf * BN * LDS;
        bf16* dv = sv + buf * BN * LDS;
        const int t0 = tile * BN;
        const bf16* gk = p.k + (size_t)(seg_start + t0) * p.kv_ld + (size_t)head * D;
        const bf16* gv = p.v + (size_t)(seg_start + t0) * p.kv_ld + (size_t)head * D;
        for (int i = threadIdx.x; i < BN * (D / 8); i += 128) {
            const int r = i / (D / 8), c = i % (D / 8);
            const bool ok = (t0 + r) < seg_len;
            cp_async16(dk + r * LDS + c * 8, gk + (size_t)(ok ? r : 0) * p.kv_ld + c * 8, ok);
            cp_async16(dv + r * LDS + c * 8, gv + (size_t)(ok ? r : 0) * p.kv_ld + c * 8, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    uint32_t qf[MT][D / 16][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            const int mat = lane >> 3, r = lane & 7;
            const bf16* ptr = sq + (size_t)((warp * MT + mt) * 16 + (mat & 1) * 8 + r) * LDS + kk * 16 + (mat >> 1) * 8;
            ldmatrix_x4(qf[mt][kk][0], qf[mt][kk][1], qf[mt][kk][2], qf[mt][kk][3], ptr);
        }
    float o[MT][D / 8][4];
    float m[MT][2], l[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        m[mt][0] = m[mt][1] = -INFINITY;
        l[mt][0] = l[mt][1] = 0.f;
#pragma unroll
        for (int d = 0; d < D / 8; ++d)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[mt][d][j] = 0.f;
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const bf16* ks = sk + buf * BN * LDS;
        const bf16* vs = sv + buf * BN * LDS;
        float s[MT][8][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[mt][n][j] = 0.f;
        // S = Q K^T : each K fragment feeds MT row tiles
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int mat = lane >> 3, r = lane & 7;
                const bf16* ptr = ks + (size_t)(nt * 16 + (mat >> 1) * 8 + r) * LDS + kk * 16 + (mat & 1) * 8;
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3, ptr);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16_16816(s[mt][nt * 2], qf[mt][kk], b0, b1);
                    mma_bf16_16816(s[mt][nt * 2 + 1], qf[mt][kk], b2, b3);
                }
            }
        }
        if ((t + 1) * BN > seg_len) {  // ragged last tile
            const int kbase = t * BN + 2 * (lane & 3);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int kidx = kbase + n * 8;
                    if (kidx >= seg_len) { s[mt][n][0] = -INFINITY; s[mt][n][2] = -INFINITY; }
                    if (kidx + 1 >= seg_len) { s[mt][n][1] = -INFINITY; s[mt][n][3] = -INFINITY; }
                }
        }
        uint32_t pf[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) softmax_step<D, 4>(s[mt], p.scale_log2, m[mt], l[mt], o[mt], pf[mt]);
        // O += P V : each V fragment feeds MT row tiles
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
            for (int dn = 0; dn < D / 16; ++dn) {
                const int mat = lane >> 3, r = lane & 7;
                const bf16* ptr = vs + (size_t)(kt * 16 + (mat & 1) * 8 + r) * LDS + dn * 16 + (mat >> 1) * 8;
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(b0, b1, b2, b3, ptr);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16_16816(o[mt][dn * 2], pf[mt][kt], b0, b1);
                    mma_bf16_16816(o[mt][dn * 2 + 1], pf[mt][kt], b2, b3);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float lr = l[mt][r];
            lr += __shfl_xor_sync(0xffffffffu, lr, 1);
            lr += __shfl_xor_sync(0xffffffffu, lr, 2);
            const int rr = row0 + (warp * MT + mt) * 16 + (lane >> 2) + r * 8;
            if (rr >= seg_len) continue;
            const float inv = lr > 0.f ? 1.f / lr : 0.f;
            bf16* dst = p.out + (size_t)(seg_start + rr) * p.o_ld + (size_t)head * D + 2 * (lane & 3);
#pragma unroll
            for (int d = 0; d < D / 8; ++d)
                *reinterpret_cast<uint32_t*>(dst + d * 8) = pack_bf16x2(o[mt][d][2 * r] * inv, o[mt][d][2 * r + 1] * inv);
        }
}

// ViT: qkv [N, 3*heads*80]; segments from cu_seqlens (device int32, nseg+1 entries); max_seg_len bounds the grid.
int vit_attention(const bf16* qkv, int ld, int64_t n_rows, bf16* out, int o_ld, const int* cu_seqlens, int nseg,
                  int max_seg_len, int heads, int head_dim, int impl, cudaStream_t s) {
    if (head_dim != 80) return -1;
    if (nseg <= 0 || max_seg_len <= 0) return 0;
    if (impl == 0) {
        const char* e = getenv("LIVECC_B200_VIT_ATTN");  // "mma" forces the mma.sync kernels
        impl = (e && e[0] == 'm') ? 1 : 2;
    }
    if (impl == 2 && cu_seqlens) return vit_attention_tc(qkv, ld, n_rows, out, o_ld, cu_seqlens, nseg, max_seg_len, heads, s);
    FlashParams p{};
    p.q = qkv; p.q_ld = ld;
    p.k = qkv + (size_t)heads * head_dim; p.v = qkv + (size_t)2 * heads * head_dim; p.kv_ld = ld;
    p.out = out; p.o_ld = o_ld; p.cu_seqlens = cu_seqlens; p.S = max_seg_len; p.past = 0; p.G = 1;
    p.Hkv = heads;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
    // 128-row CTAs (two row tiles per warp) once they still give every SM >= 2 CTAs; else the 64-row kernel
    int dev_sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas128 = ((max_seg_len + 127) / 128) * heads * nseg;
    if (cu_seqlens && ctas128 >= 2 * dev_sms) {
        constexpr int smem = (128 + 4 * 64) * 88 * 2;
        static SmemAttrOnce once;
        if (ensure_dyn_smem(once, vit_flash_kernel<2>, smem)) return -2;
        { lcc::count_launch(); vit_flash_kernel<2><<<dim3((max_seg_len + 127) / 128, heads, nseg), 128, smem, s>>>(p); }
        return 0;
    }
    dim3 grid((max_seg_len + 63) / 64, heads, nseg);
    return launch_flash<80, false, false>(p, grid, s);
}

// Decoder prefill: q rows of the fused qkv buffer (already rotated), K/V in the paged cache
// (already containing the S new tokens at positions past..past+S-1).
int attn_prefill_paged(const bf16* q, int q_ld, const bf16* kc, const bf16* vc, const int* page_table,
                       int page_size, int Hq, int Hkv, int S, int past, bf16* out, int o_ld, float* part_o,
                       float* part_ml, size_t part_capacity_rows, int num_sms, int impl, cudaStream_t s) {
    if (page_size != 64) return -1;
    if (S <= 0) return 0;
    if (impl == 0) {
        const char* e = getenv("LIVECC_B200_PREFILL_ATTN");  // "mma" forces the mma.sync kernel
        impl = (e && e[0] == 'm') ? 1 : 2;
    }
    if (impl == 2) {
        int ns = 1;
        if (int r = attn_prefill_tc(q, q_ld, kc, vc, page_table, Hq, Hkv, S, past, out, o_ld, part_o, part_ml,
                                    part_capacity_rows, num_sms, &ns, s))
            return r;
        if (ns > 1) {
            const int G = Hq / Hkv;
            lcc::count_launch();
            flash_merge_kernel<<<dim3(S * G, Hkv), 128, 0, s>>>(part_o, part_ml, ns, Hkv, S * G, G,
                                                                1.4426950408889634f / sqrtf(128.f), out, o_ld);
        }
        return 0;
    }
    FlashParams p{};
    p.q = q; p.q_ld = q_ld; p.kc = kc; p.vc = vc; p.page_table = page_table; p.Hkv = Hkv;
    p.out = out; p.o_ld = o_ld; p.cu_seqlens = nullptr; p.S = S; p.past = past; p.G = Hq / Hkv;
    p.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
    const int q_tiles = (S * p.G + 63) / 64;
    const int ctas = q_tiles * Hkv;
    const int kv_tiles = (past + S + 63) / 64;
    // split the KV range when a chunk-sized prefill over a long cache cannot fill two CTAs per SM
    int nsplit = 1;
    if (part_o && part_ml && ctas < 2 * num_sms) {
        nsplit = (2 * num_sms + ctas - 1) / ctas;
        if (nsplit > 8) nsplit = 8;
        while (nsplit > 1 && kv_tiles < 8 * nsplit) --nsplit;  // keep >= 8 tiles of work per split
        if ((size_t)nsplit * Hkv * S * p.G > part_capacity_rows) nsplit = 1;
    }
    p.nsplit = nsplit; p.part_o = part_o; p.part_ml = part_ml;
    dim3 grid(q_tiles, Hkv, nsplit);
    if (int r = launch_flash<128, true, true>(p, grid, s)) return r;
    if (nsplit > 1)
        { lcc::count_launch(); flash_merge_kernel<<<dim3(S * p.G, Hkv), 128, 0, s>>>(part_o, part_ml, nsplit, Hkv, S * p.G, p.G, p.scale_log2, out, o_ld); }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Decode attention (one token). grid = (Hkv, nsplit); CTA = 4 warps; each warp owns 16 of the 64
// tokens of a page-tile; rows of the MMA tile = the G (<=16) query heads of the KV group.
// Device-side scalars (graph-replay friendly): kv_len = tokens already in the cache; the new token
// is appended at slot kv_len by the CTA whose range contains it.
// ---------------------------------------------------------------------------------------------
struct DecodeAttnParams {
    bf16* qkv;          // [ (Hq + 2Hkv) * 128 ] raw projections (+bias) of the new token
    bf16* kc;           // layer K cache [pages, Hkv, 64, 128]
    bf16* vc;
    const int* page_table;
    const int* kv_len;      // device scalar
    const int* rope_pos;    // device scalar: position id of the new token (kv_len + rope_delta)
    const int* finished;    // device flag: non-zero => no-op
    const float* inv_freq;  // [64]
    int Hq, Hkv, nsplit;
    float* part_o;      // [nsplit, Hq, 128]
    float* part_ml;     // [nsplit, Hq, 2]  (running max (unscaled), l)
    int* counters;      // [Hkv] arrival counters, zero between launches (last CTA of a group merges the splits)
    bf16* out;          // [Hq * 128]
    float scale_log2;
    int* done_groups;   // optional: incremented once per KV group when its merged output is in `out`
};

// o_proj role of the fused attention + o_proj kernel (see attn_oproj_kernel)
struct OprojParams {
    const bf16* W;      // [N, K] o_proj weight
    int ldw, N, K;
    bf16* h;            // residual stream, updated in place
    int* sync;          // [0] = KV groups done (producer flag), [1] = o_proj CTAs finished (for the reset)
    int n_cta;          // number of o_proj CTAs
};

__device__ __forceinline__ void rope1d_row(const bf16* src, bf16* dst, int lane_dim /*0..63*/, float pos,
                                           const float* inv_freq) {
    // one (j, j+64) pair; torch bf16 semantics (each product and the sum rounded to bf16)
    const float ang = __fmul_rn(pos, inv_freq[lane_dim]);
    const float c = rbf(cosf(ang)), s = rbf(sinf(ang));
    const float x1 = bf2f(src[lane_dim]), x2 = bf2f(src[lane_dim + 64]);
    dst[lane_dim] = f2bf(rbf(rbf(x1 * c) + rbf(-x2 * s)));
    dst[lane_dim + 64] = f2bf(rbf(rbf(x2 * c) + rbf(x1 * s)));
}

__device__ __forceinline__ void attn_decode_role(const DecodeAttnParams& p, const int g, const int split,
                                                 uint8_t* smem_attn) {
    constexpr int D = 128, LDS = D + 8, BN = 64;
    pdl_wait();  // qkv of this token comes from the previous kernel
    if (p.finished && *p.finished) return;
    bf16* sq = reinterpret_cast<bf16*>(smem_attn);  // [16][LDS]
    bf16* sk = sq + 16 * LDS;                        // [2][64][LDS]
    bf16* sv = sk + 2 * BN * LDS;                    // [2][64][LDS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.Hq / p.Hkv;
    const int kv_old = *p.kv_len;
    const int T = kv_old + 1;
    const float pos = (float)(*p.rope_pos);

    const int n_tiles_total = (T + BN - 1) / BN;
    const int tiles_per_split = (n_tiles_total + p.nsplit - 1) / p.nsplit;
    const int tile_begin = split * tiles_per_split;
    const int tile_end = min(n_tiles_total, tile_begin + tiles_per_split);

    float* po = p.part_o + ((size_t)split * p.Hq + (size_t)g * G) * D;
    float* pml = p.part_ml + ((size_t)split * p.Hq + (size_t)g * G) * 2;
    const bool empty = tile_begin >= tile_end;
    if (empty) {  // empty split: neutral partial
        if (threadIdx.x < G) { pml[threadIdx.x * 2] = -INFINITY; pml[threadIdx.x * 2 + 1] = 0.f; }
    } else {

    // ---- q: rotate the G heads of this group into smem (rows >= G zero) ----
    for (int i = threadIdx.x; i < 16 * 64; i += 128) {
        const int r = i >> 6, j = i & 63;
        if (r < G) rope1d_row(p.qkv + (size_t)(g * G + r) * D, sq + r * LDS, j, pos, p.inv_freq);
        else { sq[r * LDS + j] = f2bf(0.f); sq[r * LDS + j + 64] = f2bf(0.f); }
    }
    // ---- append the new token's k (rotated) and v to the cache (the CTA owning the last tile) ----
    if (tile_end == n_tiles_total) {
        const int page = p.page_table[kv_old / BN], slot = kv_old % BN;
        bf16* dk = p.kc + (((size_t)page * p.Hkv + g) * BN + slot) * D;
        bf16* dv = p.vc + (((size_t)page * p.Hkv + g) * BN + slot) * D;
        const bf16* ksrc = p.qkv + (size_t)(p.Hq + g) * D;
        const bf16* vsrc = p.qkv + (size_t)(p.Hq + p.Hkv + g) * D;
        if (threadIdx.x < 64) rope1d_row(ksrc, dk, threadIdx.x, pos, p.inv_freq);
        else if (threadIdx.x < 80) {
            const int c = threadIdx.x - 64;
            *reinterpret_cast<uint4*>(dv + c * 8) = *reinterpret_cast<const uint4*>(vsrc + c * 8);
        }
        __threadfence();
    }
    __syncthreads();

    auto load_kv = [&](int tile, int buf) {
        bf16* dk = sk + buf * BN * LDS;
        bf16* dv = sv + buf * BN * LDS;
        const int page = p.page_table[tile];
        const bf16* gk = p.kc + ((size_t)page * p.Hkv + g) * BN * D;
        const bf16* gv = p.vc + ((size_t)page * p.Hkv + g) * BN * D;
        const int t0 = tile * BN;
        for (int i = threadIdx.x; i < BN * (D / 8); i += 128) {
            const int r = i >> 4, c = i & 15;
            const bool ok = (t0 + r) < T;
            cp_async16(dk + r * LDS + c * 8, gk + (size_t)r * D + c * 8, ok);
            cp_async16(dv + r * LDS + c * 8, gv + (size_t)r * D + c * 8, ok);
        }
    };
    load_kv(tile_begin, 0);
    cp_async_commit();

    uint32_t qf[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
        const int mat = lane >> 3, r = lane & 7;
        const bf16* ptr = sq + (size_t)((mat & 1) * 8 + r) * LDS + kk * 16 + (mat >> 1) * 8;
        ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], ptr);
    }
    float o[D / 8][4];
#pragma unroll
    for (int d = 0; d < D / 8; ++d)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[d][j] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

    for (int t = tile_begin; t < tile_end; ++t) {
        const int buf = (t - tile_begin) & 1;
        if (t + 1 < tile_end) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const bf16* ks = sk + buf * BN * LDS + warp * 16 * LDS;
        const bf16* vs = sv + buf * BN * LDS + warp * 16 * LDS;
        float s[2][4];
        warp_qk<D, LDS, 1>(qf, ks, lane, s);
        const int kbase = t * BN + warp * 16 + 2 * (lane & 3);
        if ((t + 1) * BN > T) {
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int kidx = kbase + n * 8;
                if (kidx >= T) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
                if (kidx + 1 >= T) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            }
        }
        uint32_t pf[1][4];
        softmax_step<D, 1>(s, p.scale_log2, m, l, o, pf);
        warp_pv<D, LDS, 1>(pf, vs, lane, o);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }

    // ---- merge the 4 warps (each saw a disjoint token subset) through shared memory ----
    float* so = reinterpret_cast<float*>(sk);  // [4][8][128] fp32 = 16 KB   (only rows < 8 are kept: G <= 8)
    float* sml = so + 4 * 8 * D;               // [4][8][2]
    __syncthreads();
    if ((lane >> 2) < 8) {
        const int r = lane >> 2;  // row (thread's first row); second row r+8 is padding when G <= 8
        float* dst = so + ((size_t)warp * 8 + r) * D + 2 * (lane & 3);
#pragma unroll
        for (int d = 0; d < D / 8; ++d) { dst[d * 8] = o[d][0]; dst[d * 8 + 1] = o[d][1]; }
        if ((lane & 3) == 0) { sml[(warp * 8 + r) * 2] = m[0]; sml[(warp * 8 + r) * 2 + 1] = l[0]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * D; i += 128) {
        const int r = i / D, d = i % D;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, sml[(w * 8 + r) * 2]);
        float acc = 0.f, L = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = sml[(w * 8 + r) * 2];
            const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * p.scale_log2);
            acc += f * so[((size_t)w * 8 + r) * D + d];
            L += f * sml[(w * 8 + r) * 2 + 1];
        }
        po[(size_t)r * D + d] = acc;
        if (d == 0) { pml[r * 2] = M; pml[r * 2 + 1] = L; }
    }
    }  // !empty

    // ---- the last CTA of this KV group to arrive merges all splits (replaces a separate combine launch) ----
    pdl_launch_dependents();
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(&p.counters[g], 1);
        s_last = (prev == p.nsplit - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) p.counters[g] = 0;  // ready for the next launch
    __threadfence();
    float* sm_m = reinterpret_cast<float*>(smem_attn);  // [8][64]
    float* sm_l = sm_m + 8 * 64;                         // [8][64]
    for (int i = threadIdx.x; i < G * p.nsplit; i += 128) {
        const int r = i / p.nsplit, sp = i % p.nsplit;
        const float* ml = p.part_ml + ((size_t)sp * p.Hq + (size_t)g * G + r) * 2;
        sm_m[r * 64 + sp] = __ldcg(ml);
        sm_l[r * 64 + sp] = __ldcg(ml + 1);
    }
    __syncthreads();
    // warp w merges rows w, w+4; a lane owns 4 consecutive dims (float4 loads, all splits independent)
    for (int r = warp; r < G; r += 4) {
        float M = -INFINITY;
        for (int sp = 0; sp < p.nsplit; ++sp) M = fmaxf(M, sm_m[r * 64 + sp]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float L = 0.f;
        const float* base = p.part_o + ((size_t)g * G + r) * D + lane * 4;
        const size_t sp_stride = (size_t)p.Hq * D;
#pragma unroll 8
        for (int sp = 0; sp < p.nsplit; ++sp) {
            const float ms = sm_m[r * 64 + sp];
            const bool live = ms != -INFINITY;
            const float f = live ? exp2f((ms - M) * p.scale_log2) : 0.f;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) v = __ldcg(reinterpret_cast<const float4*>(base + sp * sp_stride));
            acc.x += f * v.x; acc.y += f * v.y; acc.z += f * v.z; acc.w += f * v.w;
            L += f * sm_l[r * 64 + sp];
        }
        const float inv = L > 0.f ? 1.f / L : 0.f;
        uint2 o;
        o.x = pack_bf16x2(acc.x * inv, acc.y * inv);
        o.y = pack_bf16x2(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2*>(p.out + (size_t)(g * G + r) * D + lane * 4) = o;
    }
    if (p.done_groups) {  // publish this group's slice of `out` to the o_proj CTAs of the fused kernel
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(p.done_groups, 1);
    }
}

__global__ void __launch_bounds__(128) attn_decode_kernel(const DecodeAttnParams p) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    attn_decode_role(p, blockIdx.x, blockIdx.y, smem_attn);
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// Fused decode attention + o_proj (+ residual). One launch, two CTA roles:
//   CTAs [0, Hkv*nsplit)        : attention (as above); the last CTA of each KV group bumps sync[0]
//   CTAs [Hkv*nsplit, +n_cta)   : o_proj rows. They are resident from the start, issue the first 16
//                                 weight loads per lane, then wait (bounded spin) until sync[0] == Hkv,
//                                 stage the attention output and finish their dot products.
// This removes one kernel boundary per layer and hides o_proj's launch + first-load latency behind the
// attention. Attention CTAs have the lower block indices, so they are always dispatched first and never
// wait on anything: the o_proj CTAs cannot starve them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot8_attn(const uint4& w, const uint4& x) {
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

__device__ __forceinline__ void oproj_role(const DecodeAttnParams& p, const OprojParams& o, const int cta,
                                           uint8_t* smem) {
    constexpr int ROWS = 4, UNROLL = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = cta * 16 + warp * ROWS;
    const int K = o.K;
    int rows[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) rows[r] = min(row0 + r, o.N - 1);
    uint4 w[ROWS][UNROLL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int cc = lane * 8 + u * 256;
            w[r][u] = (cc < K) ? ld_stream16(o.W + (size_t)rows[r] * o.ldw + cc) : make_uint4(0, 0, 0, 0);
        }
    if (p.finished && *p.finished) return;
    // ---- wait for the merged attention output (bounded: a lost signal degrades the result, never hangs) ----
    if (threadIdx.x == 0) {
        int spins = 0;
        while (ld_acquire_gpu(o.sync) < p.Hkv && ++spins < (1 << 20)) __nanosleep(400);
    }
    __syncthreads();
    bf16* xs = reinterpret_cast<bf16*>(smem);
    for (int c = threadIdx.x * 8; c < K; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(xs + c) = __ldcg(reinterpret_cast<const uint4*>(p.out + c));
    __syncthreads();
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int cc = min(lane * 8 + u * 256, K - 8);
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + cc);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] += dot8_attn(w[r][u], xv);
    }
    for (int c = lane * 8 + UNROLL * 256; (c - lane * 8) < K; c += UNROLL * 256) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = c + u * 256;
                w[r][u] = (cc < K) ? ld_stream16(o.W + (size_t)rows[r] * o.ldw + cc) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int cc = min(c + u * 256, K - 8);
            const uint4 xv = *reinterpret_cast<const uint4*>(xs + cc);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += dot8_attn(w[r][u], xv);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int n = row0 + r;
            if (n < o.N) o.h[n] = f2bf(rbf(acc[r]) + bf2f(o.h[n]));  // o_proj + residual (mq2vl.py:593,645)
        }
    }
    // ---- the last o_proj CTA re-arms the flags for the next launch ----
    __syncthreads();
    if (threadIdx.x == 0) {
        const int c = atomicAdd(o.sync + 1, 1);
        if (c == o.n_cta - 1) {
            o.sync[0] = 0;
            o.sync[1] = 0;
            __threadfence();
        }
    }
}

__global__ void __launch_bounds__(128) attn_oproj_kernel(const DecodeAttnParams p, const OprojParams o) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int n_attn = p.Hkv * p.nsplit;
    if ((int)blockIdx.x < n_attn) attn_decode_role(p, blockIdx.x % p.Hkv, blockIdx.x / p.Hkv, smem_attn);
    else oproj_role(p, o, blockIdx.x - n_attn, smem_attn);
}

int attn_decode(bf16* qkv, bf16* kc, bf16* vc, const int* page_table, int page_size, const int* kv_len,
                const int* rope_pos, const int* finished, const float* inv_freq, int Hq, int Hkv, int nsplit,
                float* part_o, float* part_ml, int* counters, bf16* out, bool pdl, cudaStream_t s) {
    if (page_size != 64 || Hq % Hkv || Hq / Hkv > 8) return -1;
    constexpr int smem = (16 + 4 * 64) * 136 * 2;
    static SmemAttrOnce once;
    if (ensure_dyn_smem(once, attn_decode_kernel, smem)) return -2;
    DecodeAttnParams p{};
    p.qkv = qkv; p.kc = kc; p.vc = vc; p.page_table = page_table; p.kv_len = kv_len; p.rope_pos = rope_pos;
    p.finished = finished; p.inv_freq = inv_freq; p.Hq = Hq; p.Hkv = Hkv; p.nsplit = nsplit;
    p.part_o = part_o; p.part_ml = part_ml; p.counters = counters; p.out = out;
    p.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
    count_launch();
    if (launch_kernel(attn_decode_kernel, dim3(Hkv, nsplit), dim3(128), (size_t)smem, s, pdl, p) != cudaSuccess) return -3;
    return 0;
}

// Fused attention + o_proj launch (see attn_oproj_kernel). sync: device int[2], zero-initialised once.
int attn_oproj_decode(bf16* qkv, bf16* kc, bf16* vc, const int* page_table, int page_size, const int* kv_len,
                      const int* rope_pos, const int* finished, const float* inv_freq, int Hq, int Hkv, int nsplit,
                      float* part_o, float* part_ml, int* counters, bf16* attn_out, const bf16* o_w, int o_ldw,
                      int o_N, bf16* h, int* sync, cudaStream_t s) {
    if (page_size != 64 || Hq % Hkv || Hq / Hkv > 8) return -1;
    constexpr int smem = (16 + 4 * 64) * 136 * 2;
    if (Hq * 128 * 2 > smem) return -4;
    static SmemAttrOnce once;
    if (ensure_dyn_smem(once, attn_oproj_kernel, smem)) return -2;
    DecodeAttnParams p{};
    p.qkv = qkv; p.kc = kc; p.vc = vc; p.page_table = page_table; p.kv_len = kv_len; p.rope_pos = rope_pos;
    p.finished = finished; p.inv_freq = inv_freq; p.Hq = Hq; p.Hkv = Hkv; p.nsplit = nsplit;
    p.part_o = part_o; p.part_ml = part_ml; p.counters = counters; p.out = attn_out;
    p.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
    p.done_groups = sync;
    OprojParams o{};
    o.W = o_w; o.ldw = o_ldw; o.N = o_N; o.K = Hq * 128; o.h = h; o.sync = sync; o.n_cta = (o_N + 15) / 16;
    lcc::count_launch();
    attn_oproj_kernel<<<Hkv * nsplit + o.n_cta, 128, smem, s>>>(p, o);
    return 0;
}

}  // namespace lcc
