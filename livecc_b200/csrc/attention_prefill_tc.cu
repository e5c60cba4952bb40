// Decoder prefill attention on tcgen05 + TMEM: causal GQA attention of S new tokens over past+S tokens of the
// paged cache, head_dim 128.  Qwen2VLAttention core, mq2vl.py:572-594 + eager_attention_forward :353-375.
//
// Same structure as the ViT kernel (attention_tc.cu): one CTA = 128 TMEM lanes = 128 packed query rows of one KV
// head, warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax (thread = row).  Differences:
//   * rows are (position, query head of the group) pairs: row r of a tile = position pos0 + r/G, head kvh*G + r%G,
//     so a K/V tile is read once for the G heads that share it.  PT = 128/G positions per tile (G=7: 126 rows used).
//     Q is fetched by ONE 3-D TMA box (64 dims, G heads, PT positions) per 64-dim half, which lands in exactly this
//     row order in the SWIZZLE_128B K-major layout.
//   * a K/V tile = one 64-token page of the cache = two TMA boxes [64 tokens][64 dims] (SWIZZLE_128B): K is a K-major
//     B operand, V an MN-major B operand (N = 128 dims = 2 atoms, K = tokens) - no transposes anywhere.
//   * causal mask per row (key index <= past + position), applied only on tiles that reach the diagonal.
//   * split-KV (blockIdx.z): partial (unnormalised O, m, l) in the format of flash_merge_kernel (attention.cu).
// The tail of the last page beyond past+S is never unmasked; V there must be finite (mrope_kv_write zeroes it).
// TMEM (256 columns): S0 @0, S1 @64, O @128 (128 columns).
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "launch.h"
#include "ops.h"

namespace lcc {
namespace ptc {
constexpr int D = 128, BM = 128, BN = 64, STAGES = 4;
// P operand through TMEM instead of shared memory (see attention_tc.cu). LIVECC_B200_ATTN_PTMEM=0/1 overrides.
constexpr int kPtmemDefault = 1;
constexpr int Q_ATOM_BYTES = BM * 128;       // [128 rows][64 bf16]
constexpr int Q_BYTES = 2 * Q_ATOM_BYTES;
constexpr int KV_ATOM_BYTES = BN * 128;      // [64 tokens][64 bf16]
constexpr int KV_BYTES = 2 * KV_ATOM_BYTES;  // one K or V tile
constexpr int P_BYTES = BM * 128;            // [128 rows][64 keys]
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
constexpr int OFF_P = OFF_V + STAGES * KV_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int COL_S = 0, COL_O = 2 * BN, COL_P = 256;  // P (32 columns of packed bf16x2) only in the PTMEM variant
constexpr float RESCALE_THRESHOLD = 8.f;
}  // namespace ptc

struct PrefillTcParams {
    bf16* out;
    int o_ld;
    const int* page_table;
    int Hkv, G, PT, S, past;
    float scale_log2;
    int nsplit;
    float* part_o;   // [nsplit, Hkv, S*G, 128] unnormalised fp32
    float* part_ml;  // [nsplit, Hkv, S*G, 2]    (raw max, sum)
};

__device__ __forceinline__ float ex2_ftz_p(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// PTMEM: P goes to TMEM (tcgen05.st) and P.V reads its A operand from TMEM instead of staging P in shared memory
// (experimental, LIVECC_B200_ATTN_PTMEM=1: written without GPU time at the end of round 1, not validated yet).
template <bool PTMEM>
__global__ void __launch_bounds__(192, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, PrefillTcParams p) {
    using namespace ptc;
    constexpr int TMEM_COLS = PTMEM ? 512 : 256;
    const int kvh = blockIdx.y, split = blockIdx.z;
    const int pos0 = blockIdx.x * p.PT;
    if (pos0 >= p.S) return;
    const int npos = min(p.PT, p.S - pos0);
    const int rows_used = npos * p.G;
    const int rows_total = p.S * p.G;
    const int row_base = pos0 * p.G;  // packed row index of this tile's row 0
    // KV tiles this CTA needs: keys 0 .. past + last position (causal), restricted to its split
    const int n_tiles = (p.past + pos0 + npos + BN - 1) / BN;
    int tile_lo = 0, tile_hi = n_tiles;
    if (p.nsplit > 1) {
        const int tiles_total = (p.past + p.S + BN - 1) / BN;
        const int tps = (tiles_total + p.nsplit - 1) / p.nsplit;
        tile_lo = split * tps;
        tile_hi = min(n_tiles, tile_lo + tps);
        if (tile_lo >= tile_hi) {  // nothing for this split: neutral partial
            float* ml = p.part_ml + (((size_t)split * p.Hkv + kvh) * rows_total + row_base) * 2;
            for (int i = threadIdx.x; i < rows_used; i += blockDim.x) {
                ml[i * 2] = -INFINITY;
                ml[i * 2 + 1] = 0.f;
            }
            return;
        }
    }
    const int T = tile_hi - tile_lo;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;
    uint64_t* s_empty = s_full + 2;
    uint64_t* p_full = s_empty + 2;
    uint64_t* pv_done = p_full + 1;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_q);
        prefetch_tensormap(&tmap_k);
        prefetch_tensormap(&tmap_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // the box is always (64, G, PT): rows beyond S are zero-filled but still counted
            mbar_arrive_expect_tx(q_full, 2u * 64u * (uint32_t)p.G * (uint32_t)p.PT * 2u);
            tma_load_3d(smem + OFF_Q, &tmap_q, q_full, 0, kvh * p.G, pos0);
            tma_load_3d(smem + OFF_Q + Q_ATOM_BYTES, &tmap_q, q_full, 64, kvh * p.G, pos0);
            for (int i = 0; i < T; ++i) {
                const int st = i % STAGES;
                mbar_wait(&kv_empty[st], (((i / STAGES) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[st], 2 * KV_BYTES);
                const int page = p.page_table[tile_lo + i];
                const int row = (page * p.Hkv + kvh) * BN;
                uint8_t* sk = smem + OFF_K + st * KV_BYTES;
                uint8_t* sv = smem + OFF_V + st * KV_BYTES;
                tma_load_2d(sk, &tmap_k, &kv_full[st], 0, row);
                tma_load_2d(sk + KV_ATOM_BYTES, &tmap_k, &kv_full[st], 64, row);
                tma_load_2d(sv, &tmap_v, &kv_full[st], 0, row);
                tma_load_2d(sv + KV_ATOM_BYTES, &tmap_v, &kv_full[st], 64, row);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BN);
        constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D) | (1u << 16);  // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(smem + OFF_Q);
        const uint32_t p_addr = smem_u32(smem + OFF_P);
        auto issue_qk = [&](int i) {
            const int st = i % STAGES, b = i & 1;
            mbar_wait(&kv_full[st], (i / STAGES) & 1);
            mbar_wait(&s_empty[b], (((i >> 1) & 1) ^ 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t k_addr = smem_u32(smem + OFF_K + st * KV_BYTES);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint64_t dq = make_sw128_kmajor_desc(q_addr + a * Q_ATOM_BYTES);
                    const uint64_t dk = make_sw128_kmajor_desc(k_addr + a * KV_ATOM_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + COL_S + b * BN, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_qk,
                                     (a | k) ? 1u : 0u);
                }
                umma_commit(&s_full[b]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_qk(0);
        for (int i = 0; i < T; ++i) {
            if (i + 1 < T) issue_qk(i + 1);
            mbar_wait(p_full, i & 1);
            tc_fence_after();
            if (lane == 0) {
                const int st = i % STAGES;
                const uint32_t v_addr = smem_u32(smem + OFF_V + st * KV_BYTES);
                const uint64_t dp = make_sw128_kmajor_desc(p_addr);
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks) {  // 16 tokens per MMA: +32 B in P rows, +16 token rows in V
                    const uint64_t dv = make_sw128_mnmajor_desc(v_addr + ks * (16 * 128), KV_ATOM_BYTES);
                    if constexpr (PTMEM)
                        umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_P + ks * 8, dv, idesc_pv, (i | ks) ? 1u : 0u);
                    else
                        umma_bf16_ss(tmem_base + COL_O, dp + (uint64_t)(2 * ks), dv, idesc_pv, (i | ks) ? 1u : 0u);
                }
                umma_commit(&kv_empty[st]);
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax / correction / epilogue =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const bool row_ok = row < rows_used;
        const int rpos = row / p.G, rg = row - rpos * p.G;
        const int lim = row_ok ? p.past + pos0 + rpos : -1;  // last visible key index; invalid rows see nothing
        uint8_t* p_row = smem + OFF_P + row * 128;
        const int sw = row & 7;
        float m_used = -INFINITY, l = 0.f;
        for (int i = 0; i < T; ++i) {
            const int b = i & 1;
            mbar_wait(&s_full[b], (i >> 1) & 1);
            tc_fence_after();
            uint32_t v[2][32];
            tmem_ld_32x32b_x32(tmem_base + COL_S + b * BN + lane_off, v[0]);
            tmem_ld_32x32b_x32(tmem_base + COL_S + b * BN + 32 + lane_off, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);

            const int kbase = (tile_lo + i) * BN;
            if (__any_sync(0xffffffffu, kbase + BN - 1 > lim)) {  // tile reaches the diagonal for some row of this warp
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (kbase + c * 32 + j > lim) v[c][j] = __float_as_uint(-INFINITY);
            }
            float mx[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = __uint_as_float(v[0][j]);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c > 0 || j >= 8) mx[j & 7] = fmaxf(mx[j & 7], __uint_as_float(v[c][j]));
            float mt = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
            mt *= p.scale_log2;
            const bool need = mt > m_used + RESCALE_THRESHOLD;  // false while the row has seen no key (mt = -inf)
            const bool any_need = __any_sync(0xffffffffu, need);
            float alpha = 1.f;
            if (need) {
                alpha = ex2_ftz_p(m_used - mt);  // 0 the first time
                m_used = mt;
                l *= alpha;
            }
            const float m_sub = (m_used == -INFINITY) ? 0.f : m_used;  // all-masked so far: p = ex2(-inf) = 0, not NaN
            uint32_t pk[BN / 2];
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = ex2_ftz_p(fmaf(__uint_as_float(v[c][j]), p.scale_log2, -m_sub));
                    const float p1 = ex2_ftz_p(fmaf(__uint_as_float(v[c][j + 1]), p.scale_log2, -m_sub));
                    ls[(j >> 1) & 3] += p0 + p1;
                    pk[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
                }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);

            if (i > 0) {
                mbar_wait(pv_done, (i - 1) & 1);  // P_{i-1} V_{i-1} retired: P buffer free, O up to date
                tc_fence_after();
                if (any_need) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o32[32];
                        tmem_ld_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o32[j] = __float_as_uint(__uint_as_float(o32[j]) * alpha);
                        tmem_st_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                    }
                    tmem_st_wait();
                }
            }
            if constexpr (PTMEM) {
                // P row -> TMEM: word j = keys (2j, 2j+1), lane = row (the A-operand layout of kind::f16)
                tmem_st_32x32b_x32(tmem_base + COL_P + lane_off, pk);
                tmem_st_wait();
            } else {
                // P row (64 keys = one 128-byte swizzled row: 16-byte chunk index XOR (row & 7))
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(p_row + ((c ^ sw) << 4)) =
                        make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
                fence_proxy_async_smem();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (T - 1) & 1);
        tc_fence_after();
        if (p.nsplit > 1) {
            const size_t prow = ((size_t)split * p.Hkv + kvh) * rows_total + row_base + row;
            float* dst = p.part_o + prow * D;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t o32[32];
                tmem_ld_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        *reinterpret_cast<uint4*>(dst + c * 32 + g4 * 4) =
                            make_uint4(o32[g4 * 4], o32[g4 * 4 + 1], o32[g4 * 4 + 2], o32[g4 * 4 + 3]);
                }
            }
            if (row_ok) {
                p.part_ml[prow * 2] = (m_used == -INFINITY) ? -INFINITY : m_used / p.scale_log2;
                p.part_ml[prow * 2 + 1] = l;
            }
        } else {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            bf16* dst = p.out + (size_t)(pos0 + rpos) * p.o_ld + (size_t)(kvh * p.G + rg) * D;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t o32[32];
                tmem_ld_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(o32[g8 * 8 + 0]) * inv, __uint_as_float(o32[g8 * 8 + 1]) * inv);
                        o.y = pack_bf16x2(__uint_as_float(o32[g8 * 8 + 2]) * inv, __uint_as_float(o32[g8 * 8 + 3]) * inv);
                        o.z = pack_bf16x2(__uint_as_float(o32[g8 * 8 + 4]) * inv, __uint_as_float(o32[g8 * 8 + 5]) * inv);
                        o.w = pack_bf16x2(__uint_as_float(o32[g8 * 8 + 6]) * inv, __uint_as_float(o32[g8 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + c * 32 + g8 * 8) = o;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// q: rotated rows of the fused qkv buffer [S, >= Hq*128] (row stride q_ld); kc/vc: one layer of the page pool
// [pages, Hkv, 64, 128].  Returns the split count used through *nsplit_out (the caller merges when > 1).
int attn_prefill_tc(const bf16* q, int q_ld, const bf16* kc, const bf16* vc, const int* page_table, int Hq, int Hkv,
                    int S, int past, bf16* out, int o_ld, float* part_o, float* part_ml, size_t part_capacity_rows,
                    int num_sms, int* nsplit_out, cudaStream_t s) {
    using namespace ptc;
    const int G = Hq / Hkv;
    if (G < 1 || G > 128 || Hq % Hkv || (q_ld % 8) || (o_ld % 8)) return -1;
    const int PT = BM / G;
    const int q_tiles = (S + PT - 1) / PT;
    const int ctas = q_tiles * Hkv;
    const int kv_tiles = (past + S + BN - 1) / BN;
    int nsplit = 1;
    if (part_o && part_ml && ctas < num_sms) {  // one CTA per SM: split the KV range until the SMs are covered
        nsplit = num_sms / ctas;
        if (nsplit > 8) nsplit = 8;
        while (nsplit > 1 && kv_tiles < 8 * nsplit) --nsplit;  // keep >= 8 tiles of work per split
        if ((size_t)nsplit * Hkv * S * G > part_capacity_rows) nsplit = 1;
    }
    CUtensorMap tq, tk, tv;
    if (make_tmap_bf16_3d_sw128(&tq, q, D, Hq, S, D, q_ld, G, PT)) return -10;
    const int64_t pool_rows = (int64_t)1 << 31;  // page ids come from the page table; no meaningful row bound here
    if (make_tmap_bf16_2d_box(&tk, kc, pool_rows, D, D, 64, BN, false)) return -11;
    if (make_tmap_bf16_2d_box(&tv, vc, pool_rows, D, D, 64, BN, false)) return -11;
    static int ptmem = -1;
    if (ptmem < 0) {
        const char* pe = getenv("LIVECC_B200_ATTN_PTMEM");  // "1" / "0" force the variant; unset = kPtmemDefault
        ptmem = pe ? (pe[0] == '1' ? 1 : 0) : kPtmemDefault;
    }
    static SmemAttrOnce once_a, once_b;
    if (ensure_dyn_smem(once_a, attn_prefill_tc_kernel<false>, SMEM_BYTES) ||
        ensure_dyn_smem(once_b, attn_prefill_tc_kernel<true>, SMEM_BYTES))
        return -12;
    PrefillTcParams p{out, o_ld, page_table, Hkv, G, PT, S, past, 1.4426950408889634f / sqrtf((float)D), nsplit, part_o, part_ml};
    if (ptmem)
        { lcc::count_launch(); attn_prefill_tc_kernel<true><<<dim3(q_tiles, Hkv, nsplit), 192, SMEM_BYTES, s>>>(tq, tk, tv, p); }
    else
        { lcc::count_launch(); attn_prefill_tc_kernel<false><<<dim3(q_tiles, Hkv, nsplit), 192, SMEM_BYTES, s>>>(tq, tk, tv, p); }
    *nsplit_out = nsplit;
    return 0;
}

}  // namespace lcc
