// ViT attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 80, non-causal, one cu_seqlens
// segment per blockIdx.z.  VisionAttention, mq2vl.py:392-454.
//
// One CTA = 128 query rows of one (segment, head); TMEM lane = query row, so softmax statistics are
// per-thread (no shuffles).  192 threads:
//   warp 0     TMA producer   Q once, then K/V tiles (128 keys) through a 3-stage ring
//   warp 1     MMA issuer     S[t&1] = Q K_t^T (M128 N128 K80)   and   O += P_t V_t (M128 N80 K128); owns TMEM
//   warps 2-5  softmax        S row -> registers, online softmax with lazy rescale, P (bf16) -> smem, final O
// Operand layout: head_dim 80 = 160 B per row does not fit the 128-byte swizzle atom, so every tile is
// stored as SWIZZLE_32B "slabs" [128 rows][16 bf16]: Q/K slabs are K-major A/B operands (one slab per
// UMMA K step), V slabs form an MN-major B operand (N=80 = 5 slabs, K = keys), P slabs (16 keys each)
// are written by the softmax threads in the same swizzle as a K-major A operand.
// TMEM (512 columns): S0 @0, S1 @128, O @256 (80 columns).
// S is double buffered so Q K_{t+1}^T runs under the softmax of tile t; P V_t runs under the
// max/exp of tile t+1 (only the O rescale and the P store wait for it).
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "launch.h"
#include "ops.h"

namespace lcc {
namespace vtc {
constexpr int D = 80;
constexpr int BM = 128;
constexpr int SLAB_COLS = 16;
constexpr int DSLABS = D / SLAB_COLS;  // 5
constexpr int Q_SLAB_BYTES = BM * 32;
constexpr int Q_BYTES = DSLABS * Q_SLAB_BYTES;
constexpr float RESCALE_THRESHOLD = 8.f;  // log2 units: P stays <= 2^8 between rescales
// P operand of the P.V MMA read from TMEM (tcgen05.st) instead of shared memory. Bit-identical results, 7-11 % faster on B200
// (profiles/r02_vit_attn_ptmem.md). LIVECC_B200_ATTN_PTMEM=0/1 overrides.
constexpr int kPtmemDefault = 1;
// BN = keys per tile. BN=128: one CTA per SM (512 TMEM columns); BN=64: two CTAs per SM (256 columns, <= 113 KB smem)
// PTMEM: P goes to TMEM (tcgen05.st) and P.V reads its A operand from TMEM instead of staging P in shared memory
// (experimental, LIVECC_B200_ATTN_PTMEM=1: written without GPU time at the end of round 1, not validated yet).
template <int BN_, int STAGES_, bool PTMEM_ = false>
struct Cfg {
    static constexpr int BN = BN_, STAGES = STAGES_;
    static constexpr bool PTMEM = PTMEM_;
    static constexpr int KV_SLAB_BYTES = BN * 32;                // [BN keys][16 bf16]
    static constexpr int KV_BYTES = DSLABS * KV_SLAB_BYTES;      // one K or V tile
    static constexpr int PSLABS = BN / SLAB_COLS;                // P: PSLABS slabs of [128 rows][16 keys]
    static constexpr int P_SLAB_BYTES = BM * 32;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + Q_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
    static constexpr int OFF_P = OFF_V + STAGES * KV_BYTES;
    static constexpr int OFF_BAR = OFF_P + PSLABS * P_SLAB_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + barriers + alignment slack
    static constexpr int COL_S = 0, COL_O = 2 * BN;
    static constexpr int COL_P = COL_O + 96;  // BN/2 columns of packed bf16x2 (PTMEM only); 2*BN + 96 + BN/2 <= TMEM_COLS
    static constexpr int TMEM_COLS = (2 * BN + D) > 256 ? 512 : 256;
    static constexpr int MIN_CTAS = TMEM_COLS == 256 ? 2 : 1;
};
}  // namespace vtc

// ex2.approx.ftz: one MUFU, no denormal fix-up code (inputs <= 8, -inf -> +0)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct VitTcParams {
    bf16* out;
    int o_ld;
    const int* cu_seqlens;
    int heads;
    float scale_log2;
};

template <class C>
__global__ void __launch_bounds__(192, C::MIN_CTAS)
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   VitTcParams p) {
    using namespace vtc;
    constexpr int BN = C::BN, STAGES = C::STAGES, PSLABS = C::PSLABS, COL_S = C::COL_S, COL_O = C::COL_O;
    constexpr int OFF_Q = C::OFF_Q, OFF_K = C::OFF_K, OFF_V = C::OFF_V, OFF_P = C::OFF_P, OFF_BAR = C::OFF_BAR;
    constexpr int KV_BYTES = C::KV_BYTES, KV_SLAB_BYTES = C::KV_SLAB_BYTES, P_SLAB_BYTES = C::P_SLAB_BYTES;
    const int seg = blockIdx.z, head = blockIdx.y;
    const int seg_start = p.cu_seqlens[seg];
    const int seg_len = p.cu_seqlens[seg + 1] - seg_start;
    const int q0 = blockIdx.x * BM;
    if (q0 >= seg_len) return;  // uniform for the CTA
    const int T = (seg_len + BN - 1) / BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // STAGES
    uint64_t* kv_empty = kv_full + STAGES;   // STAGES
    uint64_t* s_full = kv_empty + STAGES;    // 2
    uint64_t* s_empty = s_full + 2;          // 2
    uint64_t* p_full = s_empty + 2;          // 1
    uint64_t* pv_done = p_full + 1;          // 1
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dim = p.heads * D;
    const int col_q = head * D, col_k = dim + head * D, col_v = 2 * dim + head * D;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_q);
        prefetch_tensormap(&tmap_kv);
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
        tmem_alloc(tmem_holder, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int j = 0; j < DSLABS; ++j)
                tma_load_2d(smem + OFF_Q + j * Q_SLAB_BYTES, &tmap_q, q_full, col_q + j * SLAB_COLS, seg_start + q0);
            for (int t = 0; t < T; ++t) {
                const int st = t % STAGES;
                mbar_wait(&kv_empty[st], (((t / STAGES) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[st], 2 * KV_BYTES);
                const int row = seg_start + t * BN;
#pragma unroll
                for (int j = 0; j < DSLABS; ++j)
                    tma_load_2d(smem + OFF_K + st * KV_BYTES + j * KV_SLAB_BYTES, &tmap_kv, &kv_full[st],
                                col_k + j * SLAB_COLS, row);
#pragma unroll
                for (int j = 0; j < DSLABS; ++j)
                    tma_load_2d(smem + OFF_V + st * KV_BYTES + j * KV_SLAB_BYTES, &tmap_kv, &kv_full[st],
                                col_v + j * SLAB_COLS, row);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BN);
        constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D) | (1u << 16);  // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(smem + OFF_Q);
        const uint32_t p_addr = smem_u32(smem + OFF_P);
        auto issue_qk = [&](int t) {
            const int st = t % STAGES, b = t & 1;
            mbar_wait(&kv_full[st], (t / STAGES) & 1);
            mbar_wait(&s_empty[b], (((t >> 1) & 1) ^ 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t k_addr = smem_u32(smem + OFF_K + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < DSLABS; ++kk)
                    umma_bf16_ss(tmem_base + COL_S + b * BN, make_sw32_desc(q_addr + kk * Q_SLAB_BYTES, 16, 256),
                                 make_sw32_desc(k_addr + kk * KV_SLAB_BYTES, 16, 256), idesc_qk, kk ? 1u : 0u);
                umma_commit(&s_full[b]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_qk(0);
        for (int t = 0; t < T; ++t) {
            if (t + 1 < T) issue_qk(t + 1);
            mbar_wait(p_full, t & 1);
            tc_fence_after();
            if (lane == 0) {
                const int st = t % STAGES;
                const uint32_t v_addr = smem_u32(smem + OFF_V + st * KV_BYTES);
#pragma unroll
                for (int ks = 0; ks < PSLABS; ++ks) {
                    const uint64_t dv = make_sw32_desc(v_addr + ks * (SLAB_COLS * 32), KV_SLAB_BYTES, 256);
                    if constexpr (C::PTMEM)  // A = P from TMEM: 16 keys = 8 columns of packed bf16x2 per MMA
                        umma_bf16_ts(tmem_base + COL_O, tmem_base + C::COL_P + ks * 8, dv, idesc_pv, (t | ks) ? 1u : 0u);
                    else
                        umma_bf16_ss(tmem_base + COL_O, make_sw32_desc(p_addr + ks * P_SLAB_BYTES, 16, 256), dv, idesc_pv,
                                     (t | ks) ? 1u : 0u);
                }
                umma_commit(&kv_empty[st]);
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax / correction / epilogue =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        uint8_t* p_row = smem + OFF_P + row * 32;
        const int sw = (row >> 2) & 1;
        float m_used = -INFINITY, l = 0.f;
        for (int t = 0; t < T; ++t) {
            const int b = t & 1;
            mbar_wait(&s_full[b], (t >> 1) & 1);
            tc_fence_after();
            constexpr int NC = BN / 32;
            uint32_t v[NC][32];
#pragma unroll
            for (int c = 0; c < NC; ++c) tmem_ld_32x32b_x32(tmem_base + COL_S + b * BN + c * 32 + lane_off, v[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);

            const int n_valid = seg_len - t * BN;  // >= 1
            if (n_valid < BN) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= n_valid) v[c][j] = __float_as_uint(-INFINITY);
            }
            float mx[8];  // 8 independent chains instead of one BN-deep one
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = __uint_as_float(v[0][j]);
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c > 0 || j >= 8) mx[j & 7] = fmaxf(mx[j & 7], __uint_as_float(v[c][j]));
            float mt = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
            mt *= p.scale_log2;
            const bool need = mt > m_used + RESCALE_THRESHOLD;  // always true on the first tile
            const bool any_need = __any_sync(0xffffffffu, need);
            float alpha = 1.f;
            if (need) {
                alpha = ex2_ftz(m_used - mt);  // 0 on the first tile
                m_used = mt;
                l *= alpha;
            }
            uint32_t pk[BN / 2];
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = ex2_ftz(fmaf(__uint_as_float(v[c][j]), p.scale_log2, -m_used));
                    const float p1 = ex2_ftz(fmaf(__uint_as_float(v[c][j + 1]), p.scale_log2, -m_used));
                    ls[(j >> 1) & 3] += p0 + p1;
                    pk[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
                }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);

            if (t > 0) {
                // P_{t-1} V_{t-1} must have retired: it reads the P buffer and updates O
                mbar_wait(pv_done, (t - 1) & 1);
                tc_fence_after();
                if (any_need) {
                    uint32_t o32[32];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tmem_ld_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o32[j] = __float_as_uint(__uint_as_float(o32[j]) * alpha);
                        tmem_st_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
                    }
                    uint32_t o16[16];
                    tmem_ld_32x32b_x16(tmem_base + COL_O + 64 + lane_off, o16);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) o16[j] = __float_as_uint(__uint_as_float(o16[j]) * alpha);
                    tmem_st_32x32b_x16(tmem_base + COL_O + 64 + lane_off, o16);
                    tmem_st_wait();
                }
            }
            if constexpr (C::PTMEM) {
                // P row -> TMEM: word j of the row = keys (2j, 2j+1), lane = row (the A-operand layout of kind::f16)
#pragma unroll
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t w32[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) w32[j] = pk[c * 32 + j];
                    tmem_st_32x32b_x32(tmem_base + C::COL_P + c * 32 + lane_off, w32);
                }
                tmem_st_wait();
            } else {
                // P row -> smem slabs (16 keys per slab; 16-byte chunk index XOR address bit 7)
#pragma unroll
                for (int ks = 0; ks < PSLABS; ++ks)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int w = ks * 8 + c * 4;
                        *reinterpret_cast<uint4*>(p_row + ks * P_SLAB_BYTES + ((c ^ sw) << 4)) =
                            make_uint4(pk[w], pk[w + 1], pk[w + 2], pk[w + 3]);
                    }
                fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the MMA (async proxy)
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l -> bf16
        mbar_wait(pv_done, (T - 1) & 1);
        tc_fence_after();
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const bool row_ok = q0 + row < seg_len;
        bf16* dst = p.out + (size_t)(seg_start + q0 + row) * p.o_ld + (size_t)head * D;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o32[32];
            tmem_ld_32x32b_x32(tmem_base + COL_O + c * 32 + lane_off, o32);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(o32[g * 8 + 0]) * inv, __uint_as_float(o32[g * 8 + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(o32[g * 8 + 2]) * inv, __uint_as_float(o32[g * 8 + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(o32[g * 8 + 4]) * inv, __uint_as_float(o32[g * 8 + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(o32[g * 8 + 6]) * inv, __uint_as_float(o32[g * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
                }
            }
        }
        uint32_t o16[16];
        tmem_ld_32x32b_x16(tmem_base + COL_O + 64 + lane_off, o16);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(o16[g * 8 + 0]) * inv, __uint_as_float(o16[g * 8 + 1]) * inv);
                o.y = pack_bf16x2(__uint_as_float(o16[g * 8 + 2]) * inv, __uint_as_float(o16[g * 8 + 3]) * inv);
                o.z = pack_bf16x2(__uint_as_float(o16[g * 8 + 4]) * inv, __uint_as_float(o16[g * 8 + 5]) * inv);
                o.w = pack_bf16x2(__uint_as_float(o16[g * 8 + 6]) * inv, __uint_as_float(o16[g * 8 + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + 64 + g * 8) = o;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <class C>
static int launch_tc(const bf16* qkv, int ld, int64_t n_rows, const VitTcParams& p, dim3 grid, cudaStream_t s) {
    using namespace vtc;
    CUtensorMap tq, tkv;
    if (make_tmap_bf16_2d_box(&tq, qkv, n_rows, (int64_t)3 * p.heads * D, ld, SLAB_COLS, BM, /*swizzle32=*/true)) return -10;
    if (make_tmap_bf16_2d_box(&tkv, qkv, n_rows, (int64_t)3 * p.heads * D, ld, SLAB_COLS, C::BN, true)) return -11;
    auto kern = vit_attn_tc_kernel<C>;
    static SmemAttrOnce once;  // per instantiation
    if (ensure_dyn_smem(once, kern, C::SMEM_BYTES)) return -12;
    { lcc::count_launch(); kern<<<grid, 192, C::SMEM_BYTES, s>>>(tq, tkv, p); }
    return 0;
}

// qkv [N, 3*heads*80] (row stride ld); tensor maps over the whole buffer, box = [16 cols, 128 | BN rows], SW32.
int vit_attention_tc(const bf16* qkv, int ld, int64_t n_rows, bf16* out, int o_ld, const int* cu_seqlens, int nseg,
                     int max_seg_len, int heads, cudaStream_t s) {
    using namespace vtc;
    if ((ld % 8) || (o_ld % 8) || !cu_seqlens) return -1;
    VitTcParams p{out, o_ld, cu_seqlens, heads, 1.4426950408889634f / sqrtf((float)D)};
    dim3 grid((max_seg_len + BM - 1) / BM, heads, nseg);
    // <= one wave of CTAs: 128-key tiles, one CTA per SM; more: 64-key tiles so that two CTAs share an SM and one
    // CTA's softmax runs under the other's waits (measured: 103 -> 77 us at 8x1024 patches, 16 heads)
    int dev_sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("LIVECC_B200_VIT_TC_BN");  // tuning hook: force keys per tile (64 | 128)
    const int variant = e ? atoi(e) : ((int)(grid.x * grid.y * grid.z) <= dev_sms ? 128 : 64);
    static int ptmem = -1;
    if (ptmem < 0) {
        const char* pe = getenv("LIVECC_B200_ATTN_PTMEM");  // "1" / "0" force the variant; unset = kPtmemDefault
        ptmem = pe ? (pe[0] == '1' ? 1 : 0) : kPtmemDefault;
    }
    if (ptmem) {
        if (variant == 128) return launch_tc<Cfg<128, 3, true>>(qkv, ld, n_rows, p, grid, s);
        return launch_tc<Cfg<64, 3, true>>(qkv, ld, n_rows, p, grid, s);
    }
    if (variant == 128) return launch_tc<Cfg<128, 3>>(qkv, ld, n_rows, p, grid, s);
    return launch_tc<Cfg<64, 3>>(qkv, ld, n_rows, p, grid, s);
}

}  // namespace lcc
