// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     C[M,N] (bf16) = epilogue( A[M,K] (bf16, K-major) x B[N,K]^T (bf16, K-major, nn.Linear layout) )
//
// Used for every M >= 16 contraction of the LiveCC hot path: ViT patch-embed / qkv / proj /
// fc1 / fc2 / merger (mq2vl.py:304-337,385-458) and the decoder prefill projections
// (mq2vl.py:491-504,539-594, lm_head :1437).
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B tiles, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one elected lane issues tcgen05.mma, fp32 accumulators in TMEM)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> bf16 stores);
//                 warp w reads TMEM lane quarter w%4 and the column half (w-4)/4 of the tile
// Pipelines: smem ring (full/empty mbarriers, STAGES deep) between TMA and MMA; a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty) between MMA and epilogue so the epilogue of tile i
// overlaps the main loop of tile i+1.
//
// The fused epilogues reproduce the reference's bf16 rounding points (one torch op = one rounding).
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "launch.h"

namespace lcc {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;

// The N tile is a RUN-TIME value (multiple of 16, 32..256): the 148 SMs are filled by choosing the tile width per
// shape (e.g. ViT qkv (1024 x 3840): 224 -> 8 x 18 = 144 tiles = one full wave; 256 or 128 leave a mostly empty
// second wave). Everything that depended on BLOCK_N (stage size, ring depth, TMEM columns, instruction descriptor,
// epilogue chunking) is derived from it on the host and passed in.
static constexpr int MAX_STAGES = 8;
static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
struct GemmTile {
    int bn;          // N tile
    int stages;      // smem ring depth
    int tmem_cols;   // power of two >= two accumulators (+16 when the last 32-column epilogue chunk is half used)
    uint32_t idesc;  // tcgen05 instruction descriptor (M = 128, N = bn)
    int smem_bytes;
};
static GemmTile make_tile(int bn, bool pair = false) {
    GemmTile t;
    t.bn = bn;
    const int stage_bytes = A_BYTES + (pair ? bn / 2 : bn) * BLOCK_K * 2;
    t.stages = (227 * 1024 - 1024 - 256) / stage_bytes;
    if (t.stages > MAX_STAGES) t.stages = MAX_STAGES;
    int cols = 2 * bn + ((bn & 31) ? 16 : 0);
    t.tmem_cols = 32;
    while (t.tmem_cols < cols) t.tmem_cols <<= 1;
    t.idesc = make_idesc_bf16(pair ? 2 * BLOCK_M : BLOCK_M, bn);
    t.smem_bytes = t.stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
    return t;
}

__device__ __forceinline__ float quick_gelu_bf16(float x) {
    // ACT2FN["quick_gelu"]: input * sigmoid(1.702 * input), every op rounded to bf16
    // (SP/transformers/activations.py:117-123).
    float t1 = rbf(1.702f * x);
    float t2 = rbf(__fdividef(1.0f, 1.0f + __expf(-t1)));
    return rbf(x * t2);
}
__device__ __forceinline__ float gelu_erf_bf16(float x) {
    return rbf(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)));
}
__device__ __forceinline__ float silu_bf16(float x) { return rbf(__fdividef(x, 1.0f + __expf(-x))); }

// Fused epilogue of one 32-column chunk of one accumulator row: v = the fp32 bits of columns [col0, col0 + 32) of output row
// `row` (one thread per row); columns >= tile_end belong to the next tile (or are padding) and are dropped.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], bf16* C, int ldc, int M, int row, bool row_ok, int col0,
                                               int tile_end, int ks, const bf16* __restrict__ bias, const bf16* residual,
                                               int ldr) {
    if (EPI == EPI_PARTIAL_F32) {
        // fp32 partial tile of split ks: P[ks][row][col] (C is the partial buffer, ldc == N)
        if (row_ok) {
            float* dst = reinterpret_cast<float*>(C) + ((size_t)ks * M + row) * (size_t)ldc + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (col0 + g * 4 < tile_end)
                    *reinterpret_cast<uint4*>(dst + g * 4) =
                        make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
    } else if (EPI == EPI_SWIGLU) {
        // Weight rows are interleaved in 32-row groups: 16 gate rows then the 16
        // matching up rows. out[j] = bf16(silu_bf16(bf16 gate) * bf16 up)  (mq2vl.py:503)
        if (row_ok && col0 < tile_end) {  // BLOCK_N % 32 == 0 for this epilogue: whole chunks only
            const int ocol0 = (col0 >> 5) << 4;
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float g0 = rbf(__uint_as_float(v[2 * j])), g1 = rbf(__uint_as_float(v[2 * j + 1]));
                float u0 = rbf(__uint_as_float(v[16 + 2 * j])), u1 = rbf(__uint_as_float(v[16 + 2 * j + 1]));
                o[j] = pack_bf16x2(silu_bf16(g0) * u0, silu_bf16(g1) * u1);
            }
            uint4* dst = reinterpret_cast<uint4*>(C + (size_t)row * ldc + ocol0);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 4 groups of 8 columns (16 B of bf16 each)
            const int col = col0 + g * 8;
            if (row_ok && col < tile_end) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]);
                if (EPI == EPI_BIAS || EPI == EPI_BIAS_QUICKGELU || EPI == EPI_BIAS_GELU ||
                    EPI == EPI_BIAS_RESIDUAL) {
                    const uint4 bb = *reinterpret_cast<const uint4*>(bias + col);
                    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = unpack_bf16x2(bw[j]);
                        x[2 * j] += f.x;
                        x[2 * j + 1] += f.y;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = rbf(x[j]);
                if (EPI == EPI_BIAS_QUICKGELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = quick_gelu_bf16(x[j]);
                } else if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = gelu_erf_bf16(x[j]);
                } else if (EPI == EPI_RESIDUAL || EPI == EPI_BIAS_RESIDUAL) {
                    const uint4 rr =
                        *reinterpret_cast<const uint4*>(residual + (size_t)row * ldr + col);
                    const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = unpack_bf16x2(rw[j]);
                        x[2 * j] += f.x;
                        x[2 * j + 1] += f.y;
                    }
                }
                uint4 o;
                o.x = pack_bf16x2(x[0], x[1]);
                o.y = pack_bf16x2(x[2], x[3]);
                o.z = pack_bf16x2(x[4], x[5]);
                o.w = pack_bf16x2(x[6], x[7]);
                *reinterpret_cast<uint4*>(C + (size_t)row * ldc + col) = o;
            }
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, bf16* C, int M, int N,
                    int K, int ldc, const bf16* __restrict__ bias,
                    const bf16* residual /* may alias C */, int ldr, int splits_arg /* >= 1 */, const GemmTile cfg) {
    const int BLOCK_N = cfg.bn;
    const int STAGE_BYTES = A_BYTES + BLOCK_N * BLOCK_K * 2;
    // split-K exists only in the EPI_PARTIAL_F32 instantiations; everywhere else splits is the constant 1 and the
    // unit arithmetic below folds back to the plain (m tile, n tile) loop.
    const int splits = (EPI == EPI_PARTIAL_F32) ? splits_arg : 1;
    const int STAGES = cfg.stages;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full = empty_bar + MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    // Work unit = (m tile, k split, n tile), m fastest so that the CTAs that share a weight tile run together.
    // splits == 1: a unit is a whole output tile. splits > 1 (EPI_PARTIAL_F32): a unit covers kbps k-blocks and
    // writes an fp32 partial tile; splitk_reduce_kernel sums the partials and applies the epilogue.
    const int num_tiles = m_tiles * n_tiles * splits;
    const int kbps = (num_k_blocks + splits - 1) / splits;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_a);
        prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, cfg.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % m_tiles, ks = (tile / m_tiles) % splits, n_blk = tile / (m_tiles * splits);
                const int kb0 = ks * kbps, kb1 = min(kb0 + kbps, num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = cfg.idesc;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            const int ks = (tile / m_tiles) % splits;
            const int kb0 = ks * kbps, kb1 = min(kb0 + kbps, num_k_blocks);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sb);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance start address by k*32 bytes (>>4 => +2k) inside the swizzle atom
                        umma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                     ((kb - kb0) | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs retire
                    if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== Epilogue =====================
        const int q = warp & 3;          // TMEM lane quarter this warp may access (warp id mod 4)
        const int half = (warp - 4) >> 2;  // which half of the tile's columns this warp converts
        const int chunks = (BLOCK_N + 31) >> 5;  // 32-column epilogue chunks; the last one may be half used (BLOCK_N % 32 == 16)
        const int chunks_h0 = (chunks + 1) >> 1;
        const int c_begin = half ? chunks_h0 : 0, c_end = half ? chunks : chunks_h0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % m_tiles, ks = (tile / m_tiles) % splits, n_blk = tile / (m_tiles * splits);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * BLOCK_M + q * 32 + lane;
            const bool row_ok = row < M;
            const int tile_end = min(N, (n_blk + 1) * BLOCK_N);  // columns past it belong to the next tile (or are padding)
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                const uint32_t taddr =
                    tmem_base + (uint32_t)(acc * BLOCK_N + c * 32) + ((uint32_t)(q * 32) << 16);
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                const int col0 = n_blk * BLOCK_N + c * 32;
                epilogue_chunk<EPI>(v, C, ldc, M, row, row_ok, col0, tile_end, ks, bias, residual, ldr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, cfg.tmem_cols);
    }
}

// ----------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x BLOCK_N output tile per pair of CTAs on a TPC.
// Each CTA stages ITS 128 rows of A and HALF of the B tile (BLOCK_N/2 rows) per k-block: shared-memory traffic per CTA and
// k-block drops from 2*(16 KB + 128*bn) to 2*(16 KB + 64*bn) bytes, which lifts the shared-memory bound of the one-CTA
// kernel (bn = 256: 768 -> 512 clocks per k-block = the tensor floor). The leader CTA (cluster rank 0) issues the MMAs for
// both; accumulators live in both CTAs' TMEM (rows 0-127 in the leader, 128-255 in the peer), each CTA runs the epilogue of
// its own rows. Barriers: `full` lives in the leader and is credited by both CTAs' TMA loads; `empty` and `tmem_full` exist
// in both CTAs and are signalled by the leader's multicast commits; `tmem_empty` lives in the leader and counts the epilogue
// warps of both CTAs.
// ----------------------------------------------------------------------------
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, bf16* C, int M,
                     int N, int K, int ldc, const bf16* __restrict__ bias, const bf16* residual /* may alias C */, int ldr,
                     int splits_arg /* >= 1 */, const GemmTile cfg) {
    const int splits = (EPI == EPI_PARTIAL_F32) ? splits_arg : 1;
    const int BLOCK_N = cfg.bn;
    const int HALF_N = BLOCK_N >> 1;
    const int STAGE_BYTES = A_BYTES + HALF_N * BLOCK_K * 2;
    const int STAGES = cfg.stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full = empty_bar + MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    const int num_units = m_pairs * n_tiles * splits;  // (m pair, k split, n tile), m fastest
    const int kbps = (num_k_blocks + splits - 1) / splits;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_a);
        prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);   // used in the leader: its producer's arrive.expect_tx (bytes of both CTAs)
            mbar_init(&empty_bar[i], 1);  // one multicast commit of the leader's MMA warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);    // one multicast commit
            mbar_init(&tmem_empty[i], 16);  // used in the leader: 8 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_cg2(tmem_holder, cfg.tmem_cols);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    __syncwarp();        // barrier.cluster is .aligned: the warp must be converged
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / TMA credit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                const int m_pair = unit % m_pairs, ks = (unit / m_pairs) % splits, n_blk = unit / (m_pairs * splits);
                const int kb0 = ks * kbps, kb1 = min(kb0 + kbps, num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    tma_load_2d_cg2(sa, &tmap_a, bar, kb * BLOCK_K, m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M);
                    tma_load_2d_cg2(sb, &tmap_b, bar, kb * BLOCK_K, n_blk * BLOCK_N + (int)rank * HALF_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = cfg.idesc;  // M = 256, N = BLOCK_N
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                const int ks = (unit / m_pairs) % splits;
                const int kb0 = ks * kbps, kb1 = min(kb0 + kbps, num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t sb = sa + A_BYTES;
                        const uint64_t da = make_sw128_kmajor_desc(sa);
                        const uint64_t db = make_sw128_kmajor_desc(sb);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16_ss_cg2(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                             ((kb - kb0) | k) != 0 ? 1u : 0u);
                        umma_commit_cg2(&empty_bar[stage], 3);  // frees the slot in both CTAs when the MMAs retire
                        if (kb == kb1 - 1) umma_commit_cg2(&tmem_full[acc], 3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== Epilogue (both CTAs, own 128 rows) =====================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int chunks = (BLOCK_N + 31) >> 5;
        const int chunks_h0 = (chunks + 1) >> 1;
        const int c_begin = half ? chunks_h0 : 0, c_end = half ? chunks : chunks_h0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = pair; unit < num_units; unit += num_pairs) {
            const int m_pair = unit % m_pairs, ks = (unit / m_pairs) % splits, n_blk = unit / (m_pairs * splits);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
            const bool row_ok = row < M;
            const int tile_end = min(N, (n_blk + 1) * BLOCK_N);
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (uint32_t)(acc * BLOCK_N + c * 32) + ((uint32_t)(q * 32) << 16);
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                const int col0 = n_blk * BLOCK_N + c * 32;
                epilogue_chunk<EPI>(v, C, ldc, M, row, row_ok, col0, tile_end, ks, bias, residual, ldr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncwarp();        // lanes that skipped a single-lane role loop wait here for their lane 0
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrives can still touch it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, cfg.tmem_cols);
    }
}

// ----------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
                cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int make_tmap_bf16_2d_box(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                          int box_rows, bool swizzle32) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

int make_tmap_bf16_3d_sw128(CUtensorMap* tm, const void* ptr, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                            int64_t stride2, int box1, int box2) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t gstride[2] = {(cuuint64_t)stride1 * 2, (cuuint64_t)stride2 * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)box1, (cuuint32_t)box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], SW128.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows) {
    return make_tmap_bf16_2d_box(tm, ptr, rows, cols, ld, BLOCK_K, box_rows, false);
}

template <int EPI>
static int launch_cfg(const GemmArgs& a, int bn, int num_sms, cudaStream_t stream) {
    const GemmTile cfg = make_tile(bn);
    CUtensorMap ta, tb;
    if (make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M)) return -10;
    if (make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, bn)) return -11;
    auto kern = gemm_bf16_tn_kernel<EPI>;
    static SmemAttrOnce once;  // per template instantiation
    if (ensure_dyn_smem(once, kern, 227 * 1024)) return -12;
    const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (a.N + bn - 1) / bn;
    const int splits = (EPI == EPI_PARTIAL_F32 && a.splits > 1) ? a.splits : 1;
    const int tiles = m_tiles * n_tiles * splits;
    const int grid = tiles < num_sms ? tiles : num_sms;
    lcc::count_launch();
    kern<<<grid, 384, cfg.smem_bytes, stream>>>(ta, tb, (bf16*)a.C, a.M, a.N, a.K, a.ldc,
                                                (const bf16*)a.bias, (const bf16*)a.residual,
                                                a.ldr, splits, cfg);
    return cudaGetLastError() == cudaSuccess ? 0 : -13;
}

template <int EPI>
static int launch_cfg2(const GemmArgs& a, int bn, int num_sms, cudaStream_t stream) {
    const GemmTile cfg = make_tile(bn, true);
    CUtensorMap ta, tb;
    if (make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M)) return -10;
    if (make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, bn / 2)) return -11;
    auto kern = gemm2_bf16_tn_kernel<EPI>;
    static SmemAttrOnce once;
    if (ensure_dyn_smem(once, kern, 227 * 1024)) return -12;
    const int m_pairs = (a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (a.N + bn - 1) / bn;
    const int splits = (EPI == EPI_PARTIAL_F32 && a.splits > 1) ? a.splits : 1;
    const int units = m_pairs * n_tiles * splits;
    const int pairs = units < num_sms / 2 ? units : num_sms / 2;
    lcc::count_launch();
    kern<<<2 * pairs, 384, cfg.smem_bytes, stream>>>(ta, tb, (bf16*)a.C, a.M, a.N, a.K, a.ldc, (const bf16*)a.bias,
                                                     (const bf16*)a.residual, a.ldr, splits, cfg);
    return cudaGetLastError() == cudaSuccess ? 0 : -13;
}

static int launch_epi2(const GemmArgs& a, int bn, int num_sms, cudaStream_t stream) {
    switch (a.epi) {
        case EPI_NONE: return launch_cfg2<EPI_NONE>(a, bn, num_sms, stream);
        case EPI_BIAS: return launch_cfg2<EPI_BIAS>(a, bn, num_sms, stream);
        case EPI_BIAS_QUICKGELU: return launch_cfg2<EPI_BIAS_QUICKGELU>(a, bn, num_sms, stream);
        case EPI_BIAS_GELU: return launch_cfg2<EPI_BIAS_GELU>(a, bn, num_sms, stream);
        case EPI_RESIDUAL: return launch_cfg2<EPI_RESIDUAL>(a, bn, num_sms, stream);
        case EPI_BIAS_RESIDUAL: return launch_cfg2<EPI_BIAS_RESIDUAL>(a, bn, num_sms, stream);
        case EPI_SWIGLU: return launch_cfg2<EPI_SWIGLU>(a, bn, num_sms, stream);
    }
    return -14;
}

static int launch_epi(const GemmArgs& a, int bn, int num_sms, cudaStream_t stream) {
    switch (a.epi) {
        case EPI_NONE: return launch_cfg<EPI_NONE>(a, bn, num_sms, stream);
        case EPI_BIAS: return launch_cfg<EPI_BIAS>(a, bn, num_sms, stream);
        case EPI_BIAS_QUICKGELU: return launch_cfg<EPI_BIAS_QUICKGELU>(a, bn, num_sms, stream);
        case EPI_BIAS_GELU: return launch_cfg<EPI_BIAS_GELU>(a, bn, num_sms, stream);
        case EPI_RESIDUAL: return launch_cfg<EPI_RESIDUAL>(a, bn, num_sms, stream);
        case EPI_BIAS_RESIDUAL: return launch_cfg<EPI_BIAS_RESIDUAL>(a, bn, num_sms, stream);
        case EPI_SWIGLU: return launch_cfg<EPI_SWIGLU>(a, bn, num_sms, stream);
    }
    return -14;
}

// Tile-width choice. What the B200 measurements say (profiles/r02_gemm_shapes.md, r02_gemm_vs_cublas.md):
//  * with one CTA per SM the main loop is bound by SHARED-MEMORY bandwidth, not by the tensor pipe: per 64-deep k-block the
//    TMA writes 16 KB (A) + bn*128 B (B) into shared memory and the four MMAs read the same bytes back: 2*(16384 + 128*bn) B
//    at 128 B/clk = 256 + 2*bn clocks against the tensor floor of 2*bn (bn = 256: 67 %, 128: 50 %, 64: 33 % — the measured
//    per-k-block times of every shape in the sweep fit this at the ~1.4 GHz the SMs run under tensor load);
//  * every launch carries ~8 us that does not depend on the tile (launch gap, barrier/TMEM set-up, first TMA round trip,
//    the un-overlapped epilogue of the last tile), which is 30-60 % of the M = 281 / 1024 GEMMs of the streaming path;
//  * because of that fixed part, filling the last wave exactly (widths like 224 or 144, which the run-time width allows)
//    measured within noise of the simple rule below (prefill qkv: 96 -> 24.3 us, 128 -> 23.5, 192 -> 25.3, 256 -> 29.5).
// So: 64-wide tiles never win; between 128 and 256 the cost is (waves over the SMs) x (128 + bn), ties to the narrower tile.
static int choose_bn(int M, int N, int num_sms) {
    if (N <= 64) return 64;
    const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    auto cost = [&](int bn) {
        const long long tiles = m_tiles * ((N + bn - 1) / bn);
        return ((tiles + num_sms - 1) / num_sms) * (long long)(128 + bn);
    };
    return cost(128) <= cost(256) ? 128 : 256;
}

// Sum of the split-K partials + the fused epilogue (same rounding points as the in-kernel epilogues):
// C = bf16(sum [+ bias]) [then bf16(. + residual)].  One thread per 8 columns; residual may alias C.
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N,
                                     const bf16* __restrict__ bias, const bf16* residual, int ldr, bf16* C, int ldc,
                                     int epi) {
    const int groups = N / 8;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * groups) return;
    const int row = (int)(idx / groups), col = (int)(idx % groups) * 8;
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ks = 0; ks < splits; ++ks) {
        const float4* src = reinterpret_cast<const float4*>(part + ((size_t)ks * M + row) * (size_t)N + col);
        const float4 a = src[0], b = src[1];
        x[0] += a.x; x[1] += a.y; x[2] += a.z; x[3] += a.w;
        x[4] += b.x; x[5] += b.y; x[6] += b.z; x[7] += b.w;
    }
    if (epi == EPI_BIAS || epi == EPI_BIAS_RESIDUAL) {
        const uint4 bb = *reinterpret_cast<const uint4*>(bias + col);
        const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(bw[j]);
            x[2 * j] += f.x;
            x[2 * j + 1] += f.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = rbf(x[j]);
    if (epi == EPI_RESIDUAL || epi == EPI_BIAS_RESIDUAL) {
        const uint4 rr = *reinterpret_cast<const uint4*>(residual + (size_t)row * ldr + col);
        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(rw[j]);
            x[2 * j] += f.x;
            x[2 * j + 1] += f.y;
        }
    }
    uint4 o;
    o.x = pack_bf16x2(x[0], x[1]);
    o.y = pack_bf16x2(x[2], x[3]);
    o.z = pack_bf16x2(x[4], x[5]);
    o.w = pack_bf16x2(x[6], x[7]);
    *reinterpret_cast<uint4*>(C + (size_t)row * ldc + col) = o;
}

// Split-K for small-M GEMMs whose tile count cannot occupy the SMs (prefill down_proj at M ~ 281: 84 tiles on 148 SMs).
// Returns 1 if it handled the GEMM, 0 if the caller should run the plain path, < 0 on error.
// Validated on B200 in round 2 (bit-exact on integer operands, test_gemm_splitk_exact_integer_operands; whole-model parity
// with it forced on). Measured (profiles/r02_gemm_shapes.md): it pays only for long K — down_proj (K = 18944) 86.7 -> 60.9 us,
// while qkv / o_proj (K = 3584) get slower (23 -> 34 us) — so the default applies it for K >= 8192 only.
// LIVECC_B200_GEMM_SPLITK=1 forces it for every eligible shape, =0 disables it.
static int try_splitk(const GemmArgs& a, int num_sms, cudaStream_t stream) {
    static int mode = -1;  // 0 off, 1 forced, 2 default (long K only)
    if (mode < 0) {
        const char* e = getenv("LIVECC_B200_GEMM_SPLITK");
        mode = !e ? 2 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2));
    }
    if (!mode || !a.splitk_ws || a.M > 384) return 0;
    if (mode == 2 && a.K < 8192) return 0;
    if (a.epi != EPI_NONE && a.epi != EPI_BIAS && a.epi != EPI_RESIDUAL && a.epi != EPI_BIAS_RESIDUAL) return 0;
    const int block_n = a.block_n ? a.block_n : (a.N >= 256 ? 256 : (a.N >= 128 ? 128 : 64));
    const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (a.N + block_n - 1) / block_n;
    const int tiles = m_tiles * n_tiles;
    const int nkb = (a.K + BLOCK_K - 1) / BLOCK_K;
    if (tiles * 2 > num_sms) return 0;
    int splits = (2 * num_sms + tiles / 2) / tiles;  // about two work units per CTA
    if (splits > 8) splits = 8;
    if (splits > nkb / 4) splits = nkb / 4;  // >= 4 k-blocks per unit
    while (splits > 1 && (splits - 1) * ((nkb + splits - 1) / splits) >= nkb) --splits;  // no empty split
    while (splits > 1 && (size_t)splits * a.M * a.N * 4 > a.splitk_ws_bytes) --splits;
    if (splits < 2) return 0;
    GemmArgs p = a;
    p.C = a.splitk_ws;
    p.ldc = a.N;
    p.epi = EPI_PARTIAL_F32;
    p.splits = splits;
    p.bias = nullptr;
    p.residual = nullptr;
    if (int r = launch_cfg<EPI_PARTIAL_F32>(p, block_n, num_sms, stream)) return r;
    const int64_t threads = (int64_t)a.M * (a.N / 8);
    lcc::count_launch();
    splitk_reduce_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(
        (const float*)a.splitk_ws, splits, a.M, a.N, (const bf16*)a.bias, (const bf16*)a.residual, a.ldr, (bf16*)a.C,
        a.ldc, a.epi);
    return cudaGetLastError() == cudaSuccess ? 1 : -13;
}

int gemm_bf16_tn(const GemmArgs& a, int num_sms, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0) return -1;
    if ((a.K % 8) || (a.lda % 8) || (a.ldb % 8) || (a.N % 8) || (a.ldc % 8)) return -2;
    if (a.epi == EPI_SWIGLU && (a.N % 32)) return -3;
    if ((a.epi == EPI_RESIDUAL || a.epi == EPI_BIAS_RESIDUAL) && (!a.residual || (a.ldr % 8))) return -4;
    if ((a.epi == EPI_BIAS || a.epi == EPI_BIAS_QUICKGELU || a.epi == EPI_BIAS_GELU ||
         a.epi == EPI_BIAS_RESIDUAL) && !a.bias) return -5;
    if (int r = try_splitk(a, num_sms, stream)) return r < 0 ? r : 0;
    int block_n = a.block_n;
    if (block_n == 0) {
        static int forced = -1;
        if (forced < 0) {
            const char* e = getenv("LIVECC_B200_GEMM_BN");  // tuning hook: force one width (0/unset = model)
            forced = e ? atoi(e) : 0;
        }
        block_n = forced > 0 ? forced : choose_bn(a.M, a.N, num_sms);
    }
    // block_n < 0 selects the CTA-pair kernel with tile width -block_n (explicit: tests and tuning)
    const bool pair = block_n < 0;
    if (pair) block_n = -block_n;
    if (block_n < 32 || block_n > 256 || (block_n % 16) || (a.epi == EPI_SWIGLU && (block_n % 32))) return -6;
    if (pair) return (block_n % 32) ? -6 : launch_epi2(a, block_n, num_sms, stream);
    return launch_epi(a, block_n, num_sms, stream);
}

}  // namespace lcc
