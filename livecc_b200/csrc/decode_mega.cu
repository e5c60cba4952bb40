// Persistent decode-step kernel ("megakernel"): ONE launch runs all decoder layers + lm_head of one generated token
// for up to 8 independent streams (Qwen2VLDecoderLayer.forward x L + norm + lm_head, mq2vl.py:613-662, 905, 1437;
// the loop body of GenerationMixin._sample, gen/utils.py:2743-2805). It replaces the 6 kernels per layer of the
// per-op path (gemv.cu + attention.cu), whose launch / first-load / tail latency kept the step at 70 % of the HBM
// roofline, and it is what makes multi-stream batching a weight-read-once
// operation.
//
// Structure (grid = #SMs, one CTA per SM, 384 threads):
//   warps 8-11 (one lane each): TMA producers. Each walks the whole step's weight/KV tile sequence of this CTA (a pure
//                        function of the model shape, the streams' KV lengths and the CTA index) and issues every 4th
//                        tile into a shared memory ring of 4 KB tiles (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier
//                        complete_tx). They never wait for a phase boundary: while the consumers sit in a grid barrier
//                        the next phase's weights keep arriving, so HBM stays busy across the 5 dependent phases of a
//                        layer. (One producer thread could not issue 4 KB tiles fast enough: 2.5 TB/s.)
//   warps 0-7          : consumers. A weight tile is 32 rows x 64 k; `mma.sync.m16n8k16` with the weights as the A
//                        operand (ldmatrix from the swizzled tile) and the <= 8 stream activations as the 8 columns of
//                        the B operand, fp32 accumulate. The K range of a row block is split across the 8 warps and
//                        reduced through shared memory in a fixed order, so a stream's result does not depend on how
//                        many other streams share the launch.
//   phases per layer   : qkv (RMSNorm + bias) | attention (RoPE, KV append, paged split-KV, merge) | o_proj + residual
//                        | gate/up (RMSNorm + SwiGLU) | down_proj + residual ; then final norm + lm_head.
//                        Phases are separated by a grid barrier (atomic counter, bounded spin).
// Attention work is cut into items of a fixed number of 32-token units per (stream, kv head) — a function of that
// stream's KV length only — so the split-KV partials, and therefore the logits, are identical whether a stream is
// decoded alone or in a batch.
// Every wait is bounded: a lost signal sets an error flag (reported through the stream scalars) instead of hanging.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "../../include/livecc_b200.h"
#include "common.cuh"
#include "gemm.h"
#include "launch.h"
#include "mega.h"
#include "mma.cuh"

namespace lcc {

namespace {

constexpr int NCW = 8;                  // consumer warps
constexpr int NPW = 2;                  // producer warps (power of two; ring groups are a multiple of it)
constexpr int GT = 4;                   // tiles per ring group: one TMA op moves GT tiles (16 KB)
constexpr int GROUP = 4096 * GT;
constexpr int NCT = NCW * 32;           // consumer threads
constexpr int TILE = 4096;              // bytes per ring slot
constexpr int QPITCH = 136;             // bf16 elements per staged q row
constexpr int SCRATCH = 45056;
// scratch carve-up (bytes)
constexpr int SC_RED = 0;               // GEMV: float [2][8][32][8]                       (16384)
constexpr int SC_PART = 16384;          // RMSNorm slice sums float [8 streams][8 slices]   (256)
constexpr int SC_QS = 0;                // attention: bf16 [16][136]                       (4352)
constexpr int SC_KNEW = 4352;           // bf16 [128]
constexpr int SC_VNEW = 4608;           // bf16 [128]
constexpr int SC_SML = 4864;            // float [9][8][2]                                 (576)
constexpr int SC_SO = 5632;             // float [9][8][128]                               (36864) -> 42496
constexpr int SC_FLAG = 42496;          // int [4]
constexpr int SC_FAC = 42512;           // float [MG_MAX_ITEMS * 8] merge weights (2048 B) -> 44560

__device__ __forceinline__ uint32_t ld_cg_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_cg_bf16(const bf16* p) {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return bf2f(__ushort_as_bfloat16(v));
}
__device__ __forceinline__ uint4 ld_cg_u128(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr unsigned long long kWaitLimitNs = 2000000000ull;  // 2 s: far beyond any legitimate wait inside one step

// Bounded mbarrier wait: a lost signal flags the error (sticky, global) and lets the kernel run to completion with
// garbage instead of hanging the GPU; once the flag is up every later wait gives up immediately.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err, int code) {
    const uint32_t a = smem_u32(bar);
    unsigned long long t0 = 0;
    for (int it = 0;; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
        if ((it & 0xfff) != 0xfff) continue;   // keep the spin tight: the abort checks cost a global round trip
        if (ld_volatile_i32(err)) return false;
        if (t0 == 0) t0 = globaltimer_ns();
        else if (globaltimer_ns() - t0 > kWaitLimitNs) { atomicExch(err, code); return false; }
    }
}

struct Shared {
    uint8_t* ring;      // ngroup x GROUP, 1024-aligned
    bf16* xs;           // [bpad][xpitch]
    uint8_t* scratch;   // SCRATCH bytes
    uint64_t* full;     // [nslot]
    uint64_t* empty;    // [nslot]
    int* s_T;           // [8] old tokens in the cache per stream
    int* s_pos;         // [8] rope position of the new token
    int* s_active;      // [8]
    unsigned* released; // [ngroup] tiles of each group slot released so far (see wait_tile)
    unsigned* prog;     // [NPW] groups walked so far by each TMA producer (paces the L2 prefetcher)
    float2* rope;       // [8 streams][64] (cos, sin) of this step's position, already rounded to bf16 (mq2vl.py:201)
};

// ------------------------------------------------------------------------------------------------------------
// work enumeration shared by the producer and the consumers
// ------------------------------------------------------------------------------------------------------------
struct PairPlan {
    int units;   // 32-token units of old tokens (>= 1)
    int cu;      // units per item
    int nitems;
};
__device__ __forceinline__ PairPlan plan_pair(int T) {
    PairPlan pp;
    pp.units = max(1, (T + 31) >> 5);
    pp.cu = 8;
    while ((pp.units + pp.cu - 1) / pp.cu > MG_MAX_ITEMS) pp.cu <<= 1;
    pp.nitems = (pp.units + pp.cu - 1) / pp.cu;
    return pp;
}

// ------------------------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------------------------
// NPW producer warps (one lane each) walk the same GROUP sequence (a group = GT = 4 consecutive tiles = 16 KB = one TMA
// op for weights, two for a K/V unit); producer `which` issues the groups with group % NPW == which. Measured on B200:
// with one 4 KB box per op the TMA path delivered ~40 GB/s per SM whatever the ring depth (per-op cost, not latency) and
// one producer thread could not even issue that; 16 KB ops cut the op count by four.
struct Ring {
    uint8_t* base;
    uint64_t *full, *empty;
    int ngroup;
    unsigned group;  // running group index of this CTA (all producers count every group)
    int* err;
    unsigned which;  // this producer's residue
    unsigned slot;   // ring slot (group granularity) and phase of this producer's NEXT group
    unsigned phase;
    volatile unsigned* prog;  // [NPW] progress counters in shared memory
    unsigned lookahead;       // prefetcher: groups it may run ahead of the slowest TMA producer
};

__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// PF = false (TMA producers): true if this producer owns the group; its slot is then armed.
// PF = true (the L2 prefetcher, one extra thread walking the same sequence): every group is its own; it only waits until
// it is at most `lookahead` groups ahead of the slowest TMA producer. The ring (160 KB/SM) is exactly what the TMA path
// needs in flight for full bandwidth, so any stall of the consumers (grid barrier, RMSNorm staging, split-K reduction:
// ~4 us per phase, measured) used to stop HBM. The prefetcher keeps pulling the stream into L2 (lookahead x 16 KB per SM,
// ~38 MB chip-wide) through those bubbles, and the ring then refills from L2.
template <bool PF>
__device__ __forceinline__ bool producer_begin(Ring& r) {
    if (PF) {
        unsigned long long t0 = 0;
        for (int it = 0;; ++it) {
            unsigned m = r.prog[0];
#pragma unroll
            for (int j = 1; j < NPW; ++j) m = min(m, r.prog[j]);
            if (r.group < m + r.lookahead) break;
            if ((it & 0xfff) != 0xfff) continue;
            if (ld_volatile_i32(r.err)) break;
            if (t0 == 0) t0 = globaltimer_ns();
            else if (globaltimer_ns() - t0 > kWaitLimitNs) break;   // pacing only: giving up is harmless
        }
        ++r.group;
        return true;
    }
    const bool mine = (r.group & (NPW - 1)) == r.which;
    ++r.group;
    r.prog[r.which] = r.group;
    if (!mine) return false;
    mbar_wait_bounded(&r.empty[r.slot], r.phase ^ 1, r.err, 2);
    mbar_arrive_expect_tx(&r.full[r.slot], GROUP);
    return true;
}
__device__ __forceinline__ void producer_end(Ring& r) {
    r.slot += NPW;
    if (r.slot >= (unsigned)r.ngroup) { r.slot -= r.ngroup; r.phase ^= 1; }
}

// weights: 3-D view [k chunk][row][64] of the [N][K] matrix (mega_make_weight_tmap), box = 4 chunks x 32 rows x 64
template <bool PF>
__device__ void producer_gemv(Ring& r, const CUtensorMap* tm, int N, int K, unsigned& rr, int cta, int G) {
    const int RB = N >> 5, KG = K >> 8;  // groups of 4 k-chunks
    const int first = (int)(((unsigned)cta + (unsigned)G - rr % (unsigned)G) % (unsigned)G);
    rr += RB;
    for (int b = first; b < RB; b += G)
        for (int kg = 0; kg < KG; ++kg)
            if (producer_begin<PF>(r)) {
                if (PF) {
                    tma_prefetch_l2_3d(tm, 0, b * 32, kg * GT);
                } else {
                    tma_load_3d(r.base + (size_t)r.slot * GROUP, tm, &r.full[r.slot], 0, b * 32, kg * GT);
                    producer_end(r);
                }
            }
}

// K/V: 3-D view [half][row][64] of a pool (128 dims = 2 halves), box = 2 halves x 32 rows x 64 = 8 KB; a unit = K op + V op
template <bool PF>
__device__ void producer_attn(Ring& r, const MegaParams& p, const Shared& sh, const CUtensorMap* tk, const CUtensorMap* tv,
                              int layer, unsigned& rr, int cta, int G) {
    unsigned gi = 0;
    for (int b = 0; b < p.B; ++b) {
        if (!sh.s_active[b]) continue;
        const PairPlan pp = plan_pair(sh.s_T[b]);
        const int* pt = p.st[b].page_table;
        for (int g = 0; g < p.Hkv; ++g, gi += pp.nitems) {
            // item `it` of this pair has global index gi + it and belongs to CTA (rr + gi + it) mod G
            const int it0 = (int)(((unsigned)cta + (unsigned)G - (rr + gi) % (unsigned)G) % (unsigned)G);
            for (int it = it0; it < pp.nitems; it += G) {
                const int u1 = min(pp.units, (it + 1) * pp.cu);
                for (int u = it * pp.cu; u < u1; ++u) {
                    if (!producer_begin<PF>(r)) continue;
                    const int t0 = u << 5;
                    const int page = pt[t0 >> 6];
                    const int row = layer * p.kv_rows_per_layer + (page * p.Hkv + g) * 64 + (t0 & 32);
                    if (PF) {
                        tma_prefetch_l2_3d(tk, 0, row, 0);
                        tma_prefetch_l2_3d(tv, 0, row, 0);
                    } else {
                        uint8_t* dst = r.base + (size_t)r.slot * GROUP;
                        tma_load_3d(dst, tk, &r.full[r.slot], 0, row, 0);
                        tma_load_3d(dst + 2 * TILE, tv, &r.full[r.slot], 0, row, 0);
                        producer_end(r);
                    }
                }
            }
        }
    }
    rr += gi;
}

// ------------------------------------------------------------------------------------------------------------
// consumers
// ------------------------------------------------------------------------------------------------------------
struct Cons {
    uint8_t* ring;
    uint64_t *full, *empty;
    unsigned* released;
    int ngroup;
    unsigned tile;  // tile index of the first tile of the current phase for this CTA (a multiple of GT)
    int* err;
    int warp, lane, tid;
    unsigned long long* trace;  // this CTA's 64 stamps or null
    int tr_n;
};
__device__ __forceinline__ void trace_stamp(Cons& c) {
    if (c.trace && c.tid == 0 && c.tr_n < 64) c.trace[c.tr_n++] = globaltimer_ns();
}

// The GT tiles of a ring group are in general consumed by DIFFERENT warps, and so are consecutive rounds of a group. A
// parity wait alone is then ambiguous: a warp that asks for round r while the group's barrier is still in round r-1
// sees "parity differs" and would read its tile a round early. Each group therefore counts its released tiles; the
// consumer of a tile of round r first waits until GT*r tiles were released (then the barrier is provably in round r)
// and only then does the parity wait.
__device__ __forceinline__ uint32_t wait_tile(const Cons& c, unsigned t) {
    const unsigned grp = t / GT, slot = grp % c.ngroup, round = grp / c.ngroup;
    const volatile unsigned* rel = c.released + slot;
    if (*rel < GT * round) {
        unsigned long long t0 = 0;
        for (int it = 0; *rel < GT * round; ++it) {
            if ((it & 0xfff) != 0xfff) continue;
            if (ld_volatile_i32(c.err)) break;
            if (t0 == 0) t0 = globaltimer_ns();
            else if (globaltimer_ns() - t0 > kWaitLimitNs) { atomicExch(c.err, 4); break; }
        }
    }
    mbar_wait_bounded(&c.full[slot], round & 1, c.err, 3);
    return smem_u32(c.ring + (size_t)slot * GROUP + (size_t)(t % GT) * TILE);
}
__device__ __forceinline__ void release_tile(const Cons& c, unsigned t) {
    __syncwarp();
    if (c.lane == 0) {
        const unsigned slot = (t / GT) % c.ngroup;
        atomicAdd(c.released + slot, 1u);
        mbar_arrive(&c.empty[slot]);   // GT arrivals free the group; release.cta orders the count before any refill
    }
}

// Grid-wide barrier among the consumer halves of all CTAs (the producers never wait here).
__device__ __forceinline__ void grid_sync(const MegaParams& p, const Cons& c, unsigned& epoch, int G) {
    consumer_sync();
    ++epoch;
    if (c.tid == 0) {
        // release at gpu scope: cumulative over everything the other consumer threads wrote before the bar.sync above
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar) : "memory");
        const unsigned target = epoch * (unsigned)G;
        unsigned long long t0 = 0;
        for (int it = 0; ld_acquire_u32(p.bar) < target; ++it) {
            if ((it & 0x3ff) != 0x3ff) continue;
            if (ld_volatile_i32(p.err)) break;
            if (t0 == 0) t0 = globaltimer_ns();
            else if (globaltimer_ns() - t0 > kWaitLimitNs) { atomicExch(p.err, 1); break; }
        }
    }
    consumer_sync();
}

// xs[s][k] = (norm_w ? norm_w[k] * bf16(x[s][k] * rsqrt(mean(x^2) + eps)) : x[s][k]) for the B live streams.
// The K range is cut into NCW fixed slices; warp w owns slice w of EVERY stream, and a stream's sum of squares is the
// sum of its 8 slice sums in slice order — the same arithmetic whether the stream is alone or one of eight.
__device__ void stage_x(const MegaParams& p, Cons& c, const Shared& sh, const bf16* src, int K, const bf16* norm_w,
                        int xpitch) {
    const int chunks = K >> 3;
    const int c0 = (int)((long long)chunks * c.warp / NCW), c1 = (int)((long long)chunks * (c.warp + 1) / NCW);
    float* part = reinterpret_cast<float*>(sh.scratch + SC_PART);  // [MG_MAXB][NCW]
    if (norm_w) {
        // the norm weights of this lane's chunks (static data): in flight together with the activation loads below
        uint4 nw[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ch = c0 + c.lane + 32 * j;
            nw[j] = ch < c1 ? *reinterpret_cast<const uint4*>(norm_w + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
        // pass 1: global (L2) -> xs raw + this slice's sum of squares; pass 2: normalise in shared memory
        for (int s = 0; s < p.B; ++s) {
            const bf16* row = src + (size_t)s * K;
            float sq = 0.f;
            for (int ch = c0 + c.lane; ch < c1; ch += 32) {
                const uint4 u = ld_cg_u128(row + ch * 8);
                *reinterpret_cast<uint4*>(sh.xs + (size_t)s * xpitch + ch * 8) = u;
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(uw[j]); sq += f.x * f.x + f.y * f.y; }
            }
            sq = warp_sum(sq);
            if (c.lane == 0) part[s * NCW + c.warp] = sq;
        }
        if (p.phase_mask & 64) trace_stamp(c);   // deep trace: activation loads done (thread 0's view)
        consumer_sync();
        if (p.phase_mask & 64) trace_stamp(c);   // deep trace: all warps' slices summed
        for (int s = 0; s < p.B; ++s) {
            float tot = 0.f;
#pragma unroll
            for (int j = 0; j < NCW; ++j) tot += part[s * NCW + j];
            const float rs = rsqrtf(tot / (float)K + p.eps);
            int jn = 0;
            for (int ch = c0 + c.lane; ch < c1; ch += 32, ++jn) {
                uint4* xp = reinterpret_cast<uint4*>(sh.xs + (size_t)s * xpitch + ch * 8);
                const uint4 u = *xp;
                const uint4 wv = jn < 2 ? nw[jn] : *reinterpret_cast<const uint4*>(norm_w + ch * 8);
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
                uint32_t ov[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(uw[j]), g = unpack_bf16x2(ww[j]);
                    ov[j] = pack_bf16x2(g.x * rbf(f.x * rs), g.y * rbf(f.y * rs));
                }
                *xp = make_uint4(ov[0], ov[1], ov[2], ov[3]);
            }
        }
    } else {
        for (int s = 0; s < p.B; ++s)
            for (int ch = c0 + c.lane; ch < c1; ch += 32)
                *reinterpret_cast<uint4*>(sh.xs + (size_t)s * xpitch + ch * 8) = ld_cg_u128(src + (size_t)s * K + ch * 8);
    }
    consumer_sync();
}

enum { EP_BIAS = 0, EP_RESIDUAL = 1, EP_SWIGLU = 2, EP_LOGITS = 3 };

struct GemvOut {
    const bf16* bias;   // EP_BIAS [N]
    bf16* out;          // EP_BIAS: [8][N]; EP_RESIDUAL: residual stream [8][N] in/out; EP_SWIGLU: [8][N/2]
    float *lg_raw, *lg_proc;  // EP_LOGITS [8][N]
};

// One GEMV phase over this CTA's row blocks. XG: activations read from global memory ([8][K], K too large for smem).
template <int EP, bool XG>
__device__ void consumer_gemv(const MegaParams& p, Cons& c, const Shared& sh, int N, int K, const bf16* xg, int xpitch,
                              const GemvOut& o, unsigned& rr, int cta, int G) {
    const int RB = N >> 5, KC = K >> 6;
    const int first = (int)(((unsigned)cta + (unsigned)G - rr % (unsigned)G) % (unsigned)G);
    rr += RB;
    float* red = reinterpret_cast<float*>(sh.scratch + SC_RED);
    const int g = c.lane >> 2, t = c.lane & 3;
    int biter = 0;
    for (int b = first; b < RB; b += G, ++biter) {
        float acc[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
        const unsigned tbase = c.tile + (unsigned)biter * KC;
        for (int kc = c.warp; kc < KC; kc += NCW) {
            uint32_t bx[4][2];
            if (XG) {  // issue the activation loads before waiting for the weight tile
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const bf16* xp = xg + (size_t)g * K + kc * 64 + kk * 16 + 2 * t;
                    bx[kk][0] = g < p.B ? ld_cg_u32(xp) : 0u;
                    bx[kk][1] = g < p.B ? ld_cg_u32(xp + 8) : 0u;
                }
            }
            const uint32_t tb = wait_tile(c, tbase + kc);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (!XG) {  // xs holds bpad rows only: columns >= B of the MMA are zero and never read back
                    const bf16* xp = sh.xs + (size_t)g * xpitch + kc * 64 + kk * 16 + 2 * t;
                    bx[kk][0] = g < p.B ? *reinterpret_cast<const uint32_t*>(xp) : 0u;
                    bx[kk][1] = g < p.B ? *reinterpret_cast<const uint32_t*>(xp + 8) : 0u;
                }
                const int r = (c.lane & 7) + ((c.lane >> 3) & 1) * 8;
                const int ch = 2 * kk + (c.lane >> 4);
                const uint32_t a0 = tb + r * 128 + ((ch ^ (r & 7)) << 4);
                uint32_t af[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3]) : "r"(a0));
                mma_bf16_16816(acc[0], af, bx[kk][0], bx[kk][1]);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3]) : "r"(a0 + 16 * 128));
                mma_bf16_16816(acc[1], af, bx[kk][0], bx[kk][1]);
            }
            release_tile(c, tbase + kc);
        }
        // cross-warp K reduction, fixed order
        float* rb = red + (biter & 1) * (NCW * 32 * 8);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float* d0 = rb + ((c.warp * 32 + m * 16 + g) * 8) + 2 * t;
            d0[0] = acc[m][0]; d0[1] = acc[m][1];
            d0[8 * 8] = acc[m][2]; d0[8 * 8 + 1] = acc[m][3];
        }
        consumer_sync();
        const int row = c.tid >> 3, col = c.tid & 7;
        if (EP == EP_SWIGLU) {
            if (c.tid < 128 && col < p.B) {
                float gt = 0.f, up = 0.f;
#pragma unroll
                for (int w = 0; w < NCW; ++w) {
                    gt += rb[(w * 32 + row) * 8 + col];
                    up += rb[(w * 32 + 16 + row) * 8 + col];
                }
                gt = rbf(gt); up = rbf(up);
                const float sl = rbf(gt / (1.0f + expf(-gt)));   // same rounding points as gemv.cu / mq2vl.py:503
                o.out[(size_t)col * (N >> 1) + b * 16 + row] = f2bf(sl * up);
            }
        } else if (col < p.B) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NCW; ++w) v += rb[(w * 32 + row) * 8 + col];
            const int n = b * 32 + row;
            if (EP == EP_BIAS) o.out[(size_t)col * N + n] = f2bf(v + bf2f(o.bias[n]));
            else if (EP == EP_RESIDUAL) {
                bf16* hp = o.out + (size_t)col * N + n;
                *hp = f2bf(rbf(v) + ld_cg_bf16(hp));   // o_proj / down_proj + residual (mq2vl.py:645,660)
            } else {
                const float l = rbf(v);
                o.lg_raw[(size_t)col * N + n] = l;
                o.lg_proc[(size_t)col * N + n] = l;
            }
        }
    }
    c.tile += (unsigned)biter * KC;
}

// Rotates 8 consecutive (j, j+64) pairs of one head: two 16-byte loads from L2 (the qkv vector was written by other CTAs
// in this launch), torch bf16 semantics (each product and the sum rounded to bf16, as attention.cu::rope1d_row).
__device__ __forceinline__ void rope_chunk(const bf16* src, bf16* dst, int j0, const float2* tab) {
    const uint4 lo = ld_cg_u128(src + j0), hi = ld_cg_u128(src + j0 + 64);
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x1 = unpack_bf16x2(lw[i]), x2 = unpack_bf16x2(hw[i]);
        const float2 c0 = tab[j0 + 2 * i], c1 = tab[j0 + 2 * i + 1];
        ol[i] = pack_bf16x2(rbf(rbf(x1.x * c0.x) + rbf(-x2.x * c0.y)), rbf(rbf(x1.y * c1.x) + rbf(-x2.y * c1.y)));
        oh[i] = pack_bf16x2(rbf(rbf(x2.x * c0.x) + rbf(x1.x * c0.y)), rbf(rbf(x2.y * c1.x) + rbf(x1.y * c1.y)));
    }
    *reinterpret_cast<uint4*>(dst + j0) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    *reinterpret_cast<uint4*>(dst + j0 + 64) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
}

// Attention phase: this CTA's items.
__device__ void consumer_attn(const MegaParams& p, Cons& c, const Shared& sh, int layer, unsigned& rr, int cta, int G) {
    constexpr int D = 128;
    const int Gq = p.Hq / p.Hkv;  // query heads per KV head (<= 8)
    bf16* qs = reinterpret_cast<bf16*>(sh.scratch + SC_QS);
    bf16* knew = reinterpret_cast<bf16*>(sh.scratch + SC_KNEW);
    bf16* vnew = reinterpret_cast<bf16*>(sh.scratch + SC_VNEW);
    float* sml = reinterpret_cast<float*>(sh.scratch + SC_SML);
    float* so = reinterpret_cast<float*>(sh.scratch + SC_SO);
    int* flag = reinterpret_cast<int*>(sh.scratch + SC_FLAG);
    const int g4 = c.lane >> 2, t4 = c.lane & 3;
    unsigned gi = 0, tiles_done = 0;
    for (int b = 0; b < p.B; ++b) {
        if (!sh.s_active[b]) continue;
        const int T = sh.s_T[b];
        const PairPlan pp = plan_pair(T);
        const bf16* qkv = p.qkv + (size_t)b * p.qkv_dim;
        for (int g = 0; g < p.Hkv; ++g, gi += pp.nitems) {
            const int it0 = (int)(((unsigned)cta + (unsigned)G - (rr + gi) % (unsigned)G) % (unsigned)G);
            for (int it = it0; it < pp.nitems; it += G) {
                const int u0 = it * pp.cu, u1 = min(pp.units, (it + 1) * pp.cu);
                const bool last_item = it == pp.nitems - 1;
                // ---- stage the rotated q heads of this KV group (rows >= Gq zero); the owner of the pair's last
                //      item also rotates + appends the new token's k and v ----
                const float2* tab = sh.rope + b * 64;
                if (c.tid < 128) {   // 16 rows x 8 chunks of 8 pairs
                    const int r = c.tid >> 3, j0 = (c.tid & 7) * 8;
                    if (r < Gq) rope_chunk(qkv + (size_t)(g * Gq + r) * D, qs + r * QPITCH, j0, tab);
                    else {
                        *reinterpret_cast<uint4*>(qs + r * QPITCH + j0) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(qs + r * QPITCH + j0 + 64) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                if (last_item) {
                    const int page = p.st[b].page_table[T >> 6], slot = T & 63;
                    const size_t base = (size_t)layer * p.layer_stride + (((size_t)page * p.Hkv + g) * 64 + slot) * D;
                    if (c.tid >= 128 && c.tid < 136) {
                        const int j0 = (c.tid - 128) * 8;
                        rope_chunk(qkv + (size_t)(p.Hq + g) * D, knew, j0, tab);
                        *reinterpret_cast<uint4*>(p.k_pool + base + j0) = *reinterpret_cast<const uint4*>(knew + j0);
                        *reinterpret_cast<uint4*>(p.k_pool + base + j0 + 64) = *reinterpret_cast<const uint4*>(knew + j0 + 64);
                    } else if (c.tid >= 136 && c.tid < 152) {
                        const int ch = c.tid - 136;
                        const uint4 v = ld_cg_u128(qkv + (size_t)(p.Hq + p.Hkv + g) * D + ch * 8);
                        *reinterpret_cast<uint4*>(vnew + ch * 8) = v;
                        *reinterpret_cast<uint4*>(p.v_pool + base + ch * 8) = v;
                    }
                    if (slot == 0) {  // first token of a (possibly recycled) page: later rows must be finite for P = 0
                        bf16* vt = p.v_pool + base + D;
                        for (int i = c.tid; i < 63 * 16; i += NCT) *reinterpret_cast<uint4*>(vt + i * 8) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                consumer_sync();
                if (p.phase_mask & 32) trace_stamp(c);   // deep trace: staged
                // ---- per warp: flash pass over its 32-token units ----
                uint32_t qf[D / 16][4];
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const int mat = c.lane >> 3, r = c.lane & 7;
                    ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3],
                                qs + (size_t)((mat & 1) * 8 + r) * QPITCH + kk * 16 + (mat >> 1) * 8);
                }
                float oacc[D / 8][4];
#pragma unroll
                for (int d = 0; d < D / 8; ++d)
#pragma unroll
                    for (int j = 0; j < 4; ++j) oacc[d][j] = 0.f;
                float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
                for (int u = u0 + c.warp; u < u1; u += NCW) {
                    const unsigned tb = c.tile + tiles_done + (unsigned)(u - u0) * 4;
                    const uint32_t k_lo = wait_tile(c, tb), k_hi = wait_tile(c, tb + 1);
                    const uint32_t v_lo = wait_tile(c, tb + 2), v_hi = wait_tile(c, tb + 3);
                    const int valid = min(32, T - (u << 5));  // tokens of this unit that are old cache entries
                    if (valid < 32) {  // rows >= valid may hold anything (recycled page / stale prefetch): P = 0 needs finite V
                        for (int i = c.lane; i < (32 - valid) * 16; i += 32) {
                            const int r = valid + (i >> 4), ch = i & 15;
                            const uint32_t a = ((ch & 8) ? v_hi : v_lo) + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
                            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
                        }
                        fence_proxy_async_smem();  // these generic-proxy stores must be ordered before the slot's next TMA fill
                        __syncwarp();
                    }
                    float s[4][4];
#pragma unroll
                    for (int n = 0; n < 4; ++n)
#pragma unroll
                        for (int j = 0; j < 4; ++j) s[n][j] = 0.f;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t kt = kk < 4 ? k_lo : k_hi;
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            // x4: (tok 0-7,k 0-7) (tok 0-7,k 8-15) (tok 8-15,k 0-7) (tok 8-15,k 8-15)
                            const int mat = c.lane >> 3, r = nt * 16 + (mat >> 1) * 8 + (c.lane & 7);
                            const int ch = 2 * (kk & 3) + (mat & 1);
                            uint32_t b0, b1, b2, b3;
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                                         : "r"(kt + r * 128 + ((ch ^ (r & 7)) << 4)));
                            mma_bf16_16816(s[nt * 2], qf[kk], b0, b1);
                            mma_bf16_16816(s[nt * 2 + 1], qf[kk], b2, b3);
                        }
                    }
                    if (valid < 32) {
#pragma unroll
                        for (int n = 0; n < 4; ++n) {
                            const int kidx = n * 8 + 2 * t4;
                            if (kidx >= valid) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
                            if (kidx + 1 >= valid) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
                        }
                    }
                    uint32_t pf[2][4];
                    softmax_step<D, 2>(s, p.scale_log2, m, l, oacc, pf);
#pragma unroll
                    for (int kt2 = 0; kt2 < 2; ++kt2)
#pragma unroll
                        for (int dn = 0; dn < D / 16; ++dn) {
                            // trans x4: (tok 0-7,d 0-7) (tok 8-15,d 0-7) (tok 0-7,d 8-15) (tok 8-15,d 8-15)
                            const int mat = c.lane >> 3, r = kt2 * 16 + (mat & 1) * 8 + (c.lane & 7);
                            const int ch = 2 * (dn & 3) + (mat >> 1);
                            const uint32_t vt = dn < 4 ? v_lo : v_hi;
                            uint32_t b0, b1, b2, b3;
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                                         : "r"(vt + r * 128 + ((ch ^ (r & 7)) << 4)));
                            mma_bf16_16816(oacc[dn * 2], pf[kt2], b0, b1);
                            mma_bf16_16816(oacc[dn * 2 + 1], pf[kt2], b2, b3);
                        }
                    release_tile(c, tb); release_tile(c, tb + 1); release_tile(c, tb + 2); release_tile(c, tb + 3);
                }
                tiles_done += (unsigned)(u1 - u0) * 4;
                l[0] += __shfl_xor_sync(0xffffffffu, l[0], 1);
                l[0] += __shfl_xor_sync(0xffffffffu, l[0], 2);
                // ---- merge the warps (+ the new token) of this item through shared memory; rows g4 < 8 only ----
                consumer_sync();  // everybody is done with qs
                if (p.phase_mask & 32) trace_stamp(c);   // deep trace: units done (SC_SO overlaps nothing, but SC_QS is re-staged next item)
                {
                    float* dst = so + ((size_t)c.warp * 8 + g4) * D + 2 * t4;
#pragma unroll
                    for (int d = 0; d < D / 8; ++d) { dst[d * 8] = oacc[d][0]; dst[d * 8 + 1] = oacc[d][1]; }
                    if (t4 == 0) { sml[(c.warp * 8 + g4) * 2] = m[0]; sml[(c.warp * 8 + g4) * 2 + 1] = l[0]; }
                }
                if (last_item) {  // 9th contributor: the new token itself (score from the staged q and the rotated k)
                    if (c.warp < Gq) {
                        float dot = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) dot += bf2f(qs[c.warp * QPITCH + c.lane * 4 + j]) * bf2f(knew[c.lane * 4 + j]);
                        dot = warp_sum(dot);
#pragma unroll
                        for (int j = 0; j < 4; ++j) so[((size_t)8 * 8 + c.warp) * D + c.lane * 4 + j] = bf2f(vnew[c.lane * 4 + j]);
                        if (c.lane == 0) { sml[(8 * 8 + c.warp) * 2] = dot; sml[(8 * 8 + c.warp) * 2 + 1] = 1.f; }
                    }
                }
                consumer_sync();
                const int nsrc = last_item ? 9 : 8;
                const size_t pbase = (((size_t)b * p.Hkv + g) * MG_MAX_ITEMS + it) * 8;
                float* fac = reinterpret_cast<float*>(sh.scratch + SC_FAC);   // [9][8] weights of the contributors
                if (c.tid < nsrc * 8) {
                    const int r = c.tid & 7;
                    float M = -INFINITY;
                    for (int w = 0; w < nsrc; ++w) M = fmaxf(M, sml[(w * 8 + r) * 2]);
                    const float mw = sml[c.tid * 2];
                    fac[c.tid] = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * p.scale_log2);
                    if (c.tid < 8 && r < Gq) p.part_ml[(pbase + r) * 2] = M;
                }
                consumer_sync();
                for (int i = c.tid; i < Gq * D; i += NCT) {
                    const int r = i / D, d = i % D;
                    float acc = 0.f;
                    for (int w = 0; w < nsrc; ++w) acc += fac[w * 8 + r] * so[((size_t)w * 8 + r) * D + d];
                    p.part_o[(pbase + r) * D + d] = acc;
                    if (d == 0) {
                        float L = 0.f;
                        for (int w = 0; w < nsrc; ++w) L += fac[w * 8 + r] * sml[(w * 8 + r) * 2 + 1];
                        p.part_ml[(pbase + r) * 2 + 1] = L;
                    }
                }
                consumer_sync();
                if (p.phase_mask & 32) trace_stamp(c);   // deep trace: partial written
                if (c.tid == 0) {
                    int prev;   // acq_rel at gpu scope: publishes this item's partials, and the last arriver sees all of them
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(prev) : "l"(&p.pair_cnt[b * p.Hkv + g]) : "memory");
                    flag[0] = (prev == pp.nitems - 1) ? 1 : 0;
                    if (flag[0]) p.pair_cnt[b * p.Hkv + g] = 0;  // re-armed for the next layer
                }
                consumer_sync();
                if (p.phase_mask & 32) trace_stamp(c);   // deep trace: counted
                if (flag[0]) {  // last item of this (stream, kv head) to finish: merge the items in index order
                    // (m, l) of every item first, in parallel, into shared memory (the accumulators `so` are free now);
                    // then each thread folds the items' float4 partials with the loads of several items in flight
                    const size_t pb = ((size_t)b * p.Hkv + g) * MG_MAX_ITEMS * 8;
                    float* sm_m = so;                       // [MG_MAX_ITEMS][8]
                    float* sm_l = so + MG_MAX_ITEMS * 8;    // [MG_MAX_ITEMS][8]
                    for (int i = c.tid; i < pp.nitems * 8; i += NCT) {
                        const float2 ml = __ldcg(reinterpret_cast<const float2*>(p.part_ml + (pb + i) * 2));
                        sm_m[i] = ml.x;
                        sm_l[i] = ml.y;
                    }
                    consumer_sync();
                    float* wq = reinterpret_cast<float*>(sh.scratch + SC_FAC);   // [nitems][8] weight of item q for row r
                    for (int i = c.tid; i < pp.nitems * 8; i += NCT) {
                        const int r = i & 7;
                        float M = -INFINITY;
                        for (int q = 0; q < pp.nitems; ++q) M = fmaxf(M, sm_m[q * 8 + r]);
                        const float mq = sm_m[i];
                        wq[i] = (mq == -INFINITY) ? 0.f : exp2f((mq - M) * p.scale_log2);
                    }
                    consumer_sync();
                    for (int i = c.tid; i < Gq * (D / 4); i += NCT) {
                        const int r = i / (D / 4), d4 = (i % (D / 4)) * 4;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        float L = 0.f;
                        const float* src = p.part_o + (pb + r) * D + d4;
#pragma unroll 8
                        for (int q = 0; q < pp.nitems; ++q) {
                            const float f = wq[q * 8 + r];
                            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)q * 8 * D));
                            acc.x += f * v.x; acc.y += f * v.y; acc.z += f * v.z; acc.w += f * v.w;
                            L += f * sm_l[q * 8 + r];
                        }
                        const float inv = L > 0.f ? 1.f / L : 0.f;
                        uint2 o;
                        o.x = pack_bf16x2(acc.x * inv, acc.y * inv);
                        o.y = pack_bf16x2(acc.z * inv, acc.w * inv);
                        *reinterpret_cast<uint2*>(p.attn + (size_t)b * p.Hq * D + (size_t)(g * Gq + r) * D + d4) = o;
                    }
                }
                consumer_sync();
            }
        }
    }
    rr += gi;
    c.tile += tiles_done;
}

}  // namespace

__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                   const MegaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int bpad = p.B <= 1 ? 1 : (p.B <= 2 ? 2 : (p.B <= 4 ? 4 : 8));
    const int xpitch = p.H + 8;
    Shared sh;
    sh.ring = sm;
    sh.xs = reinterpret_cast<bf16*>(sm + (size_t)p.ngroup * GROUP);
    sh.scratch = reinterpret_cast<uint8_t*>(sh.xs) + (((size_t)bpad * xpitch * 2 + 127) & ~size_t(127));
    sh.full = reinterpret_cast<uint64_t*>(sh.scratch + SCRATCH);
    sh.empty = sh.full + MG_MAX_SLOTS;
    sh.s_T = reinterpret_cast<int*>(sh.empty + MG_MAX_SLOTS);
    sh.s_pos = sh.s_T + 8;
    sh.s_active = sh.s_pos + 8;
    sh.released = reinterpret_cast<unsigned*>(sh.s_active + 8);
    sh.prog = sh.released + MG_MAX_SLOTS;
    sh.rope = reinterpret_cast<float2*>(sh.prog + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.ngroup; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], GT); sh.released[i] = 0u; }
        for (int i = 0; i < NPW; ++i) sh.prog[i] = 0u;
        fence_barrier_init();
    }
    if (threadIdx.x < 8) {
        const int b = threadIdx.x;
        int T = 0, pos = 0, act = 0;
        if (b < p.B) {
            const int* sc = p.st[b].scalars;
            T = sc[LCC_SC_KV_LEN]; pos = sc[LCC_SC_ROPE_POS]; act = sc[LCC_SC_FINISHED] ? 0 : 1;
        }
        sh.s_T[b] = T; sh.s_pos[b] = pos; sh.s_active[b] = act;
    }
    __syncthreads();
    int any = 0;
    for (int b = 0; b < p.B; ++b) any |= sh.s_active[b];
    if (!any) return;  // every stream is finished: the step is a no-op (graph replay after EOS)
    // The position of the new token is fixed for the whole step: its 64 (cos, sin) pairs per stream are computed once
    // (fp32 angle, cos/sin rounded to bf16 as the reference does) instead of by every attention item of every layer.
    for (int i = threadIdx.x; i < p.B * 64; i += blockDim.x) {
        const float ang = __fmul_rn((float)sh.s_pos[i >> 6], p.inv_freq[i & 63]);
        sh.rope[i] = make_float2(rbf(cosf(ang)), rbf(sinf(ang)));
    }
    __syncthreads();

    const bool ph_qkv = p.phase_mask & 1, ph_attn = p.phase_mask & 2, ph_o = p.phase_mask & 4, ph_gu = p.phase_mask & 8,
               ph_down = p.phase_mask & 16;
    unsigned rr = 0;
    if (warp >= NCW) {
        // ================================= producers =================================
        if (lane == 0 && (warp < NCW + NPW || p.lookahead > 0)) {
            const bool pf = warp == NCW + NPW;   // the last warp is the L2 prefetcher
            Ring r{sh.ring, sh.full, sh.empty, p.ngroup, 0u, p.err, (unsigned)(warp - NCW), (unsigned)(warp - NCW), 0u, sh.prog,
                   (unsigned)(p.ngroup + p.lookahead)};
            auto walk = [&](auto tag) {
                constexpr bool PF = decltype(tag)::value;
                for (int l = p.layer_begin; l < p.layer_end; ++l) {
                    const CUtensorMap* wm = p.wmaps + 4 * l;
                    if (ph_qkv) producer_gemv<PF>(r, wm + 0, p.qkv_dim, p.H, rr, cta, G);
                    if (ph_attn) producer_attn<PF>(r, p, sh, &tmap_k, &tmap_v, l, rr, cta, G);
                    if (ph_o) producer_gemv<PF>(r, wm + 1, p.H, p.Hq * 128, rr, cta, G);
                    if (ph_gu) producer_gemv<PF>(r, wm + 2, 2 * p.I, p.H, rr, cta, G);
                    if (ph_down) producer_gemv<PF>(r, wm + 3, p.H, p.I, rr, cta, G);
                }
                if (p.do_head) producer_gemv<PF>(r, p.wmaps + 4 * p.L, p.V, p.H, rr, cta, G);
            };
            if (pf) walk(std::true_type{});
            else {
                walk(std::false_type{});
                sh.prog[warp - NCW] = 0x7fffffffu;   // done: never hold the prefetcher back
            }
        }
        return;
    }
    // ================================= consumers =================================
    Cons c{sh.ring, sh.full, sh.empty, sh.released, p.ngroup, 0u, p.err, warp, lane, (int)threadIdx.x,
           p.trace ? p.trace + (size_t)cta * 64 : nullptr, 0};
    unsigned epoch = 0;
    // optional timeline (LIVECC_B200_MEGA_TRACE=1): globaltimer stamps of consumer thread 0, 64 per CTA
#define MG_TRACE() trace_stamp(c)
    MG_TRACE();
    for (int l = p.layer_begin; l < p.layer_end; ++l) {
        const MegaLayer& ly = p.layers[l];
        if (ph_qkv) {
            stage_x(p, c, sh, p.h, p.H, ly.ln1_w, xpitch);
            MG_TRACE();
            GemvOut o{ly.qkv_b, p.qkv, nullptr, nullptr};
            consumer_gemv<EP_BIAS, false>(p, c, sh, p.qkv_dim, p.H, nullptr, xpitch, o, rr, cta, G);
            MG_TRACE();
            grid_sync(p, c, epoch, G);
            MG_TRACE();
        }
        if (ph_attn) {
            consumer_attn(p, c, sh, l, rr, cta, G);
            MG_TRACE();
            grid_sync(p, c, epoch, G);
            MG_TRACE();
        }
        if (ph_o) {
            stage_x(p, c, sh, p.attn, p.Hq * 128, nullptr, xpitch);
            MG_TRACE();
            GemvOut o{nullptr, p.h, nullptr, nullptr};
            consumer_gemv<EP_RESIDUAL, false>(p, c, sh, p.H, p.Hq * 128, nullptr, xpitch, o, rr, cta, G);
            MG_TRACE();
            grid_sync(p, c, epoch, G);
            MG_TRACE();
        }
        if (ph_gu) {
            stage_x(p, c, sh, p.h, p.H, ly.ln2_w, xpitch);
            MG_TRACE();
            GemvOut o{nullptr, p.act, nullptr, nullptr};
            consumer_gemv<EP_SWIGLU, false>(p, c, sh, 2 * p.I, p.H, nullptr, xpitch, o, rr, cta, G);
            MG_TRACE();
            grid_sync(p, c, epoch, G);
            MG_TRACE();
        }
        if (ph_down) {
            GemvOut o{nullptr, p.h, nullptr, nullptr};
            consumer_gemv<EP_RESIDUAL, true>(p, c, sh, p.H, p.I, p.act, xpitch, o, rr, cta, G);
            MG_TRACE();
            grid_sync(p, c, epoch, G);
            MG_TRACE();
        }
    }
    if (p.do_head) {
        stage_x(p, c, sh, p.h, p.H, p.final_norm_w, xpitch);
        GemvOut o{nullptr, nullptr, p.logits_raw, p.logits_proc};
        consumer_gemv<EP_LOGITS, false>(p, c, sh, p.V, p.H, nullptr, xpitch, o, rr, cta, G);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
int mega_smem_bytes(int H, int B, int* ngroup_out) {
    const int bpad = B <= 1 ? 1 : (B <= 2 ? 2 : (B <= 4 ? 4 : 8));
    const int xs = ((bpad * (H + 8) * 2) + 127) & ~127;
    const int fixed = 1024 /*align*/ + xs + SCRATCH + MG_MAX_SLOTS * 16 + 384 + 4096;  // barriers, scalars, released[], prog[], rope table
    int ngroup = (232448 - fixed) / GROUP;
    if (ngroup > MG_MAX_SLOTS) ngroup = MG_MAX_SLOTS;
    // A multiple of the producer count: group slot s is then always armed by producer s % NPW, so two rounds of one slot
    // are ordered inside one thread (a producer a full ring round ahead of its neighbour would otherwise pass the parity
    // wait of a slot whose previous round has not even been issued -> double arm -> launch failure).
    ngroup &= ~(NPW - 1);
    *ngroup_out = ngroup;
    return fixed + ngroup * GROUP;
}

// [N][K] bf16 weights as a 3-D tensor {64 cols, N rows, K/64 chunks}: one box = 4 consecutive k-chunks of 32 rows, landing
// in shared memory as 4 consecutive 4 KB tiles [32 rows][128 B], each in the SWIZZLE_128B layout ldmatrix expects.
int mega_make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K) {
    if (K % 256) return -1;
    return make_tmap_bf16_3d_sw128(tm, w, 64, N, K / 64, K, 64, 32, GT);
}

int decode_mega_launch(const MegaParams& p_in, const void* k_pool, const void* v_pool, long long pool_rows, int num_sms,
                       cudaStream_t s) {
    MegaParams p = p_in;
    if (p.B < 1 || p.B > MG_MAXB || p.Hq % p.Hkv || p.Hq / p.Hkv > 8) return -1;
    if ((p.qkv_dim % 32) || (p.H % 256) || (p.I % 256) || ((2 * p.I) % 32) || (p.V % 32)) return -2;
    int ngroup = 0;
    const int smem = mega_smem_bytes(p.H, p.B, &ngroup);
    if (ngroup < 4) return -3;
    p.ngroup = ngroup;
    // pools viewed as {64 cols, rows, 2 halves}: one box = both 64-dim halves of 32 tokens -> tiles (lo, hi)
    CUtensorMap tk, tv;
    if (make_tmap_bf16_3d_sw128(&tk, k_pool, 64, pool_rows, 2, 128, 64, 32, 2)) return -10;
    if (make_tmap_bf16_3d_sw128(&tv, v_pool, 64, pool_rows, 2, 128, 64, 32, 2)) return -11;
    static SmemAttrOnce once;
    if (ensure_dyn_smem(once, decode_mega_kernel, 232448)) return -12;  // the opt-in maximum: the request varies with B
    count_launch();
    decode_mega_kernel<<<num_sms, MG_THREADS, smem, s>>>(tk, tv, p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[livecc_b200] decode_mega_kernel launch failed: %s (B=%d, smem=%d, ngroup=%d)\n", cudaGetErrorString(e),
                p.B, smem, ngroup);
        return -13;
    }
    return 0;
}

}  // namespace lcc
