// Token selection for one stream, fused into one single-CTA kernel per step (the reference spends ~10
// tiny launches and a host sync per token here: gen/utils.py:2762-2805):
//   RepetitionPenaltyLogitsProcessor (logits_process.py:407-410) over the whole id history,
//   optional ThresholdLogitsProcessor (REF/demo/infer.py:10-23),
//   argmax (greedy; gen/utils.py:2793), EOS / max_new_tokens stop, append to the sequence buffer,
//   advance of the device-side stream scalars, and the embedding row of the chosen token for the
//   next decode step (mq2vl.py:1256).
// Everything the next step needs stays on the device, so a whole generate() call runs without a host sync.
#include "../../include/livecc_b200.h"
#include "common.cuh"
#include "launch.h"
#include "ops.h"

namespace lcc {

__device__ __forceinline__ void sample_greedy_body(const SampleArgs& a) {
    int* sc = a.scalars;
    if (sc[LCC_SC_FINISHED]) return;
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    __shared__ float bcast[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hist = sc[LCC_SC_SEQ_LEN];

    // 1. repetition penalty: gather from raw, scatter to proc (idempotent under duplicate ids)
    if (a.repetition_penalty != 1.0f) {
        // torch's CUDA `scores / penalty` multiplies by fp32(1.0 / (double)penalty); the host passes that value
        const float inv_pen = a.inv_repetition_penalty != 0.f ? a.inv_repetition_penalty : 1.0f / a.repetition_penalty;
#pragma unroll 4
        for (int i = tid; i < hist; i += blockDim.x) {
            const int64_t id = a.seq[i];
            if (id >= 0 && id < a.V) {
                const float v = a.logits_raw[id];
                a.logits_proc[id] = v < 0.f ? v * a.repetition_penalty : v * inv_pen;
            }
        }
        __syncthreads();
    }

    // 2. optional streaming-EOS threshold: if softmax(scores)[tok] <= thr then scores[tok] = -inf
    if (a.thr_token >= 0 && a.thr_token < a.V) {
        float mx = -INFINITY;
        for (int i = tid; i < a.V; i += blockDim.x) mx = fmaxf(mx, a.logits_proc[i]);
        mx = warp_max(mx);
        if (lane == 0) red_v[warp] = mx;
        __syncthreads();
        if (warp == 0) {
            float t = red_v[lane];
            t = warp_max(t);
            if (lane == 0) bcast[0] = t;
        }
        __syncthreads();
        mx = bcast[0];
        float se = 0.f;
        for (int i = tid; i < a.V; i += blockDim.x) se += expf(a.logits_proc[i] - mx);
        se = warp_sum(se);
        __syncthreads();
        if (lane == 0) red_v[warp] = se;
        __syncthreads();
        if (warp == 0) {
            float t = red_v[lane];
            t = warp_sum(t);
            if (lane == 0) bcast[1] = t;
        }
        __syncthreads();
        if (tid == 0) {
            const float thr = a.thr_base + a.thr_step * (float)sc[LCC_SC_N_GENERATED];
            const float pr = expf(a.logits_proc[a.thr_token] - mx) / bcast[1];
            if (pr <= thr) a.logits_proc[a.thr_token] = -INFINITY;
        }
        __syncthreads();
    }

    // 3. argmax, lowest index on ties. float4 loads, 8 independent loads in flight per thread.
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    const int V4 = a.V >> 2;
    const float4* lp4 = reinterpret_cast<const float4*>(a.logits_proc);
#pragma unroll 8
    for (int i = tid; i < V4; i += blockDim.x) {
        const float4 v = lp4[i];
        const int b = i << 2;
        if (v.x > bv) { bv = v.x; bi = b; }  // indices increase per thread => first max kept
        if (v.y > bv) { bv = v.y; bi = b + 1; }
        if (v.z > bv) { bv = v.z; bi = b + 2; }
        if (v.w > bv) { bv = v.w; bi = b + 3; }
    }
    for (int i = (V4 << 2) + tid; i < a.V; i += blockDim.x) {
        const float v = a.logits_proc[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = red_v[lane];
        bi = red_i[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) red_i[0] = bi;
    }
    __syncthreads();
    const int tok = red_i[0] == 0x7fffffff ? 0 : red_i[0];

    // 4. bookkeeping (single thread) + 5. embedding row for the next step (all threads)
    if (tid == 0) {
        a.seq[hist] = tok;
        sc[LCC_SC_SEQ_LEN] = hist + 1;
        const int n = sc[LCC_SC_N_GENERATED] + 1;
        sc[LCC_SC_N_GENERATED] = n;
        sc[LCC_SC_LAST_TOKEN] = tok;
        if (a.advance_kv) {
            sc[LCC_SC_KV_LEN] += 1;
            sc[LCC_SC_ROPE_POS] += 1;
        }
        if (tok == a.eos_token_id || tok == a.eos_token_id2 || n >= a.max_new_tokens) sc[LCC_SC_FINISHED] = 1;
    }
    const bf16* src = a.embed + (size_t)tok * a.H;
    for (int c = tid * 8; c < a.H; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(a.h + c) = *reinterpret_cast<const uint4*>(src + c);
}

__global__ void __launch_bounds__(1024) sample_greedy_kernel(const SampleArgs a) { sample_greedy_body(a); }

// One CTA per stream of a batched decode step: stream b uses row b of the logits / hidden buffers and its own id
// buffer and scalars. Also publishes the decode kernel's sticky error flag in scalars[LCC_SC_NATIVE_ERROR].
__global__ void __launch_bounds__(1024) sample_greedy_batch_kernel(const SampleArgs base, const SampleBatch sb) {
    const int b = blockIdx.x;
    SampleArgs a = base;
    a.logits_raw = base.logits_raw + (size_t)b * base.V;
    a.logits_proc = base.logits_proc + (size_t)b * base.V;
    a.h = base.h + (size_t)b * base.H;
    a.seq = sb.seq[b];
    a.scalars = sb.scalars[b];
    if (sb.err && threadIdx.x == 0) {
        const int e = *sb.err;
        if (e) a.scalars[LCC_SC_NATIVE_ERROR] = e;
    }
    sample_greedy_body(a);
}

int sample_greedy(const SampleArgs& a, cudaStream_t s) {
    if (a.V <= 0 || a.H % 8) return -1;
    lcc::count_launch();
    sample_greedy_kernel<<<1, 1024, 0, s>>>(a);
    return 0;
}

int sample_greedy_batch(const SampleArgs& base, const SampleBatch& sb, cudaStream_t s) {
    if (base.V <= 0 || base.H % 8 || sb.B < 1 || sb.B > 8) return -1;
    lcc::count_launch();
    sample_greedy_batch_kernel<<<sb.B, 1024, 0, s>>>(base, sb);
    return 0;
}

}  // namespace lcc
