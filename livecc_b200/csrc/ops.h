// Internal launch interfaces of the op kernels (elementwise.cu, attention.cu, gemv.cu, sampling.cu).
// All return 0 on success, negative on invalid arguments; launches are asynchronous on `s`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lcc {
typedef __nv_bfloat16 bf16;

// elementwise.cu
int cast_f32_bf16(const float* in, bf16* out, int64_t n, int num_sms, cudaStream_t s);
int patchify_u8(const uint8_t* frames, int T, int H, int W, bf16* out, const float* mean255, const float* std255,
                cudaStream_t s);
int layernorm(const bf16* x, int ldx, const bf16* w, const bf16* b, bf16* y, int ldy, int rows, int dim,
              float eps, cudaStream_t s);
int rmsnorm(const bf16* x, int ldx, const bf16* w, bf16* y, int ldy, int rows, int dim, float eps,
            cudaStream_t s);
int vit_rope_table(float* cos_t, float* sin_t, int t, int h, int w, int merge, int head_dim,
                   const float* inv_freq, cudaStream_t s);
int vit_rope_apply(bf16* qkv, int ld, const float* cos_t, const float* sin_t, int N, int heads, int hd,
                   cudaStream_t s);
int embed_gather(const int64_t* ids, const bf16* table, const bf16* video, int n_video_rows, int64_t video_id, bf16* out,
                 int* rank_ws, int* total_video, int S, int H, int64_t vocab, cudaStream_t s);
int mrope_kv_write(bf16* qkv, int ld, const int* pos3, int S, const float* inv_freq, int sec_t, int sec_h,
                   int Hq, int Hkv, bf16* kc, bf16* vc, const int* page_table, int page_size, int kv_start,
                   cudaStream_t s);

// attention.cu
// n_rows = rows of qkv/out (= cu_seqlens[nseg]); impl: 0 = default, 1 = mma.sync kernels, 2 = tcgen05 kernel
int vit_attention(const bf16* qkv, int ld, int64_t n_rows, bf16* out, int o_ld, const int* cu_seqlens, int nseg,
                  int max_seg_len, int heads, int head_dim, int impl, cudaStream_t s);
// attention_tc.cu
int vit_attention_tc(const bf16* qkv, int ld, int64_t n_rows, bf16* out, int o_ld, const int* cu_seqlens, int nseg,
                     int max_seg_len, int heads, cudaStream_t s);
// impl: 0 = default, 1 = mma.sync kernel, 2 = tcgen05 kernel (LCC_ATTN_*)
int attn_prefill_paged(const bf16* q, int q_ld, const bf16* kc, const bf16* vc, const int* page_table,
                       int page_size, int Hq, int Hkv, int S, int past, bf16* out, int o_ld, float* part_o,
                       float* part_ml, size_t part_capacity_rows, int num_sms, int impl, cudaStream_t s);
// attention_prefill_tc.cu
int attn_prefill_tc(const bf16* q, int q_ld, const bf16* kc, const bf16* vc, const int* page_table, int Hq, int Hkv,
                    int S, int past, bf16* out, int o_ld, float* part_o, float* part_ml, size_t part_capacity_rows,
                    int num_sms, int* nsplit_out, cudaStream_t s);
int attn_decode(bf16* qkv, bf16* kc, bf16* vc, const int* page_table, int page_size, const int* kv_len,
                const int* rope_pos, const int* finished, const float* inv_freq, int Hq, int Hkv, int nsplit,
                float* part_o, float* part_ml, int* counters, bf16* out, bool pdl, cudaStream_t s);

int attn_oproj_decode(bf16* qkv, bf16* kc, bf16* vc, const int* page_table, int page_size, const int* kv_len,
                      const int* rope_pos, const int* finished, const float* inv_freq, int Hq, int Hkv, int nsplit,
                      float* part_o, float* part_ml, int* counters, bf16* attn_out, const bf16* o_w, int o_ldw,
                      int o_N, bf16* h, int* sync, cudaStream_t s);

// gemv.cu
int gemv_norm_bias(const bf16* W, int ldw, const bf16* x, const bf16* norm_w, float eps, const bf16* bias,
                   bf16* out, int N, int K, const int* finished, const void* pf_ptr, size_t pf_bytes, int num_sms, bool pdl,
                   cudaStream_t s);
int gemv_residual(const bf16* W, int ldw, const bf16* x, bf16* h_inout, int N, int K, const int* finished,
                  const void* pf_ptr, size_t pf_bytes, int num_sms, bool pdl, cudaStream_t s);
int gemv_norm_swiglu(const bf16* W_gu, int ldw, const bf16* x, const bf16* norm_w, float eps, bf16* act,
                     int N2, int K, const int* finished, const void* pf_ptr, size_t pf_bytes, int num_sms, bool pdl,
                     cudaStream_t s);
int gemv_norm_logits(const bf16* W, int ldw, const bf16* x, const bf16* norm_w, float eps, float* logits,
                     float* logits_copy, int N, int K, const int* finished, int num_sms, bool pdl, cudaStream_t s);

// sampling.cu
struct SampleArgs {
    const float* logits_raw;  // [V]
    float* logits_proc;       // [V] copy of raw; modified in place
    int V;
    int64_t* seq;             // device [cap]: input ids followed by generated ids
    int* scalars;             // device int32: see LCC_SC_* in include/livecc_b200.h
    float repetition_penalty;
    float inv_repetition_penalty;  // see lcc_sampling
    int thr_token;            // < 0: disabled (ThresholdLogitsProcessor, REF/demo/infer.py:10-23)
    float thr_base, thr_step;
    int eos_token_id;
    int eos_token_id2;        // < 0: none
    int max_new_tokens;
    int advance_kv;           // 1 after a decode forward, 0 after prefill
    const bf16* embed;        // [V, H]
    bf16* h;                  // [H]: receives embed[token] for the next decode step
    int H;
};
int sample_greedy(const SampleArgs& a, cudaStream_t s);
struct SampleBatch {
    int B;
    int64_t* seq[8];
    int* scalars[8];
    const int* err;  // optional: sticky error flag of the decode kernel, copied into scalars[LCC_SC_NATIVE_ERROR]
};
// `base` carries the shared arguments; logits_raw / logits_proc / h are [B][V] / [B][V] / [B][H] row arrays.
int sample_greedy_batch(const SampleArgs& base, const SampleBatch& sb, cudaStream_t s);

}  // namespace lcc
