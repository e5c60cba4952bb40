/*
 * livecc_b200.h — C ABI of liblivecc_sm100a.so, the B200-native (sm_100a) kernels and model runtime
 * behind LiveCC's per-chunk encode -> prefill -> decode hot path.
 *
 * The reference (showlab/livecc) has no native/FFI boundary of its own: the path sits behind Python
 * surfaces (REF/demo/infer.py:43-50,165-174 -> transformers Qwen2VLForConditionalGeneration). This
 * header is therefore the boundary *we* define one level below that surface;
 * each entry point cites the reference interface (file:line) whose arithmetic it replaces.
 *   mq2vl.py     = transformers/models/qwen2_vl/modeling_qwen2_vl.py   (transformers 5.5.0)
 *   gen/utils.py = transformers/generation/utils.py
 *   REF/         = showlab/livecc
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (PyTorch storage in our host code);
 *     the library never allocates device memory.
 *   - Every call is asynchronous on `stream`; nothing synchronises the device.
 *   - Return value: 0 on success, negative on error; lcc_last_error(ctx) returns a message.
 *   - A ctx is bound to one device and one host thread at a time; no hidden globals.
 *   - bf16 everywhere unless the name says otherwise; "bf16 rounding points" follow the reference's
 *     eager graph (one torch op = one rounding).
 *   - The paged KV cache uses pages of LCC_PAGE_SIZE tokens; a layer's K (or V) pool is
 *     [num_pages, kv_heads, LCC_PAGE_SIZE, 128] bf16.
 */
#ifndef LIVECC_B200_H
#define LIVECC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCC_ABI_VERSION 8
#define LCC_PAGE_SIZE 64

typedef struct lcc_ctx lcc_ctx;
typedef struct lcc_model lcc_model;
typedef void* lcc_stream_t; /* cudaStream_t */

/* GEMM epilogues (lcc_gemm_bf16) */
#define LCC_EPI_NONE 0            /* C = bf16(acc) */
#define LCC_EPI_BIAS 1            /* C = bf16(acc + bias) */
#define LCC_EPI_BIAS_QUICKGELU 2  /* ViT fc1: mq2vl.py:329-337 */
#define LCC_EPI_BIAS_GELU 3       /* merger mlp[0..1]: mq2vl.py:317-326 */
#define LCC_EPI_RESIDUAL 4        /* o_proj / down_proj + residual: mq2vl.py:645-660 */
#define LCC_EPI_BIAS_RESIDUAL 5   /* ViT proj / fc2 + residual: mq2vl.py:479-487 */
#define LCC_EPI_SWIGLU 6          /* gate|up (16-row interleaved) -> silu(gate)*up: mq2vl.py:502-504 */

/* Device-side per-stream scalars (int32[LCC_SC_COUNT]); they replace the host-side loop state of
 * GenerationMixin._sample (gen/utils.py:2743-2805) so that a generate() call needs no host sync. */
#define LCC_SC_KV_LEN 0       /* tokens in the KV cache */
#define LCC_SC_ROPE_POS 1     /* position id of the next token to forward = kv_len + rope_delta (mq2vl.py:1212-1222) */
#define LCC_SC_FINISHED 2     /* non-zero once EOS was produced or max_new_tokens reached */
#define LCC_SC_N_GENERATED 3  /* tokens generated in this generate() call */
#define LCC_SC_SEQ_LEN 4      /* ids in the sequence buffer (history + generated) */
#define LCC_SC_LAST_TOKEN 5
#define LCC_SC_VIDEO_TOKENS 6 /* number of video placeholder ids seen by the last lcc_prefill (mq2vl.py:1169-1175 check) */
#define LCC_SC_NATIVE_ERROR 7 /* non-zero: a bounded wait inside the persistent decode kernel gave up (results invalid) */
#define LCC_SC_COUNT 8

int lcc_abi_version(void);

/* ---- context ---------------------------------------------------------------------------- */
/* lcc_create returns NULL unless `device` is an sm_100 GPU (there is no fallback path). */
lcc_ctx* lcc_create(int device);
void lcc_destroy(lcc_ctx* ctx);
const char* lcc_last_error(lcc_ctx* ctx);
int lcc_num_sms(lcc_ctx* ctx);
/* Kernel launches issued by this library in this process so far (launches recorded into a CUDA graph count once,
 * at capture). */
uint64_t lcc_launch_count(void);

/* ---- op level (each is unit-tested against the oracle) ---------------------------------- */

/* C[M,N] = epilogue(A[M,K] x B[N,K]^T), bf16 in, fp32 accumulate (tcgen05/TMEM), bf16 out.
 * Replaces every nn.Linear / Conv3d-as-GEMM call with M >= 16 on the path:
 * mq2vl.py:304-310 (patch embed), :385,:401-403,:456 (ViT qkv/proj), :329-337 (ViT MLP),
 * :317-326 (merger), :539-541,:559-565,:593 (decoder q/k/v/o), :502-504 (MLP).
 * K, N and all leading dimensions must be multiples of 8 elements. block_n (N tile): 0 = dispatch rule, else a multiple of 16 in [32,256] (SwiGLU: of 32);
 * a negative value selects the CTA-pair kernel (cta_group::2, 256 x |block_n| tiles, |block_n| a multiple of 32).
 * splitk_ws: optional fp32 scratch (>= 8*M*N*4 bytes) that allows split-K for M <= 384 when the tile count cannot
 * occupy the SMs; used for K >= 8192 by default (LIVECC_B200_GEMM_SPLITK=1 forces it for every eligible shape, =0 disables). NULL = never. */
int lcc_gemm_bf16(lcc_ctx* ctx, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                  int M, int N, int K, const void* bias, const void* residual, int ldr, int epilogue,
                  int block_n, void* splitk_ws, int64_t splitk_ws_bytes, lcc_stream_t stream);

/* hidden_states.to(bf16) of the f32 patch rows (mq2vl.py:309). */
int lcc_cast_f32_bf16(lcc_ctx* ctx, const float* in, void* out, int64_t n, lcc_stream_t stream);

/* nn.LayerNorm(dim, eps) over rows (ViT norm1/norm2/ln_q, mq2vl.py:464-465,317). */
int lcc_layernorm(lcc_ctx* ctx, const void* x, int ldx, const void* w, const void* b, void* y, int ldy,
                  int rows, int dim, float eps, lcc_stream_t stream);

/* Qwen2VLRMSNorm over rows (mq2vl.py:126-131). */
int lcc_rmsnorm(lcc_ctx* ctx, const void* x, int ldx, const void* w, void* y, int ldy, int rows, int dim,
                float eps, lcc_stream_t stream);

/* ViT 2-D rotary: table for a (t,h,w) grid in merge-window patch order (rot_pos_emb, mq2vl.py:725-752;
 * inv_freq[head_dim/4] fp32 from VisionRotaryEmbedding :271-284), then in-place application to the
 * q and k thirds of qkv [N, 3*heads*head_dim] (apply_rotary_pos_emb_vision, :257-268). */
int lcc_vit_rope_table(lcc_ctx* ctx, float* cos_t, float* sin_t, int t, int h, int w, int merge,
                       int head_dim, const float* inv_freq, lcc_stream_t stream);
int lcc_vit_rope_apply(lcc_ctx* ctx, void* qkv, int ld, const float* cos_t, const float* sin_t, int N,
                       int heads, int head_dim, lcc_stream_t stream);

/* VisionAttention core (mq2vl.py:392-454): non-causal attention inside each cu_seqlens segment,
 * q/k/v read from qkv [n_rows, 3*heads*80] (already rotated), out [n_rows, heads*80]. cu_seqlens: device
 * int32[nseg+1] with cu_seqlens[nseg] == n_rows. impl: LCC_VIT_ATTN_DEFAULT, or force one kernel family. */
#define LCC_VIT_ATTN_DEFAULT 0
#define LCC_VIT_ATTN_MMA 1 /* mma.sync flash kernels */
#define LCC_VIT_ATTN_TC 2  /* tcgen05 + TMEM kernel */
int lcc_vit_attention(lcc_ctx* ctx, const void* qkv, int ld, int64_t n_rows, void* out, int o_ld,
                      const int32_t* cu_seqlens, int nseg, int max_seg_len, int heads, int head_dim, int impl,
                      lcc_stream_t stream);

/* embed_tokens gather + masked_scatter of the video features (mq2vl.py:1255-1272).
 * ids: device int64[S]; video_embeds: bf16 [n_video_rows, H] or NULL (then placeholders keep their text embedding,
 * as in the reference when no pixel values are passed); placeholders beyond n_video_rows also keep their text
 * embedding (nothing is read out of bounds; the caller compares the count and raises like mq2vl.py:1169-1175).
 * rank_ws: device int32[S+1] scratch (last entry receives the video-token count). */
int lcc_embed_gather(lcc_ctx* ctx, const int64_t* ids, const void* table, const void* video_embeds,
                     int n_video_rows, int64_t video_token_id, void* out, int32_t* rank_ws, int S, int H, int64_t vocab,
                     lcc_stream_t stream);

/* M-RoPE (mq2vl.py:188-201,212-254) on the q and k parts of qkv [S, (Hq+2Hkv)*128] (q rotated in
 * place) and append of rotated k / v to the paged cache at positions kv_start..kv_start+S-1
 * (DynamicLayer.update, cache_utils.py:102-121). pos3: device int32[3,S]; inv_freq: device f32[64]. */
int lcc_mrope_kv_write(lcc_ctx* ctx, void* qkv, int ld, const int32_t* pos3, int S, const float* inv_freq,
                       int sec_t, int sec_h, int Hq, int Hkv, void* k_cache, void* v_cache,
                       const int32_t* page_table, int kv_start, lcc_stream_t stream);

/* Causal GQA attention of S new tokens over past+S cached tokens (Qwen2VLAttention core,
 * mq2vl.py:572-594; eager_attention_forward :353-375). q: rotated rows of qkv; out [S, Hq*128].
 * part_o / part_ml: optional split-KV scratch, f32[part_rows,128] and f32[part_rows,2] with
 * part_rows >= nsplit*S*Hq (nsplit <= 8); NULL => no KV split. impl: LCC_ATTN_DEFAULT | LCC_ATTN_MMA | LCC_ATTN_TC.
 * The V slots of the last page behind token past+S-1 must hold finite values (lcc_mrope_kv_write zeroes them). */
#define LCC_ATTN_DEFAULT 0
#define LCC_ATTN_MMA 1 /* mma.sync flash kernel */
#define LCC_ATTN_TC 2  /* tcgen05 + TMEM kernel */
int lcc_attn_prefill(lcc_ctx* ctx, const void* q, int q_ld, const void* k_cache, const void* v_cache,
                     const int32_t* page_table, int Hq, int Hkv, int S, int past, void* out, int o_ld,
                     float* part_o, float* part_ml, int64_t part_rows, int impl, lcc_stream_t stream);

/* One-token attention: RoPE of q, RoPE + append of the new k/v at slot scalars[LCC_SC_KV_LEN],
 * split-KV attention over the paged cache and merge. qkv: raw projections(+bias) of the new token.
 * part_o: f32[nsplit,Hq,128], part_ml: f32[nsplit,Hq,2] scratch; counters: int32[Hkv], zero-initialised
 * once by the caller (the kernel leaves them zero); out: bf16[Hq*128]. */
int lcc_attn_decode(lcc_ctx* ctx, void* qkv, void* k_cache, void* v_cache, const int32_t* page_table,
                    const int32_t* scalars, const float* inv_freq, int Hq, int Hkv, int nsplit, float* part_o,
                    float* part_ml, int32_t* counters, void* out, lcc_stream_t stream);

/* Decode-step projections (M = 1), each fused with its neighbours in the layer
 * (Qwen2VLDecoderLayer.forward, mq2vl.py:613-662). `scalars` may be NULL (no early-exit flag). */
int lcc_gemv_norm_bias(lcc_ctx* ctx, const void* W, int ldw, const void* x, const void* norm_w, float eps,
                       const void* bias, void* out, int N, int K, const int32_t* scalars, lcc_stream_t stream);
int lcc_gemv_residual(lcc_ctx* ctx, const void* W, int ldw, const void* x, void* h_inout, int N, int K,
                      const int32_t* scalars, lcc_stream_t stream);
int lcc_gemv_norm_swiglu(lcc_ctx* ctx, const void* W_gate_up, int ldw, const void* x, const void* norm_w,
                         float eps, void* act, int N2, int K, const int32_t* scalars, lcc_stream_t stream);
/* final norm + lm_head on one token -> fp32 logits (mq2vl.py:905,1437-1438; gen/utils.py:2762). */
int lcc_gemv_norm_logits(lcc_ctx* ctx, const void* W, int ldw, const void* x, const void* norm_w, float eps,
                         float* logits, float* logits_copy, int N, int K, const int32_t* scalars,
                         lcc_stream_t stream);

/* Logits processing + greedy token selection + stream bookkeeping (gen/utils.py:2762-2805;
 * logits_process.py:407-410; REF/demo/infer.py:10-23). */
typedef struct {
    float repetition_penalty; /* 1.0 = off */
    int32_t thr_token;        /* ThresholdLogitsProcessor token id, < 0 = off */
    float thr_base, thr_step;
    int32_t eos_token_id;
    int32_t max_new_tokens;
    float inv_repetition_penalty; /* fp32(1.0 / (double)penalty): torch's CUDA `tensor / python_float` multiplies by
                                     this value (measured: tools/penalty_probe.py); 0 = derive as 1.0f / repetition_penalty */
    int32_t eos_token_id2;        /* second stop id (generation_config.json eos_token_id = [151645, 151643] in the
                                     Qwen2-VL family; gen/utils.py stops on any of them), < 0 = none */
} lcc_sampling;

int lcc_sample_greedy(lcc_ctx* ctx, const float* logits_raw, float* logits_proc, int V, int64_t* seq,
                      int32_t* scalars, const lcc_sampling* sp, int advance_kv, const void* embed, void* h,
                      int H, lcc_stream_t stream);

/* ---- model level (native runtime: one call launches a whole phase) ---------------------- */
typedef struct {
    /* vision (Qwen2VLVisionConfig) */
    int32_t vit_depth, vit_dim, vit_heads, vit_mlp, patch_dim, merge, vit_out;
    /* text (Qwen2VLTextConfig) */
    int32_t hidden, inter, layers, q_heads, kv_heads, vocab;
    float rms_eps, rope_theta;
    int32_t mrope_t, mrope_h; /* mrope_section[0], [1] */
    int64_t video_token_id;
} lcc_model_config;

typedef struct {
    const void *norm1_w, *norm1_b, *norm2_w, *norm2_b;
    const void *qkv_w, *qkv_b, *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} lcc_vit_block_weights;

typedef struct {
    const void *ln1_w, *qkv_w, *qkv_b, *o_w, *ln2_w, *gate_up_w /* 16-row interleaved */, *down_w;
} lcc_layer_weights;

typedef struct {
    const void* patch_w; /* [vit_dim, patch_dim] */
    const lcc_vit_block_weights* vit_blocks;
    const void *merger_ln_w, *merger_ln_b, *merger_fc1_w, *merger_fc1_b, *merger_fc2_w, *merger_fc2_b;
    const void* embed; /* [vocab, hidden] */
    const lcc_layer_weights* layers;
    const void* final_norm_w;
    const void* lm_head; /* [vocab, hidden] */
    /* fp32 rotary tables computed by the host exactly as the reference does:
     * text_inv_freq[64]  = 1/theta^(2i/128)            (mq2vl.py:175-183)
     * vit_inv_freq[hd/4] = 1/10000^(2i/(hd/2))         (mq2vl.py:274-279) */
    const float* text_inv_freq;
    const float* vit_inv_freq;
} lcc_model_weights;

/* One video stream's device state (all caller-owned). */
typedef struct {
    void* k_pool;               /* [layers][num_pages][kv_heads][LCC_PAGE_SIZE][128] bf16 */
    void* v_pool;
    int64_t layer_stride;       /* elements between consecutive layers of a pool */
    const int32_t* page_table;  /* device int32: logical page -> physical page */
    int32_t* scalars;           /* device int32[LCC_SC_COUNT] */
    int64_t* seq;               /* device int64[cap]: input ids then generated ids */
} lcc_stream_state;

lcc_model* lcc_model_create(lcc_ctx* ctx, const lcc_model_config* cfg, const lcc_model_weights* w);
void lcc_model_destroy(lcc_model* m);

/* Scratch bytes needed by the calls below for at most `max_patches` ViT rows / `max_tokens` prefill rows,
 * and binding of a caller-owned scratch buffer of that size to the model. */
size_t lcc_workspace_bytes(const lcc_model* m, int max_patches, int max_tokens);
int lcc_model_bind_workspace(lcc_model* m, void* ws, size_t ws_bytes, int max_patches, int max_tokens);

/* Qwen2VisionTransformerPretrainedModel.forward (mq2vl.py:757-795) for one video of grid (t,h,w):
 * pixel_values f32 [t*h*w, patch_dim] -> out bf16 [t*h*w/merge^2, vit_out]. */
int lcc_vit_forward(lcc_model* m, const float* pixel_values, int t, int h, int w, void* out,
                    lcc_stream_t stream);

/* Same, starting from the resized uint8 frames [T,3,H,W] on the device (H, W multiples of 28): the
 * rescale/normalize/patchify of Qwen2VLVideoProcessor._preprocess (video_processing_qwen2_vl.py:240-272) and the
 * bf16 cast of PatchEmbed (mq2vl.py:309) are fused into one kernel; bit-identical to the f32 path.
 * mean255/std255: HOST float[3] = fp32(image_mean)*255, fp32(image_std)*255 (image_processing_backends.py:301-304). */
int lcc_vit_forward_frames(lcc_model* m, const uint8_t* frames, int T, int H, int W, const float* mean255,
                           const float* std255, void* out, lcc_stream_t stream);

/* Prefill of S new tokens (Qwen2VLModel.forward + lm_head on the last token, mq2vl.py:1230-1300,
 * 828-910, 1437) followed by the first token selection. ids: device int64[S] (the new tokens);
 * pos3: device int32[3,S]; video_embeds: bf16 [n_video_rows, hidden] or NULL; past = tokens already
 * cached. Host must have set scalars {KV_LEN = past+S, ROPE_POS, FINISHED=0, N_GENERATED=0, SEQ_LEN}.
 * slot (0..7): which row of the decode-step buffers receives the first token's embedding and logits; 0 for the
 * one-stream path, b for stream b of a lcc_decode_batch group. */
int lcc_prefill(lcc_model* m, const lcc_stream_state* st, const int64_t* ids, const int32_t* pos3, int S,
                int past, const void* video_embeds, int n_video_rows, const lcc_sampling* sp, int slot, lcc_stream_t stream);

/* n_steps x (one-token forward + token selection) (the loop body of _sample, gen/utils.py:2743-2805).
 * No-ops once scalars[LCC_SC_FINISHED] is set. Capturable in a CUDA graph. nsplit: KV splits (1..64). */
int lcc_decode_steps(lcc_model* m, const lcc_stream_state* st, int n_steps, int nsplit, const lcc_sampling* sp,
                     lcc_stream_t stream);

/* Batched decode (the reference demo admits 5 concurrent sessions on one model,
 * REF/demo/app.py:178): n_steps x (one persistent kernel = every decoder layer + lm_head for all n_streams <= 8 streams,
 * each weight byte read once per step, then one token selection per stream). states[b] must have been prefilled with
 * slot = b and share one page pool. A stream's ids and logits are bit-identical to decoding it alone (same kernel,
 * same reduction order). lcc_decode_steps (one stream) uses the per-op kernels unless LIVECC_B200_MEGA=1: they are ~5 % faster
 * for a single stream (measured); with 2..8 streams the persistent kernel reads every weight byte once per step for all. */
int lcc_decode_batch(lcc_model* m, const lcc_stream_state* states, int n_streams, int n_steps, const lcc_sampling* sp,
                     lcc_stream_t stream);
/* Test hook: run layers [layer_begin, layer_end) and the phases in phase_mask (1 qkv, 2 attention, 4 o_proj,
 * 8 gate/up, 16 down_proj) of the persistent decode kernel, optionally the lm_head, without token selection. */
int lcc_decode_mega_debug(lcc_model* m, const lcc_stream_state* states, int n_streams, int layer_begin, int layer_end,
                          int phase_mask, int do_head, lcc_stream_t stream);

/* ---- frame ingest ------------------------------------------------------------------------------
 * transforms.functional.resize(clip, [H, W], interpolation=BICUBIC, antialias=True) on a uint8 TCHW clip
 * (REF/livecc-utils/src/livecc_utils/video_process_patch.py:101-106 and :150-155), bit-identical to torchvision 0.26 /
 * ATen's CPU kernel: float32 width pass, float32 height pass, clamp, round half to even, uint8. */
/* Window/weight table of one axis (ATen _compute_indices_min_size_weights_aa<float>, Keys cubic a = -0.5), host only:
 * lcc_resize_aa_taps = max window length; xmin/xsize: int32[out_size]; weights: float[taps][out_size] (tap-major). */
int lcc_resize_aa_taps(int in_size, int out_size);
int lcc_resize_aa_table(int in_size, int out_size, int32_t* xmin, int32_t* xsize, float* weights);
/* A plan owns the device copy of both axes' tables for (h, w) -> (H, W). rows_per_cta: 0 = choose (tuning hook).
 * NULL (and lcc_last_error) when a window does not fit in shared memory or a size is out of range. */
typedef struct lcc_resize_plan lcc_resize_plan;
lcc_resize_plan* lcc_resize_plan_create(lcc_ctx* ctx, int h, int w, int H, int W, int rows_per_cta);
void lcc_resize_plan_destroy(lcc_resize_plan* plan);
int lcc_resize_plan_info(const lcc_resize_plan* plan, int* rows_per_cta, int* max_rows, int64_t* smem_bytes);
/* src: device uint8 [planes, h, w] (planes = T*C), dst: device uint8 [planes, H, W]. One kernel. */
int lcc_resize_bicubic_aa_u8(lcc_ctx* ctx, const lcc_resize_plan* plan, const uint8_t* src, int planes, uint8_t* dst,
                             lcc_stream_t stream);

/* Debug/parity hooks: byte offsets of buffers inside the bound workspace. */
#define LCC_WS_PREFILL_HIDDEN 0 /* bf16 [S, hidden] residual stream of the last prefill */
#define LCC_WS_LOGITS 1         /* f32 [8][vocab] raw logits of the last token selection (row = slot) */
#define LCC_WS_DECODE_HIDDEN 2  /* bf16 [8][hidden] embedding row / residual stream of the decode step (row = slot) */
#define LCC_WS_DECODE_QKV 3     /* bf16 [8][(Hq+2Hkv)*128] */
#define LCC_WS_DECODE_ATTN 4    /* bf16 [8][Hq*128] */
#define LCC_WS_DECODE_ACT 5     /* bf16 [8][inter] */
#define LCC_WS_LOGITS_PROC 6    /* f32 [8][vocab] processed logits */
#define LCC_WS_MEGA_ERROR 7     /* int32: sticky error flag of the persistent decode kernel (0 = ok) */
#define LCC_WS_MEGA_TRACE 8     /* u64 [256][64]: per-CTA globaltimer stamps when LIVECC_B200_MEGA_TRACE=1 */
size_t lcc_ws_offset(const lcc_model* m, int which);

#ifdef __cplusplus
}
#endif
#endif /* LIVECC_B200_H */
