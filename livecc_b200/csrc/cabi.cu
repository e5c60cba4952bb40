// C-ABI entry points of liblivecc_sm100a.so (declared in include/livecc_b200.h).
// Plain pointers and sizes only; never throws; every call is asynchronous on the given stream.
#include "../../include/livecc_b200.h"
#include "cabi_common.h"
#include "gemm.h"
#include "launch.h"
#include "ops.h"
#include "resize.h"

using lcc::bf16;

extern "C" {

int lcc_abi_version(void) { return LCC_ABI_VERSION; }

lcc_ctx* lcc_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
    if (prop.major != 10) return nullptr;  // sm_100a only: there is no fallback path
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    lcc_ctx* c = new lcc_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->err[0] = 0;
    return c;
}

void lcc_destroy(lcc_ctx* ctx) { delete ctx; }

const char* lcc_last_error(lcc_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

int lcc_num_sms(lcc_ctx* ctx) { return ctx ? ctx->num_sms : -1; }

uint64_t lcc_launch_count(void) { return lcc::g_launches.load(std::memory_order_relaxed); }

int lcc_gemm_bf16(lcc_ctx* ctx, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                  int M, int N, int K, const void* bias, const void* residual, int ldr, int epilogue,
                  int block_n, void* splitk_ws, int64_t splitk_ws_bytes, lcc_stream_t stream) {
    if (!ctx) return -1;
    lcc::GemmArgs a;
    a.splitk_ws = splitk_ws;
    a.splitk_ws_bytes = splitk_ws_bytes > 0 ? (size_t)splitk_ws_bytes : 0;
    a.A = A; a.B = B; a.C = C;
    a.M = M; a.N = N; a.K = K;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.bias = bias; a.residual = residual; a.ldr = ldr;
    a.epi = epilogue; a.block_n = block_n;
    int r = lcc::gemm_bf16_tn(a, ctx->num_sms, (cudaStream_t)stream);
    if (r) LCC_FAIL(ctx, r, "lcc_gemm_bf16 failed (code %d; M=%d N=%d K=%d epi=%d)", r, M, N, K, epilogue);
    return 0;
}


#define OP_RET(ctx, expr, name)                                              \
    do {                                                                     \
        if (!(ctx)) return -1;                                               \
        int rc__ = (expr);                                                   \
        if (rc__) LCC_FAIL(ctx, rc__, name " failed (code %d)", rc__);       \
        LCC_CHECK_LAUNCH(ctx, name);                                         \
        return 0;                                                            \
    } while (0)

int lcc_cast_f32_bf16(lcc_ctx* ctx, const float* in, void* out, int64_t n, lcc_stream_t stream) {
    OP_RET(ctx, lcc::cast_f32_bf16(in, (bf16*)out, n, ctx->num_sms, (cudaStream_t)stream), "lcc_cast_f32_bf16");
}

int lcc_layernorm(lcc_ctx* ctx, const void* x, int ldx, const void* w, const void* b, void* y, int ldy,
                  int rows, int dim, float eps, lcc_stream_t stream) {
    OP_RET(ctx, lcc::layernorm((const bf16*)x, ldx, (const bf16*)w, (const bf16*)b, (bf16*)y, ldy, rows, dim, eps,
                               (cudaStream_t)stream), "lcc_layernorm");
}

int lcc_rmsnorm(lcc_ctx* ctx, const void* x, int ldx, const void* w, void* y, int ldy, int rows, int dim,
                float eps, lcc_stream_t stream) {
    OP_RET(ctx, lcc::rmsnorm((const bf16*)x, ldx, (const bf16*)w, (bf16*)y, ldy, rows, dim, eps, (cudaStream_t)stream),
           "lcc_rmsnorm");
}

int lcc_vit_rope_table(lcc_ctx* ctx, float* cos_t, float* sin_t, int t, int h, int w, int merge,
                       int head_dim, const float* inv_freq, lcc_stream_t stream) {
    OP_RET(ctx, lcc::vit_rope_table(cos_t, sin_t, t, h, w, merge, head_dim, inv_freq, (cudaStream_t)stream),
           "lcc_vit_rope_table");
}

int lcc_vit_rope_apply(lcc_ctx* ctx, void* qkv, int ld, const float* cos_t, const float* sin_t, int N,
                       int heads, int head_dim, lcc_stream_t stream) {
    OP_RET(ctx, lcc::vit_rope_apply((bf16*)qkv, ld, cos_t, sin_t, N, heads, head_dim, (cudaStream_t)stream),
           "lcc_vit_rope_apply");
}

int lcc_vit_attention(lcc_ctx* ctx, const void* qkv, int ld, int64_t n_rows, void* out, int o_ld, const int32_t* cu_seqlens,
                      int nseg, int max_seg_len, int heads, int head_dim, int impl, lcc_stream_t stream) {
    OP_RET(ctx, lcc::vit_attention((const bf16*)qkv, ld, n_rows, (bf16*)out, o_ld, cu_seqlens, nseg, max_seg_len, heads,
                                   head_dim, impl, (cudaStream_t)stream), "lcc_vit_attention");
}

int lcc_embed_gather(lcc_ctx* ctx, const int64_t* ids, const void* table, const void* video_embeds,
                     int n_video_rows, int64_t video_token_id, void* out, int32_t* rank_ws, int S, int H, int64_t vocab,
                     lcc_stream_t stream) {
    OP_RET(ctx, lcc::embed_gather(ids, (const bf16*)table, (const bf16*)video_embeds, n_video_rows, video_token_id, (bf16*)out,
                                  rank_ws, rank_ws + S, S, H, vocab, (cudaStream_t)stream), "lcc_embed_gather");
}

int lcc_mrope_kv_write(lcc_ctx* ctx, void* qkv, int ld, const int32_t* pos3, int S, const float* inv_freq,
                       int sec_t, int sec_h, int Hq, int Hkv, void* k_cache, void* v_cache,
                       const int32_t* page_table, int kv_start, lcc_stream_t stream) {
    OP_RET(ctx, lcc::mrope_kv_write((bf16*)qkv, ld, pos3, S, inv_freq, sec_t, sec_h, Hq, Hkv, (bf16*)k_cache,
                                    (bf16*)v_cache, page_table, LCC_PAGE_SIZE, kv_start, (cudaStream_t)stream),
           "lcc_mrope_kv_write");
}

int lcc_attn_prefill(lcc_ctx* ctx, const void* q, int q_ld, const void* k_cache, const void* v_cache,
                     const int32_t* page_table, int Hq, int Hkv, int S, int past, void* out, int o_ld,
                     float* part_o, float* part_ml, int64_t part_rows, int impl, lcc_stream_t stream) {
    OP_RET(ctx, lcc::attn_prefill_paged((const bf16*)q, q_ld, (const bf16*)k_cache, (const bf16*)v_cache, page_table,
                                        LCC_PAGE_SIZE, Hq, Hkv, S, past, (bf16*)out, o_ld, part_o, part_ml,
                                        (size_t)(part_rows > 0 ? part_rows : 0), ctx->num_sms, impl, (cudaStream_t)stream),
           "lcc_attn_prefill");
}

int lcc_attn_decode(lcc_ctx* ctx, void* qkv, void* k_cache, void* v_cache, const int32_t* page_table,
                    const int32_t* scalars, const float* inv_freq, int Hq, int Hkv, int nsplit, float* part_o,
                    float* part_ml, int32_t* counters, void* out, lcc_stream_t stream) {
    if (nsplit < 1 || nsplit > 64) LCC_FAIL(ctx, -2, "lcc_attn_decode: nsplit out of range");
    OP_RET(ctx, lcc::attn_decode((bf16*)qkv, (bf16*)k_cache, (bf16*)v_cache, page_table, LCC_PAGE_SIZE,
                                 scalars + LCC_SC_KV_LEN, scalars + LCC_SC_ROPE_POS, scalars + LCC_SC_FINISHED,
                                 inv_freq, Hq, Hkv, nsplit, part_o, part_ml, counters, (bf16*)out, false,
                                 (cudaStream_t)stream),
           "lcc_attn_decode");
}

static inline const int* fin_ptr(const int32_t* scalars) { return scalars ? scalars + LCC_SC_FINISHED : nullptr; }

int lcc_gemv_norm_bias(lcc_ctx* ctx, const void* W, int ldw, const void* x, const void* norm_w, float eps,
                       const void* bias, void* out, int N, int K, const int32_t* scalars, lcc_stream_t stream) {
    OP_RET(ctx, lcc::gemv_norm_bias((const bf16*)W, ldw, (const bf16*)x, (const bf16*)norm_w, eps, (const bf16*)bias,
                                    (bf16*)out, N, K, fin_ptr(scalars), nullptr, 0, ctx->num_sms, false, (cudaStream_t)stream),
           "lcc_gemv_norm_bias");
}

int lcc_gemv_residual(lcc_ctx* ctx, const void* W, int ldw, const void* x, void* h_inout, int N, int K,
                      const int32_t* scalars, lcc_stream_t stream) {
    OP_RET(ctx, lcc::gemv_residual((const bf16*)W, ldw, (const bf16*)x, (bf16*)h_inout, N, K, fin_ptr(scalars), nullptr, 0, ctx->num_sms, false,
                                   (cudaStream_t)stream), "lcc_gemv_residual");
}

int lcc_gemv_norm_swiglu(lcc_ctx* ctx, const void* W_gate_up, int ldw, const void* x, const void* norm_w,
                         float eps, void* act, int N2, int K, const int32_t* scalars, lcc_stream_t stream) {
    OP_RET(ctx, lcc::gemv_norm_swiglu((const bf16*)W_gate_up, ldw, (const bf16*)x, (const bf16*)norm_w, eps, (bf16*)act,
                                      N2, K, fin_ptr(scalars), nullptr, 0, ctx->num_sms, false, (cudaStream_t)stream),
           "lcc_gemv_norm_swiglu");
}

int lcc_gemv_norm_logits(lcc_ctx* ctx, const void* W, int ldw, const void* x, const void* norm_w, float eps,
                         float* logits, float* logits_copy, int N, int K, const int32_t* scalars,
                         lcc_stream_t stream) {
    OP_RET(ctx, lcc::gemv_norm_logits((const bf16*)W, ldw, (const bf16*)x, (const bf16*)norm_w, eps, logits,
                                      logits_copy, N, K, fin_ptr(scalars), ctx->num_sms, false, (cudaStream_t)stream),
           "lcc_gemv_norm_logits");
}

int lcc_sample_greedy(lcc_ctx* ctx, const float* logits_raw, float* logits_proc, int V, int64_t* seq,
                      int32_t* scalars, const lcc_sampling* sp, int advance_kv, const void* embed, void* h,
                      int H, lcc_stream_t stream) {
    if (!sp) return -1;
    lcc::SampleArgs a{};
    a.logits_raw = logits_raw; a.logits_proc = logits_proc; a.V = V; a.seq = seq; a.scalars = scalars;
    a.repetition_penalty = sp->repetition_penalty; a.inv_repetition_penalty = sp->inv_repetition_penalty; a.thr_token = sp->thr_token; a.thr_base = sp->thr_base;
    a.thr_step = sp->thr_step; a.eos_token_id = sp->eos_token_id; a.eos_token_id2 = sp->eos_token_id2;
    a.max_new_tokens = sp->max_new_tokens;
    a.advance_kv = advance_kv; a.embed = (const bf16*)embed; a.h = (bf16*)h; a.H = H;
    OP_RET(ctx, lcc::sample_greedy(a, (cudaStream_t)stream), "lcc_sample_greedy");
}

// ---- frame ingest: antialiased bicubic uint8 resize (resize.cu) ----
struct lcc_resize_plan {
    lcc::ResizePlan* impl;
};

int lcc_resize_aa_taps(int in_size, int out_size) { return lcc::resize_aa_taps(in_size, out_size); }

int lcc_resize_aa_table(int in_size, int out_size, int32_t* xmin, int32_t* xsize, float* weights) {
    if (!xmin || !xsize || !weights) return -1;
    return lcc::resize_aa_table(in_size, out_size, xmin, xsize, weights);
}

lcc_resize_plan* lcc_resize_plan_create(lcc_ctx* ctx, int h, int w, int H, int W, int rows_per_cta) {
    if (!ctx) return nullptr;
    lcc::ResizePlan* impl = lcc::resize_plan_create(h, w, H, W, rows_per_cta);
    if (!impl) {
        snprintf(ctx->err, sizeof(ctx->err),
                 "lcc_resize_plan_create(%dx%d -> %dx%d): size out of range, window larger than shared memory, or "
                 "device allocation failed: %s", h, w, H, W, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    lcc_resize_plan* p = new lcc_resize_plan();
    p->impl = impl;
    return p;
}

void lcc_resize_plan_destroy(lcc_resize_plan* plan) {
    if (!plan) return;
    lcc::resize_plan_destroy(plan->impl);
    delete plan;
}

int lcc_resize_plan_info(const lcc_resize_plan* plan, int* rows_per_cta, int* max_rows, int64_t* smem_bytes) {
    if (!plan) return -1;
    lcc::resize_plan_info(plan->impl, rows_per_cta, max_rows, smem_bytes);
    return 0;
}

int lcc_resize_bicubic_aa_u8(lcc_ctx* ctx, const lcc_resize_plan* plan, const uint8_t* src, int planes, uint8_t* dst,
                             lcc_stream_t stream) {
    if (!plan) LCC_FAIL(ctx, -1, "lcc_resize_bicubic_aa_u8: null plan");
    OP_RET(ctx, lcc::resize_bicubic_aa_u8(plan->impl, src, planes, dst, (cudaStream_t)stream), "lcc_resize_bicubic_aa_u8");
}

}  // extern "C"
