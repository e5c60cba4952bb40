// Native model runtime: one C call launches a whole phase of the LiveCC hot path on the given stream.
//   lcc_vit_forward   — Qwen2VisionTransformerPretrainedModel.forward   (mq2vl.py:757-795)
//   lcc_prefill       — Qwen2VLModel.forward over S new tokens + first token selection (mq2vl.py:1230-1300)
//   lcc_decode_steps  — the loop body of GenerationMixin._sample         (gen/utils.py:2743-2805)
// The host (Python) owns every buffer; this file only sequences kernel launches.
#include <stdlib.h>

#include <vector>

#include "../../include/livecc_b200.h"
#include "cabi_common.h"
#include "gemm.h"
#include "launch.h"
#include "mega.h"
#include "ops.h"

using lcc::bf16;

struct lcc_model {
    lcc_ctx* ctx;
    lcc_model_config cfg;
    lcc_model_weights w;
    std::vector<lcc_vit_block_weights> vit_blocks;
    std::vector<lcc_layer_weights> layers;
    // bound workspace (lcc_model_bind_workspace)
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    int cap_patches = 0, cap_tokens = 0;
    int mega_lookahead = 0;  // LIVECC_B200_MEGA_LOOKAHEAD: 16 KB groups per CTA an L2 prefetcher thread runs ahead of the ring (measured slower: off)
    bool mega_trace = false;  // LIVECC_B200_MEGA_TRACE=1: per-CTA phase timeline in the workspace (LCC_WS_MEGA_TRACE)
    bool use_mega = false;  // one-stream decode through the persistent kernel (LIVECC_B200_MEGA=1); batched decode always uses it
    bool mega_ok = false;   // geometry supported by the persistent kernel and its tables are uploaded
    bool fuse_attn_oproj = false;  // LIVECC_B200_FUSE=1: one launch for decode attention + o_proj (flag-synchronised roles)
    bool use_pdl = false;  // programmatic dependent launch between the decode-step kernels (LIVECC_B200_PDL=1 enables)
};

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Workspace carve-up (deterministic function of the model config and the two capacities).
struct WsLayout;
int get_layout(const lcc_model* m, int need_patches, int need_tokens, WsLayout* out);

struct WsLayout {
    size_t vx, vh, vn, vqkv, vattn, vmlp, vcos, vsin, vcu;  // ViT
    size_t hid, normed, qkv, attn, act, rank, pf_part_o, pf_part_ml, pf_splitk;  // prefill
    size_t h1, qkv1, attn1, act1, logits_raw, logits_proc, part_o, part_ml, attn_cnt;  // decode (8 stream rows each)
    size_t mg_part_o, mg_part_ml, mg_cnt, mg_tmaps, mg_layers, mg_trace;  // persistent decode kernel
    size_t total;
};

constexpr int kMaxSplit = 64;
constexpr size_t kSplitKRows = 8 * 384;  // split-K GEMM partials: splits (<= 8) x M (<= 384) fp32 rows of the widest projection
constexpr size_t kPrefillSplitRows = 8 * 1024;  // split-KV prefill partials: nsplit * S <= this many positions

WsLayout make_layout(const lcc_model_config& c, int NP, int NT) {
    WsLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t np = (size_t)(NP > 0 ? NP : 1), nt = (size_t)(NT > 0 ? NT : 1);
    const size_t qkv_dim = (size_t)(c.q_heads + 2 * c.kv_heads) * 128;
    L.vx = take(np * c.patch_dim * 2);
    L.vh = take(np * c.vit_dim * 2);
    L.vn = take(np * c.vit_dim * 2);
    L.vqkv = take(np * 3 * c.vit_dim * 2);
    L.vattn = take(np * c.vit_dim * 2);
    L.vmlp = take(np * c.vit_mlp * 2);
    L.vcos = take(np * (c.vit_dim / c.vit_heads / 2) * 4);
    L.vsin = take(np * (c.vit_dim / c.vit_heads / 2) * 4);
    L.vcu = take((np + 2) * 4);
    L.hid = take(nt * c.hidden * 2);
    L.normed = take(nt * c.hidden * 2);
    L.qkv = take(nt * qkv_dim * 2);
    L.attn = take(nt * c.q_heads * 128 * 2);
    L.act = take(nt * c.inter * 2);
    L.rank = take((nt + 2) * 4);
    L.pf_part_o = take(kPrefillSplitRows * (size_t)c.q_heads * 128 * 4);
    L.pf_part_ml = take(kPrefillSplitRows * (size_t)c.q_heads * 2 * 4);
    L.pf_splitk = take(kSplitKRows * (qkv_dim > (size_t)c.hidden ? qkv_dim : (size_t)c.hidden) * 4);
    // decode-step vectors: row b belongs to stream slot b of a batched step (slot 0 = the single-stream path)
    L.h1 = take((size_t)lcc::MG_MAXB * c.hidden * 2);
    L.qkv1 = take((size_t)lcc::MG_MAXB * qkv_dim * 2);
    L.attn1 = take((size_t)lcc::MG_MAXB * c.q_heads * 128 * 2);
    L.act1 = take((size_t)lcc::MG_MAXB * c.inter * 2);
    L.logits_raw = take((size_t)lcc::MG_MAXB * c.vocab * 4);
    L.logits_proc = take((size_t)lcc::MG_MAXB * c.vocab * 4);
    L.part_o = take((size_t)kMaxSplit * c.q_heads * 128 * 4);
    L.part_ml = take((size_t)kMaxSplit * c.q_heads * 2 * 4);
    L.attn_cnt = take(64 * 4);
    L.mg_part_o = take((size_t)lcc::MG_MAXB * c.kv_heads * lcc::MG_MAX_ITEMS * 8 * 128 * 4);
    L.mg_part_ml = take((size_t)lcc::MG_MAXB * c.kv_heads * lcc::MG_MAX_ITEMS * 8 * 2 * 4);
    L.mg_cnt = take(128 * 4);  // [0, 8*kv_heads): pair counters; [96]: grid barrier; [97]: sticky error flag
    L.mg_tmaps = take((size_t)(4 * c.layers + 1) * sizeof(CUtensorMap));
    L.mg_layers = take((size_t)c.layers * sizeof(lcc::MegaLayer));
    L.mg_trace = take((size_t)256 * 64 * 8);
    L.total = off;
    return L;
}

__global__ void fill_cu_seqlens_kernel(int* cu, int t, int hw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= t) cu[i] = i * hw;
}

int gemm(lcc_model* m, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
         const void* bias, const void* res, int ldr, int epi, cudaStream_t s, void* splitk_ws = nullptr,
         size_t splitk_ws_bytes = 0) {
    lcc::GemmArgs a;
    a.A = A; a.B = B; a.C = C; a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.bias = bias; a.residual = res; a.ldr = ldr; a.epi = epi; a.block_n = 0;
    a.splitk_ws = splitk_ws; a.splitk_ws_bytes = splitk_ws_bytes;  // used only with LIVECC_B200_GEMM_SPLITK=1
    return lcc::gemm_bf16_tn(a, m->ctx->num_sms, s);
}

#define STEP(expr, what)                                                                      \
    do {                                                                                      \
        int rc__ = (expr);                                                                    \
        if (rc__) LCC_FAIL(m->ctx, rc__, "%s failed (code %d)", what, rc__);                  \
    } while (0)

lcc::SampleArgs make_sample(const lcc_model* m, const lcc_stream_state* st, const lcc_sampling* sp, uint8_t* ws,
                            const WsLayout& L, int advance, int slot = 0) {
    lcc::SampleArgs a{};
    a.logits_raw = reinterpret_cast<const float*>(ws + L.logits_raw) + (size_t)slot * m->cfg.vocab;
    a.logits_proc = reinterpret_cast<float*>(ws + L.logits_proc) + (size_t)slot * m->cfg.vocab;
    a.V = m->cfg.vocab;
    a.seq = st ? st->seq : nullptr;
    a.scalars = st ? st->scalars : nullptr;
    a.repetition_penalty = sp->repetition_penalty;
    a.inv_repetition_penalty = sp->inv_repetition_penalty;
    a.thr_token = sp->thr_token;
    a.thr_base = sp->thr_base;
    a.thr_step = sp->thr_step;
    a.eos_token_id = sp->eos_token_id;
    a.eos_token_id2 = sp->eos_token_id2;
    a.max_new_tokens = sp->max_new_tokens;
    a.advance_kv = advance;
    a.embed = reinterpret_cast<const bf16*>(m->w.embed);
    a.h = reinterpret_cast<bf16*>(ws + L.h1) + (size_t)slot * m->cfg.hidden;
    a.H = m->cfg.hidden;
    return a;
}

}  // namespace

extern "C" {

lcc_model* lcc_model_create(lcc_ctx* ctx, const lcc_model_config* cfg, const lcc_model_weights* w) {
    if (!ctx || !cfg || !w) return nullptr;
    if (cfg->hidden != cfg->q_heads * 128 || cfg->vit_dim != cfg->vit_heads * 80 || cfg->q_heads % cfg->kv_heads ||
        cfg->q_heads / cfg->kv_heads > 8 || cfg->inter % 16 || cfg->merge != 2) {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported model geometry (need head_dim 128 / ViT head_dim 80 / GQA group <= 8)");
        return nullptr;
    }
    lcc_model* m = new lcc_model();
    m->ctx = ctx;
    m->cfg = *cfg;
    m->w = *w;
    m->vit_blocks.assign(w->vit_blocks, w->vit_blocks + cfg->vit_depth);
    m->layers.assign(w->layers, w->layers + cfg->layers);
    m->w.vit_blocks = m->vit_blocks.data();
    m->w.layers = m->layers.data();
    const char* mega_env = getenv("LIVECC_B200_MEGA");
    // Measured on B200 (profiles/r02_decode_paths.md): for ONE stream the per-op kernels are 5 % faster than the persistent
    // kernel (3.21 vs 3.37 ms/step in situ), so lcc_decode_steps keeps them by default; lcc_decode_batch (2..8 streams) is
    // always the persistent kernel.
    m->use_mega = mega_env && mega_env[0] == '1';
    if (const char* te = getenv("LIVECC_B200_MEGA_TRACE")) m->mega_trace = te[0] == '1';
    if (const char* la = getenv("LIVECC_B200_MEGA_LOOKAHEAD")) m->mega_lookahead = atoi(la) < 0 ? 0 : atoi(la);
    const char* fuse_env = getenv("LIVECC_B200_FUSE");
    m->fuse_attn_oproj = fuse_env && fuse_env[0] == '1';
#ifdef LCC_ENABLE_PDL
    const char* pdl_env = getenv("LIVECC_B200_PDL");
    m->use_pdl = pdl_env && pdl_env[0] == '1';
#else
    m->use_pdl = false;  // griddepcontrol is compiled out (see common.cuh)
#endif
    return m;
}

void lcc_model_destroy(lcc_model* m) { delete m; }

size_t lcc_workspace_bytes(const lcc_model* m, int max_patches, int max_tokens) {
    return m ? make_layout(m->cfg, max_patches, max_tokens).total : 0;
}
size_t lcc_ws_offset(const lcc_model* m, int which) {
    if (!m) return 0;
    const WsLayout L = make_layout(m->cfg, m->cap_patches, m->cap_tokens);
    switch (which) {
        case LCC_WS_PREFILL_HIDDEN: return L.hid;
        case LCC_WS_LOGITS: return L.logits_raw;
        case LCC_WS_DECODE_HIDDEN: return L.h1;
        case LCC_WS_DECODE_QKV: return L.qkv1;
        case LCC_WS_DECODE_ATTN: return L.attn1;
        case LCC_WS_DECODE_ACT: return L.act1;
        case LCC_WS_LOGITS_PROC: return L.logits_proc;
        case LCC_WS_MEGA_ERROR: return L.mg_cnt + 97 * 4;
        case LCC_WS_MEGA_TRACE: return L.mg_trace;
    }
    return 0;
}

// Shared body of the two ViT entry points: `pixel_values` (f32 patch rows from the HF processor) or
// `frames` (uint8 [T,3,H,W] on the device: fused normalize + patchify, no f32 round trip).
static int vit_forward_impl(lcc_model* m, const float* pixel_values, const uint8_t* frames, int T_frames, int t,
                            int h, int w, const float* mean255, const float* std255, void* out, lcc_stream_t stream) {
    if (!m) return -1;
    const lcc_model_config& c = m->cfg;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = t * h * w;
    if (N <= 0 || (h % c.merge) || (w % c.merge)) LCC_FAIL(m->ctx, -2, "lcc_vit_forward: bad grid %d,%d,%d", t, h, w);
    WsLayout L;
    if (get_layout(m, N, 0, &L)) LCC_FAIL(m->ctx, -3, "lcc_vit_forward: no workspace bound or too small for %d patches", N);
    uint8_t* ws = m->ws;
    bf16* vx = (bf16*)(ws + L.vx); bf16* vh = (bf16*)(ws + L.vh); bf16* vn = (bf16*)(ws + L.vn);
    bf16* vqkv = (bf16*)(ws + L.vqkv); bf16* vattn = (bf16*)(ws + L.vattn); bf16* vmlp = (bf16*)(ws + L.vmlp);
    float* vcos = (float*)(ws + L.vcos); float* vsin = (float*)(ws + L.vsin);
    int* vcu = (int*)(ws + L.vcu);
    const int dim = c.vit_dim, hd = dim / c.vit_heads;

    if (frames) {
        if (c.patch_dim != 1176) LCC_FAIL(m->ctx, -4, "frame ingest is specialised for 3x2x14x14 patches");
        STEP(lcc::patchify_u8(frames, T_frames, h * 14, w * 14, vx, mean255, std255, s), "vit frame ingest");
    } else {
        STEP(lcc::cast_f32_bf16(pixel_values, vx, (int64_t)N * c.patch_dim, m->ctx->num_sms, s), "vit cast");
    }
    STEP(gemm(m, vx, c.patch_dim, m->w.patch_w, c.patch_dim, vh, dim, N, dim, c.patch_dim, nullptr, nullptr, 0,
              lcc::EPI_NONE, s), "vit patch_embed");
    STEP(lcc::vit_rope_table(vcos, vsin, t, h, w, c.merge, hd, m->w.vit_inv_freq, s), "vit rope table");
    lcc::count_launch();
    fill_cu_seqlens_kernel<<<(t + 256) / 256, 256, 0, s>>>(vcu, t, h * w);
    for (int i = 0; i < c.vit_depth; ++i) {
        const lcc_vit_block_weights& b = m->vit_blocks[i];
        STEP(lcc::layernorm(vh, dim, (const bf16*)b.norm1_w, (const bf16*)b.norm1_b, vn, dim, N, dim, 1e-6f, s), "vit norm1");
        STEP(gemm(m, vn, dim, b.qkv_w, dim, vqkv, 3 * dim, N, 3 * dim, dim, b.qkv_b, nullptr, 0, lcc::EPI_BIAS, s), "vit qkv");
        STEP(lcc::vit_rope_apply(vqkv, 3 * dim, vcos, vsin, N, c.vit_heads, hd, s), "vit rope");
        STEP(lcc::vit_attention(vqkv, 3 * dim, N, vattn, dim, vcu, t, h * w, c.vit_heads, hd, 0, s), "vit attention");
        STEP(gemm(m, vattn, dim, b.proj_w, dim, vh, dim, N, dim, dim, b.proj_b, vh, dim, lcc::EPI_BIAS_RESIDUAL, s), "vit proj");
        STEP(lcc::layernorm(vh, dim, (const bf16*)b.norm2_w, (const bf16*)b.norm2_b, vn, dim, N, dim, 1e-6f, s), "vit norm2");
        STEP(gemm(m, vn, dim, b.fc1_w, dim, vmlp, c.vit_mlp, N, c.vit_mlp, dim, b.fc1_b, nullptr, 0, lcc::EPI_BIAS_QUICKGELU, s), "vit fc1");
        STEP(gemm(m, vmlp, c.vit_mlp, b.fc2_w, c.vit_mlp, vh, dim, N, dim, c.vit_mlp, b.fc2_b, vh, dim, lcc::EPI_BIAS_RESIDUAL, s), "vit fc2");
    }
    // PatchMerger (mq2vl.py:313-326): LN, view [N/4, 4*dim], Linear+GELU, Linear
    const int md = dim * c.merge * c.merge, NM = N / (c.merge * c.merge);
    STEP(lcc::layernorm(vh, dim, (const bf16*)m->w.merger_ln_w, (const bf16*)m->w.merger_ln_b, vn, dim, N, dim, 1e-6f, s), "merger ln");
    STEP(gemm(m, vn, md, m->w.merger_fc1_w, md, vmlp, md, NM, md, md, m->w.merger_fc1_b, nullptr, 0, lcc::EPI_BIAS_GELU, s), "merger fc1");
    STEP(gemm(m, vmlp, md, m->w.merger_fc2_w, md, out, c.vit_out, NM, c.vit_out, md, m->w.merger_fc2_b, nullptr, 0, lcc::EPI_BIAS, s), "merger fc2");
    LCC_CHECK_LAUNCH(m->ctx, "lcc_vit_forward");
    return 0;
}

int lcc_vit_forward(lcc_model* m, const float* pixel_values, int t, int h, int w, void* out,
                    lcc_stream_t stream) {
    return vit_forward_impl(m, pixel_values, nullptr, 0, t, h, w, nullptr, nullptr, out, stream);
}

int lcc_vit_forward_frames(lcc_model* m, const uint8_t* frames, int T, int H, int W, const float* mean255,
                           const float* std255, void* out, lcc_stream_t stream) {
    if (!m) return -1;
    if (!frames || T <= 0 || H % 28 || W % 28 || !mean255 || !std255)
        LCC_FAIL(m->ctx, -2, "lcc_vit_forward_frames: frames must be uint8 [T,3,H,W] with H,W multiples of 28");
    return vit_forward_impl(m, nullptr, frames, T, (T + 1) / 2, H / 14, W / 14, mean255, std255, out, stream);
}

int lcc_prefill(lcc_model* m, const lcc_stream_state* st, const int64_t* ids, const int32_t* pos3, int S,
                int past, const void* video_embeds, int n_video_rows, const lcc_sampling* sp, int slot, lcc_stream_t stream) {
    if (!m || !st || !sp) return -1;
    if (slot < 0 || slot >= lcc::MG_MAXB) LCC_FAIL(m->ctx, -2, "lcc_prefill: slot must be in [0,%d)", lcc::MG_MAXB);
    const lcc_model_config& c = m->cfg;
    cudaStream_t s = (cudaStream_t)stream;
    if (S <= 0) LCC_FAIL(m->ctx, -2, "lcc_prefill: S must be positive");
    WsLayout L;
    if (get_layout(m, 0, S, &L)) LCC_FAIL(m->ctx, -3, "lcc_prefill: no workspace bound or too small for %d tokens", S);
    uint8_t* ws = m->ws;
    bf16* hid = (bf16*)(ws + L.hid); bf16* normed = (bf16*)(ws + L.normed); bf16* qkv = (bf16*)(ws + L.qkv);
    bf16* attn = (bf16*)(ws + L.attn); bf16* act = (bf16*)(ws + L.act); int* rank = (int*)(ws + L.rank);
    const int H = c.hidden, Hq = c.q_heads, Hkv = c.kv_heads, qkv_dim = (Hq + 2 * Hkv) * 128;
    void* skw = ws + L.pf_splitk;  // split-K scratch of the small-N projections (opt-in, see gemm_tcgen05.cu::try_splitk)
    const size_t skb = kSplitKRows * (size_t)(qkv_dim > H ? qkv_dim : H) * 4;

    STEP(lcc::embed_gather(ids, (const bf16*)m->w.embed, (const bf16*)video_embeds, n_video_rows, c.video_token_id, hid, rank,
                           st->scalars + LCC_SC_VIDEO_TOKENS, S, H, c.vocab, s), "embed gather");
    for (int i = 0; i < c.layers; ++i) {
        const lcc_layer_weights& lw = m->layers[i];
        bf16* kc = (bf16*)st->k_pool + (size_t)i * st->layer_stride;
        bf16* vc = (bf16*)st->v_pool + (size_t)i * st->layer_stride;
        STEP(lcc::rmsnorm(hid, H, (const bf16*)lw.ln1_w, normed, H, S, H, c.rms_eps, s), "input_layernorm");
        STEP(gemm(m, normed, H, lw.qkv_w, H, qkv, qkv_dim, S, qkv_dim, H, lw.qkv_b, nullptr, 0, lcc::EPI_BIAS, s, skw, skb), "qkv proj");
        STEP(lcc::mrope_kv_write(qkv, qkv_dim, pos3, S, m->w.text_inv_freq, c.mrope_t, c.mrope_h, Hq, Hkv, kc, vc,
                                 st->page_table, LCC_PAGE_SIZE, past, s), "mrope + kv write");
        STEP(lcc::attn_prefill_paged(qkv, qkv_dim, kc, vc, st->page_table, LCC_PAGE_SIZE, Hq, Hkv, S, past, attn,
                                     Hq * 128, (float*)(ws + L.pf_part_o), (float*)(ws + L.pf_part_ml),
                                     kPrefillSplitRows * (size_t)Hq, m->ctx->num_sms, 0, s), "prefill attention");
        STEP(gemm(m, attn, Hq * 128, lw.o_w, Hq * 128, hid, H, S, H, Hq * 128, nullptr, hid, H, lcc::EPI_RESIDUAL, s, skw, skb), "o_proj");
        STEP(lcc::rmsnorm(hid, H, (const bf16*)lw.ln2_w, normed, H, S, H, c.rms_eps, s), "post_attention_layernorm");
        STEP(gemm(m, normed, H, lw.gate_up_w, H, act, c.inter, S, 2 * c.inter, H, nullptr, nullptr, 0, lcc::EPI_SWIGLU, s), "gate_up");
        STEP(gemm(m, act, c.inter, lw.down_w, c.inter, hid, H, S, H, c.inter, nullptr, hid, H, lcc::EPI_RESIDUAL, s, skw, skb), "down_proj");
    }
    // final norm + lm_head on the last position only (logits_to_keep = 1, gen/utils.py:2487-2491)
    STEP(lcc::gemv_norm_logits((const bf16*)m->w.lm_head, H, hid + (size_t)(S - 1) * H, (const bf16*)m->w.final_norm_w,
                               c.rms_eps, (float*)(ws + L.logits_raw) + (size_t)slot * c.vocab,
                               (float*)(ws + L.logits_proc) + (size_t)slot * c.vocab, c.vocab, H,
                               nullptr, m->ctx->num_sms, false, s), "lm_head");
    STEP(lcc::sample_greedy(make_sample(m, st, sp, ws, L, 0, slot), s), "token selection");
    LCC_CHECK_LAUNCH(m->ctx, "lcc_prefill");
    return 0;
}

// ---- persistent decode-step kernel (decode_mega.cu) --------------------------------------------------------------
static int mega_step_params(lcc_model* m, const lcc_stream_state* sts, int B, const WsLayout& L, lcc::MegaParams* out) {
    const lcc_model_config& c = m->cfg;
    uint8_t* ws = m->ws;
    lcc::MegaParams p{};
    p.wmaps = reinterpret_cast<const CUtensorMap*>(ws + L.mg_tmaps);
    p.layers = reinterpret_cast<const lcc::MegaLayer*>(ws + L.mg_layers);
    p.final_norm_w = (const bf16*)m->w.final_norm_w;
    p.inv_freq = m->w.text_inv_freq;
    p.L = c.layers; p.H = c.hidden; p.I = c.inter; p.Hq = c.q_heads; p.Hkv = c.kv_heads; p.V = c.vocab;
    p.qkv_dim = (c.q_heads + 2 * c.kv_heads) * 128;
    p.eps = c.rms_eps;
    p.B = B;
    for (int b = 0; b < B; ++b) {
        if (sts[b].k_pool != sts[0].k_pool || sts[b].v_pool != sts[0].v_pool || sts[b].layer_stride != sts[0].layer_stride)
            return -1;  // all streams of a batch live in one page pool
        p.st[b].page_table = sts[b].page_table;
        p.st[b].scalars = sts[b].scalars;
    }
    p.k_pool = (bf16*)sts[0].k_pool; p.v_pool = (bf16*)sts[0].v_pool;
    p.layer_stride = sts[0].layer_stride;
    p.kv_rows_per_layer = (int)(sts[0].layer_stride / 128);
    p.h = (bf16*)(ws + L.h1); p.qkv = (bf16*)(ws + L.qkv1); p.attn = (bf16*)(ws + L.attn1); p.act = (bf16*)(ws + L.act1);
    p.logits_raw = (float*)(ws + L.logits_raw); p.logits_proc = (float*)(ws + L.logits_proc);
    p.part_o = (float*)(ws + L.mg_part_o); p.part_ml = (float*)(ws + L.mg_part_ml);
    int* cnt = (int*)(ws + L.mg_cnt);
    p.pair_cnt = cnt; p.bar = (unsigned*)(cnt + 96); p.err = cnt + 97;
    p.trace = m->mega_trace ? (unsigned long long*)(ws + L.mg_trace) : nullptr;
    p.layer_begin = 0; p.layer_end = c.layers; p.phase_mask = 31; p.do_head = 1;
    p.lookahead = m->mega_lookahead;
    p.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
    *out = p;
    return 0;
}

static int mega_launch(lcc_model* m, const lcc::MegaParams& p, cudaStream_t s) {
    const cudaError_t me = cudaMemsetAsync(p.bar, 0, sizeof(unsigned), s);  // grid barrier counter starts at 0
    if (me != cudaSuccess) {
        fprintf(stderr, "[livecc_b200] cudaMemsetAsync(barrier) failed: %s\n", cudaGetErrorString(me));
        return -20;
    }
    const long long pool_rows = (long long)p.kv_rows_per_layer * p.L;
    return lcc::decode_mega_launch(p, p.k_pool, p.v_pool, pool_rows, m->ctx->num_sms, s);
}

int lcc_decode_batch(lcc_model* m, const lcc_stream_state* states, int n_streams, int n_steps, const lcc_sampling* sp,
                     lcc_stream_t stream) {
    if (!m || !states || !sp) return -1;
    if (n_streams < 1 || n_streams > lcc::MG_MAXB) LCC_FAIL(m->ctx, -2, "lcc_decode_batch: 1..%d streams per launch", lcc::MG_MAXB);
    if (!m->mega_ok) LCC_FAIL(m->ctx, -4, "lcc_decode_batch: model geometry is not supported by the persistent decode kernel");
    WsLayout L;
    if (get_layout(m, 0, 0, &L)) LCC_FAIL(m->ctx, -3, "lcc_decode_batch: no workspace bound");
    cudaStream_t s = (cudaStream_t)stream;
    lcc::MegaParams p;
    if (mega_step_params(m, states, n_streams, L, &p)) LCC_FAIL(m->ctx, -5, "lcc_decode_batch: streams must share one page pool");
    lcc::SampleBatch sb{};
    sb.B = n_streams;
    for (int b = 0; b < n_streams; ++b) { sb.seq[b] = states[b].seq; sb.scalars[b] = states[b].scalars; }
    sb.err = p.err;
    const lcc::SampleArgs base = make_sample(m, nullptr, sp, m->ws, L, 1);
    for (int step = 0; step < n_steps; ++step) {
        STEP(mega_launch(m, p, s), "persistent decode step");
        STEP(lcc::sample_greedy_batch(base, sb, s), "batched token selection");
    }
    LCC_CHECK_LAUNCH(m->ctx, "lcc_decode_batch");
    return 0;
}

int lcc_decode_mega_debug(lcc_model* m, const lcc_stream_state* states, int n_streams, int layer_begin, int layer_end,
                          int phase_mask, int do_head, lcc_stream_t stream) {
    if (!m || !states) return -1;
    if (!m->mega_ok) LCC_FAIL(m->ctx, -4, "lcc_decode_mega_debug: geometry not supported");
    WsLayout L;
    if (get_layout(m, 0, 0, &L)) LCC_FAIL(m->ctx, -3, "lcc_decode_mega_debug: no workspace bound");
    lcc::MegaParams p;
    if (n_streams < 1 || n_streams > lcc::MG_MAXB || mega_step_params(m, states, n_streams, L, &p))
        LCC_FAIL(m->ctx, -5, "lcc_decode_mega_debug: bad streams");
    if (layer_begin < 0 || layer_end > m->cfg.layers || layer_begin > layer_end) LCC_FAIL(m->ctx, -6, "bad layer range");
    p.layer_begin = layer_begin; p.layer_end = layer_end; p.phase_mask = phase_mask; p.do_head = do_head;
    STEP(mega_launch(m, p, (cudaStream_t)stream), "persistent decode step (sub-range)");
    LCC_CHECK_LAUNCH(m->ctx, "lcc_decode_mega_debug");
    return 0;
}

int lcc_decode_steps(lcc_model* m, const lcc_stream_state* st, int n_steps, int nsplit, const lcc_sampling* sp,
                     lcc_stream_t stream) {
    if (!m || !st || !sp) return -1;
    const lcc_model_config& c = m->cfg;
    cudaStream_t s = (cudaStream_t)stream;
    if (nsplit < 1 || nsplit > kMaxSplit) LCC_FAIL(m->ctx, -2, "lcc_decode_steps: nsplit must be in [1,%d]", kMaxSplit);
    if (m->use_mega && m->mega_ok) return lcc_decode_batch(m, st, 1, n_steps, sp, stream);  // persistent kernel, 1 stream
    WsLayout L;
    if (get_layout(m, 0, 0, &L)) LCC_FAIL(m->ctx, -3, "lcc_decode_steps: no workspace bound");
    uint8_t* ws = m->ws;
    bf16* h = (bf16*)(ws + L.h1); bf16* qkv = (bf16*)(ws + L.qkv1); bf16* attn = (bf16*)(ws + L.attn1);
    bf16* act = (bf16*)(ws + L.act1);
    float* part_o = (float*)(ws + L.part_o); float* part_ml = (float*)(ws + L.part_ml);
    int* attn_cnt = (int*)(ws + L.attn_cnt);
    const bool pdl = m->use_pdl;
    const int H = c.hidden, Hq = c.q_heads, Hkv = c.kv_heads, qkv_dim = (Hq + 2 * Hkv) * 128;
    const int* fin = st->scalars + LCC_SC_FINISHED;
    // L2 prefetch window handed from each GEMV to its successor (see l2_prefetch_tail in gemv.cu)
    const size_t kPf = 32u << 20;
    const size_t qkv_bytes = (size_t)qkv_dim * H * 2, o_bytes = (size_t)H * Hq * 128 * 2;
    const size_t gu_bytes = (size_t)2 * c.inter * H * 2, down_bytes = (size_t)H * c.inter * 2;
    auto cap = [&](size_t b) { return b < kPf ? b : kPf; };
    for (int step = 0; step < n_steps; ++step) {
        for (int i = 0; i < c.layers; ++i) {
            const lcc_layer_weights& lw = m->layers[i];
            bf16* kc = (bf16*)st->k_pool + (size_t)i * st->layer_stride;
            bf16* vc = (bf16*)st->v_pool + (size_t)i * st->layer_stride;
            const void* next_qkv = (i + 1 < c.layers) ? m->layers[i + 1].qkv_w : m->w.lm_head;
            const size_t next_qkv_bytes = (i + 1 < c.layers) ? qkv_bytes : (size_t)c.vocab * H * 2;  // lm_head: capped below
            STEP(lcc::gemv_norm_bias((const bf16*)lw.qkv_w, H, h, (const bf16*)lw.ln1_w, c.rms_eps, (const bf16*)lw.qkv_b,
                                     qkv, qkv_dim, H, fin, lw.o_w, cap(o_bytes), m->ctx->num_sms, pdl, s), "decode qkv");
            if (m->fuse_attn_oproj) {
                STEP(lcc::attn_oproj_decode(qkv, kc, vc, st->page_table, LCC_PAGE_SIZE, st->scalars + LCC_SC_KV_LEN,
                                            st->scalars + LCC_SC_ROPE_POS, fin, m->w.text_inv_freq, Hq, Hkv, nsplit,
                                            part_o, part_ml, attn_cnt, attn, (const bf16*)lw.o_w, Hq * 128, H, h,
                                            attn_cnt + 32, s), "decode attention + o_proj");
            } else {
                STEP(lcc::attn_decode(qkv, kc, vc, st->page_table, LCC_PAGE_SIZE, st->scalars + LCC_SC_KV_LEN,
                                      st->scalars + LCC_SC_ROPE_POS, fin, m->w.text_inv_freq, Hq, Hkv, nsplit, part_o,
                                      part_ml, attn_cnt, attn, pdl, s), "decode attention");
                STEP(lcc::gemv_residual((const bf16*)lw.o_w, Hq * 128, attn, h, H, Hq * 128, fin, lw.gate_up_w,
                                        cap(gu_bytes), m->ctx->num_sms, pdl, s), "decode o_proj");
            }
            STEP(lcc::gemv_norm_swiglu((const bf16*)lw.gate_up_w, H, h, (const bf16*)lw.ln2_w, c.rms_eps, act,
                                       2 * c.inter, H, fin, lw.down_w, cap(down_bytes), m->ctx->num_sms, pdl, s), "decode gate_up");
            STEP(lcc::gemv_residual((const bf16*)lw.down_w, c.inter, act, h, H, c.inter, fin, next_qkv, cap(next_qkv_bytes),
                                    m->ctx->num_sms, pdl, s), "decode down_proj");
        }
        STEP(lcc::gemv_norm_logits((const bf16*)m->w.lm_head, H, h, (const bf16*)m->w.final_norm_w, c.rms_eps,
                                   (float*)(ws + L.logits_raw), (float*)(ws + L.logits_proc), c.vocab, H, fin, m->ctx->num_sms, pdl, s),
             "decode lm_head");
        STEP(lcc::sample_greedy(make_sample(m, st, sp, ws, L, 1), s), "decode token selection");
    }
    LCC_CHECK_LAUNCH(m->ctx, "lcc_decode_steps");
    return 0;
}

}  // extern "C"

namespace {
int get_layout(const lcc_model* m, int need_patches, int need_tokens, WsLayout* out) {
    if (!m->ws) return -1;
    if (need_patches > m->cap_patches || need_tokens > m->cap_tokens) return -2;
    *out = make_layout(m->cfg, m->cap_patches, m->cap_tokens);
    return out->total > m->ws_bytes ? -3 : 0;
}
}  // namespace

extern "C" int lcc_model_bind_workspace(lcc_model* m, void* ws, size_t ws_bytes, int max_patches, int max_tokens) {
    if (!m || !ws) return -1;
    if (make_layout(m->cfg, max_patches, max_tokens).total > ws_bytes)
        LCC_FAIL(m->ctx, -2, "lcc_model_bind_workspace: %zu bytes is too small", ws_bytes);
    m->ws = reinterpret_cast<uint8_t*>(ws);
    m->ws_bytes = ws_bytes;
    // split-KV arrival counters start at zero (the attention kernel re-zeroes them itself)
    if (cudaMemset(m->ws + make_layout(m->cfg, max_patches, max_tokens).attn_cnt, 0, 64 * 4) != cudaSuccess)
        LCC_FAIL(m->ctx, -3, "lcc_model_bind_workspace: cudaMemset failed");
    m->cap_patches = max_patches;
    m->cap_tokens = max_tokens;
    // tables of the persistent decode kernel: one TMA descriptor per weight matrix (box = 32 rows x 64 k, SWIZZLE_128B)
    // and the per-layer vector pointers; uploaded synchronously (bind happens outside the hot loop)
    const lcc_model_config& c = m->cfg;
    const WsLayout L = make_layout(c, max_patches, max_tokens);
    const int qkv_dim = (c.q_heads + 2 * c.kv_heads) * 128;
    m->mega_ok = false;
    if (!(qkv_dim % 32) && !(c.hidden % 256) && !(c.inter % 256) && !(c.vocab % 32) && c.q_heads / c.kv_heads <= 8) {
        std::vector<CUtensorMap> maps(4 * c.layers + 1);
        std::vector<lcc::MegaLayer> lays(c.layers);
        bool ok = true;
        for (int i = 0; i < c.layers && ok; ++i) {
            const lcc_layer_weights& lw = m->layers[i];
            ok = !lcc::mega_make_weight_tmap(&maps[4 * i + 0], lw.qkv_w, qkv_dim, c.hidden) &&
                 !lcc::mega_make_weight_tmap(&maps[4 * i + 1], lw.o_w, c.hidden, c.q_heads * 128) &&
                 !lcc::mega_make_weight_tmap(&maps[4 * i + 2], lw.gate_up_w, 2 * c.inter, c.hidden) &&
                 !lcc::mega_make_weight_tmap(&maps[4 * i + 3], lw.down_w, c.hidden, c.inter);
            lays[i] = lcc::MegaLayer{(const bf16*)lw.ln1_w, (const bf16*)lw.qkv_b, (const bf16*)lw.ln2_w};
        }
        ok = ok && !lcc::mega_make_weight_tmap(&maps[4 * c.layers], m->w.lm_head, c.vocab, c.hidden);
        if (ok &&
            cudaMemcpy(m->ws + L.mg_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(m->ws + L.mg_layers, lays.data(), lays.size() * sizeof(lcc::MegaLayer), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemset(m->ws + L.mg_cnt, 0, 128 * 4) == cudaSuccess)
            m->mega_ok = true;
    }
    return 0;
}
