// HBM-bound row kernels of the LiveCC path: casts, LayerNorm, RMSNorm, ViT 2-D rotary, M-RoPE +
// paged-KV append, embedding gather/scatter. One warp per row wherever a row fits a warp's registers;
// 16-byte vectorised, coalesced accesses; grids sized from the row count.
#include "common.cuh"
#include "launch.h"
#include "ops.h"

namespace lcc {

// ------------------------------------------------------------------------------------------
// f32 -> bf16 (pixel_values_videos.to(bf16), mq2vl.py:309)
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (; i + 8 <= n; i += stride) {
        float4 a = *reinterpret_cast<const float4*>(in + i);
        float4 b = *reinterpret_cast<const float4*>(in + i + 4);
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
        *reinterpret_cast<uint4*>(out + i) = o;
    }
    // tail (n % 8): handled by the first threads
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        int64_t j = (n & ~int64_t(7)) + threadIdx.x;
        out[j] = f2bf(in[j]);
    }
}

int cast_f32_bf16(const float* in, bf16* out, int64_t n, int num_sms, cudaStream_t s) {
    if (n <= 0) return 0;
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
    { lcc::count_launch(); cast_f32_bf16_kernel<<<(int)blocks, 256, 0, s>>>(in, out, n); }
    return 0;
}

// ------------------------------------------------------------------------------------------
// GPU frame ingest: uint8 frames [T,3,H,W] -> bf16 patch rows [N, 3*2*14*14] in one pass,
// replacing the host-side rescale/normalize/patchify of Qwen2VLVideoProcessor._preprocess
// (video_processing_qwen2_vl.py:240-272; fused mean/std: image_processing_backends.py:301-304,327) plus the
// f32 H2D copy and the bf16 cast (mq2vl.py:309). Bit-identical to cast_bf16(host patchify): the value is
// bf16( (float(u8) - mean255[c]) / std255[c] ) with IEEE fp32 subtract and divide; odd T repeats the last frame.
// Row order (t, h/2, w/2, 2, 2), column order (c, tp, 14, 14).
// ------------------------------------------------------------------------------------------
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ frames, int T, int H, int W,
                                   bf16* __restrict__ out, float m0, float m1, float m2, float s0, float s1,
                                   float s2) {
    constexpr int P = 14, TP = 2, MS = 2, COLS = 3 * TP * P * P;
    const int gh = H / P, gw = W / P;
    const int grid_t = (T + TP - 1) / TP;
    const int64_t total = (int64_t)grid_t * gh * gw * (COLS / 2);  // two adjacent px per thread
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = (int)(idx % (COLS / 2)) * 2;
    const int row = (int)(idx / (COLS / 2));
    // row -> (t, hb, wb, hi, wi)
    const int in_win = row % (MS * MS);
    const int win = row / (MS * MS);
    const int wbn = gw / MS, hbn = gh / MS;
    const int wb = win % wbn, hb = (win / wbn) % hbn, t = win / (wbn * hbn);
    const int ph = hb * MS + in_win / MS, pw = wb * MS + in_win % MS;
    // col -> (c, tp, py, px)
    const int px = col % P, py = (col / P) % P, tp = (col / (P * P)) % TP, c = col / (P * P * TP);
    const int tf = min(t * TP + tp, T - 1);
    const uint8_t* src = frames + (((size_t)tf * 3 + c) * H + (size_t)ph * P + py) * W + (size_t)pw * P + px;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float a = __fdiv_rn(__fsub_rn((float)src[0], mean), sd);
    const float b = __fdiv_rn(__fsub_rn((float)src[1], mean), sd);  // px is even, px+1 < 14: same patch row
    *reinterpret_cast<uint32_t*>(out + (size_t)row * COLS + col) = pack_bf16x2(a, b);
}

int patchify_u8(const uint8_t* frames, int T, int H, int W, bf16* out, const float* mean255, const float* std255,
                cudaStream_t s) {
    if (T <= 0 || H % 28 || W % 28) return -1;
    const int64_t total = (int64_t)((T + 1) / 2) * (H / 14) * (W / 14) * (1176 / 2);
    lcc::count_launch();
    patchify_u8_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(frames, T, H, W, out, mean255[0], mean255[1],
                                                                  mean255[2], std255[0], std255[1], std255[2]);
    return 0;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over rows (nn.LayerNorm eps=1e-6, mq2vl.py:464-465,317): fp32 two-pass statistics,
// y = bf16((x - mean) * rstd * w + b). One warp per row, dim % 8 == 0.
// ------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                                 const bf16* __restrict__ b, bf16* __restrict__ y, int ldy, int rows,
                                 int dim, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const bf16* xr = x + (size_t)warp * ldx;
    float sum = 0.f;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = unpack_bf16x2(uw[j]); sum += f.x + f.y; }
    }
    const float mean = warp_sum(sum) / (float)dim;
    float sq = 0.f;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(uw[j]);
            float d0 = f.x - mean, d1 = f.y - mean;
            sq += d0 * d0 + d1 * d1;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)dim + eps);
    bf16* yr = y + (size_t)warp * ldy;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        uint4 wv = *reinterpret_cast<const uint4*>(w + c);
        uint4 bv = *reinterpret_cast<const uint4*>(b + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(uw[j]), g = unpack_bf16x2(ww[j]), h = unpack_bf16x2(bw[j]);
            o[j] = pack_bf16x2((f.x - mean) * rstd * g.x + h.x, (f.y - mean) * rstd * g.y + h.y);
        }
        *reinterpret_cast<uint4*>(yr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int layernorm(const bf16* x, int ldx, const bf16* w, const bf16* b, bf16* y, int ldy, int rows, int dim,
              float eps, cudaStream_t s) {
    if (rows <= 0) return 0;
    if (dim % 8 || ldx % 8 || ldy % 8) return -1;
    const int warps_per_block = 8;
    lcc::count_launch();
    layernorm_kernel<<<(rows + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(
        x, ldx, w, b, y, ldy, rows, dim, eps);
    return 0;
}

// ------------------------------------------------------------------------------------------
// RMSNorm over rows (Qwen2VLRMSNorm, mq2vl.py:126-131): y = w * bf16(x * rsqrt(mean(x^2) + eps)).
// ------------------------------------------------------------------------------------------
__global__ void rmsnorm_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                               bf16* __restrict__ y, int ldy, int rows, int dim, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const bf16* xr = x + (size_t)warp * ldx;
    float sq = 0.f;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = unpack_bf16x2(uw[j]); sq += f.x * f.x + f.y * f.y; }
    }
    const float rs = rsqrtf(warp_sum(sq) / (float)dim + eps);
    bf16* yr = y + (size_t)warp * ldy;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        uint4 wv = *reinterpret_cast<const uint4*>(w + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(uw[j]), g = unpack_bf16x2(ww[j]);
            o[j] = pack_bf16x2(g.x * rbf(f.x * rs), g.y * rbf(f.y * rs));
        }
        *reinterpret_cast<uint4*>(yr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int rmsnorm(const bf16* x, int ldx, const bf16* w, bf16* y, int ldy, int rows, int dim, float eps,
            cudaStream_t s) {
    if (rows <= 0) return 0;
    if (dim % 8 || ldx % 8 || ldy % 8) return -1;
    const int warps_per_block = 8;
    lcc::count_launch();
    rmsnorm_kernel<<<(rows + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(
        x, ldx, w, y, ldy, rows, dim, eps);
    return 0;
}

// ------------------------------------------------------------------------------------------
// ViT 2-D rotary (mq2vl.py:725-752 table, :257-268 application).
// Table: cos/sin [N, head_dim/2] fp32 where column j < hd/4 uses the patch's h id and
// hd/4 <= j < hd/2 its w id, angle = id * inv_freq[j mod hd/4]; (h,w) ids follow the merge-window
// patch order of the processor. Built once per grid shape.
// ------------------------------------------------------------------------------------------
__global__ void vit_rope_table_kernel(float* __restrict__ cos_t, float* __restrict__ sin_t, int t, int h,
                                      int w, int merge, int half /* = head_dim/2 */,
                                      const float* __restrict__ inv_freq /* [half/2] */) {
    const int per_frame = h * w;
    const int n = t * per_frame;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * half) return;
    const int p = idx / half, j = idx % half;
    int r = p % per_frame;
    // merge-window order: (h/m, w/m, m, m)
    const int wm = w / merge;
    const int win = r / (merge * merge), in = r % (merge * merge);
    const int hy = (win / wm) * merge + in / merge;
    const int wx = (win % wm) * merge + in % merge;
    const int q = half / 2;
    const float pos = (float)(j < q ? hy : wx);
    const float ang = __fmul_rn(pos, inv_freq[j % q]);
    cos_t[idx] = cosf(ang);
    sin_t[idx] = sinf(ang);
}

int vit_rope_table(float* cos_t, float* sin_t, int t, int h, int w, int merge, int head_dim,
                   const float* inv_freq, cudaStream_t s) {
    const int half = head_dim / 2;
    const int total = t * h * w * half;
    if (total <= 0) return 0;
    { lcc::count_launch(); vit_rope_table_kernel<<<(total + 255) / 256, 256, 0, s>>>(cos_t, sin_t, t, h, w, merge, half, inv_freq); }
    return 0;
}

// In-place rotation of the q and k thirds of qkv [N, 3*heads*hd] (fp32 math, one rounding to bf16).
// out[i] = x[i]*cos[i] + rot(x)[i]*sin[i], rot(x) = cat(-x[hd/2:], x[:hd/2]); cos[i] = table[i mod hd/2].
__global__ void vit_rope_apply_kernel(bf16* __restrict__ qkv, int ld, const float* __restrict__ cos_t,
                                      const float* __restrict__ sin_t, int N, int heads, int hd) {
    // one thread = 8 consecutive rotation pairs (x[j..j+8), x[j+half..j+half+8)): two 16-byte loads/stores of
    // bf16 and four float4 loads of cos/sin; requires half % 8 == 0 (hd = 80 -> half = 40)
    const int half = hd / 2;
    const int chunks = half / 8;
    const int per_row = 2 * heads * chunks;  // q and k
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * per_row) return;
    const int row = (int)(idx / per_row);
    const int r = (int)(idx % per_row);
    const int which = r / (heads * chunks);  // 0 = q, 1 = k
    const int hh = (r / chunks) % heads;
    const int j0 = (r % chunks) * 8;
    bf16* base = qkv + (size_t)row * ld + (size_t)which * heads * hd + (size_t)hh * hd + j0;
    const uint4 u1 = *reinterpret_cast<const uint4*>(base);
    const uint4 u2 = *reinterpret_cast<const uint4*>(base + half);
    const float4* cp = reinterpret_cast<const float4*>(cos_t + (size_t)row * half + j0);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + (size_t)row * half + j0);
    const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const uint32_t w1[4] = {u1.x, u1.y, u1.z, u1.w}, w2[4] = {u2.x, u2.y, u2.z, u2.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 a = unpack_bf16x2(w1[k]), b = unpack_bf16x2(w2[k]);
        // (q * cos) + (rotate_half(q) * sin) with separately rounded products (no FMA contraction)
        const float p0 = __fadd_rn(__fmul_rn(a.x, c[2 * k]), __fmul_rn(-b.x, sn[2 * k]));
        const float p1 = __fadd_rn(__fmul_rn(a.y, c[2 * k + 1]), __fmul_rn(-b.y, sn[2 * k + 1]));
        const float q0 = __fadd_rn(__fmul_rn(b.x, c[2 * k]), __fmul_rn(a.x, sn[2 * k]));
        const float q1 = __fadd_rn(__fmul_rn(b.y, c[2 * k + 1]), __fmul_rn(a.y, sn[2 * k + 1]));
        o1[k] = pack_bf16x2(p0, p1);
        o2[k] = pack_bf16x2(q0, q1);
    }
    *reinterpret_cast<uint4*>(base) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(base + half) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
}

int vit_rope_apply(bf16* qkv, int ld, const float* cos_t, const float* sin_t, int N, int heads, int hd,
                   cudaStream_t s) {
    if ((hd / 2) % 8 || ld % 8) return -1;
    const int64_t total = (int64_t)N * 2 * heads * (hd / 16);
    if (total <= 0) return 0;
    { lcc::count_launch(); vit_rope_apply_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(qkv, ld, cos_t, sin_t, N, heads, hd); }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Embedding gather + scatter of video features (mq2vl.py:1255-1272).
// out[s] = is_video(ids[s]) ? video_embeds[rank of s among video tokens] : table[ids[s]].
// Kernel 1 (single CTA) computes the exclusive rank by a block scan; kernel 2 copies rows.
// ------------------------------------------------------------------------------------------
__global__ void video_rank_kernel(const int64_t* __restrict__ ids, int S, int64_t video_id,
                                  int* __restrict__ rank, int* __restrict__ total) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < S; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int flag = (i < S && ids[i] == video_id) ? 1 : 0;
        int v = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int t = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += n;
            }
            warp_tot[lane] = t;  // inclusive scan of warp totals
        }
        __syncthreads();
        const int prefix = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + v - flag;
        if (i < S) rank[i] = flag ? prefix : -1;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void embed_gather_kernel(const int64_t* __restrict__ ids, const int* __restrict__ rank,
                                    const bf16* __restrict__ table, const bf16* __restrict__ video,
                                    bf16* __restrict__ out, int S, int H, int64_t vocab, int n_video_rows) {
    const int row = blockIdx.x;
    if (row >= S) return;
    const int rk = rank[row];
    int64_t id = ids[row];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    // A placeholder beyond the last feature row (more <|video_pad|> ids than ViT rows) keeps its text embedding:
    // nothing is read past the end of `video`; the host raises on the count mismatch (mq2vl.py:1169-1175).
    const bf16* src = (rk >= 0 && rk < n_video_rows) ? video + (size_t)rk * H : table + (size_t)id * H;
    bf16* dst = out + (size_t)row * H;
    for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
}

int embed_gather(const int64_t* ids, const bf16* table, const bf16* video, int n_video_rows, int64_t video_id, bf16* out,
                 int* rank_ws, int* total_video, int S, int H, int64_t vocab, cudaStream_t s) {
    if (S <= 0) return 0;
    if (H % 8) return -1;
    { lcc::count_launch(); video_rank_kernel<<<1, 1024, 0, s>>>(ids, S, video ? video_id : (int64_t)-1, rank_ws, total_video); }
    lcc::count_launch();
    embed_gather_kernel<<<S, 128, 0, s>>>(ids, rank_ws, table, video, out, S, H, vocab, video ? n_video_rows : 0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// M-RoPE on q,k of the fused qkv rows + append of (rotated k, v) to the paged KV cache
// (mq2vl.py:188-201 cos/sin in fp32 rounded to bf16, :244-254 section select + rotate-half,
// cache_utils.py:102-121 append). Prefill form: S rows.
//   qkv  [S, (Hq + 2*Hkv) * 128] : q is rotated in place, k/v are read.
//   pos3 [3, S] int32            : temporal / height / width position of every new token.
//   cache layer base: K at kc, V at vc, each [num_pages, Hkv, P, 128]; token (kv_start + s) lives in
//   page page_table[(kv_start+s)/P], slot (kv_start+s)%P.
// q*cos + rotate_half(q)*sin is evaluated the way torch does on bf16 tensors: each product and the
// sum are rounded to bf16.
// ------------------------------------------------------------------------------------------
struct MropeSections { int s0, s1; };  // boundaries in [0, 64): j < s0 -> T, j < s0+s1 -> H, else W

__device__ __forceinline__ void mrope_cos_sin(int j /*0..63*/, int pt, int ph, int pw, MropeSections sec,
                                              const float* inv_freq, float& c, float& s) {
    const int p = j < sec.s0 ? pt : (j < sec.s0 + sec.s1 ? ph : pw);
    const float ang = __fmul_rn((float)p, inv_freq[j]);
    c = rbf(cosf(ang));
    s = rbf(sinf(ang));
}

__global__ void mrope_kv_write_kernel(bf16* __restrict__ qkv, int ld, const int* __restrict__ pos3, int S,
                                      const float* __restrict__ inv_freq, MropeSections sec, int Hq, int Hkv,
                                      bf16* __restrict__ kc, bf16* __restrict__ vc,
                                      const int* __restrict__ page_table, int page_size, int kv_start) {
    // one warp per (token, head) over q heads, k heads and v heads; lane handles dims {2l,2l+1} and +64
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int heads_total = Hq + 2 * Hkv;
    if (gw >= S * heads_total) return;
    const int s_idx = gw / heads_total, hh = gw % heads_total;
    bf16* row = qkv + (size_t)s_idx * ld + (size_t)hh * 128;
    const int tok = kv_start + s_idx;
    const int page = page_table[tok / page_size], slot = tok % page_size;
    if (hh >= Hq + Hkv) {  // V: plain copy into the cache
        const int kvh = hh - Hq - Hkv;
        bf16* dst = vc + (((size_t)page * Hkv + kvh) * page_size + slot) * 128;
        if (lane < 16) *reinterpret_cast<uint4*>(dst + lane * 8) = *reinterpret_cast<const uint4*>(row + lane * 8);
        // The V slots of the last page behind the newest token are zeroed: the tensor-core prefill kernel multiplies
        // them by P = 0 (masked keys), which must not meet NaN/Inf bit patterns of a recycled page.
        if (s_idx == S - 1)
            for (int i = lane; i < (page_size - 1 - slot) * 16; i += 32)
                *reinterpret_cast<uint4*>(dst + 128 + i * 8) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int pt = pos3[s_idx], ph = pos3[S + s_idx], pw = pos3[2 * S + s_idx];
    // lane owns rotation pairs j = 2*lane, 2*lane+1  (x[j], x[j+64])
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(row + 2 * lane);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(row + 64 + 2 * lane);
    const float2 x1 = unpack_bf16x2(lo), x2 = unpack_bf16x2(hi);
    float c0, s0, c1, s1;
    mrope_cos_sin(2 * lane, pt, ph, pw, sec, inv_freq, c0, s0);
    mrope_cos_sin(2 * lane + 1, pt, ph, pw, sec, inv_freq, c1, s1);
    const float o1a = rbf(rbf(x1.x * c0) + rbf(-x2.x * s0));
    const float o1b = rbf(rbf(x1.y * c1) + rbf(-x2.y * s1));
    const float o2a = rbf(rbf(x2.x * c0) + rbf(x1.x * s0));
    const float o2b = rbf(rbf(x2.y * c1) + rbf(x1.y * s1));
    const uint32_t olo = pack_bf16x2(o1a, o1b), ohi = pack_bf16x2(o2a, o2b);
    if (hh < Hq) {
        *reinterpret_cast<uint32_t*>(row + 2 * lane) = olo;
        *reinterpret_cast<uint32_t*>(row + 64 + 2 * lane) = ohi;
    } else {
        const int kvh = hh - Hq;
        bf16* dst = kc + (((size_t)page * Hkv + kvh) * page_size + slot) * 128;
        *reinterpret_cast<uint32_t*>(dst + 2 * lane) = olo;
        *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = ohi;
    }
}

int mrope_kv_write(bf16* qkv, int ld, const int* pos3, int S, const float* inv_freq, int sec_t, int sec_h,
                   int Hq, int Hkv, bf16* kc, bf16* vc, const int* page_table, int page_size, int kv_start,
                   cudaStream_t s) {
    if (S <= 0) return 0;
    const int64_t warps = (int64_t)S * (Hq + 2 * Hkv);
    const int wpb = 8;
    MropeSections sec{sec_t, sec_h};
    lcc::count_launch();
    mrope_kv_write_kernel<<<(int)((warps + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        qkv, ld, pos3, S, inv_freq, sec, Hq, Hkv, kc, vc, page_table, page_size, kv_start);
    return 0;
}

}  // namespace lcc
