// GPU frame ingest: the antialiased bicubic resize of
//   REF/livecc-utils/src/livecc_utils/video_process_patch.py:101-106,150-155
//   transforms.functional.resize(clip_u8_TCHW, [H, W], BICUBIC, antialias=True)
// reproduced bit for bit: torchvision casts uint8 -> float32, ATen's CPU upsample_bicubic2d_aa runs a width pass then a
// height pass with a float32 intermediate, torchvision clamps to [0,255], rounds half to even and casts to uint8.
// The window/weight tables are built on the host with the same mixed float/double arithmetic as ATen
// (_compute_indices_min_size_weights_aa<float>, aa_filter with a = -0.5); the rounding order of the compiled wheel
// (FMA-contracted polynomial; tap sums with 4-wide unfused groups and a fused remainder) is stated in
// oracle/resize_aa.py and pinned by tests/test_resize_cpu.py / tests/golden/resize_aa_golden.json.
//
// One kernel, both passes: a CTA owns TH output rows of one plane. It copies the source rows its vertical windows cover
// (contiguous in HBM, 16-byte loads) into shared memory, runs the width pass from shared memory into a float tile
// [rows][W] in shared memory, then the height pass, clamp, round, and writes 4 output bytes per thread. The float
// intermediate never leaves the SM; HBM traffic = source (re-read of window overlap is served by L2) + output.
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "launch.h"
#include "resize.h"

namespace lcc {

// ---- host: ATen's window / weight table ----------------------------------------------------------------------------
// x86-64 baseline has no FMA instruction, so the compiler cannot contract the float expressions below; every fused
// operation is written as fmaf (correctly rounded in glibc with or without hardware FMA).
static inline float cubic_weight(float x) {
    x = fabsf(x);
    if (x < 1.0f) {
        const float t1 = fmaf(1.5f, x, -2.5f);  // (a + 2) x - (a + 3), a = -0.5
        return fmaf(t1 * x, x, 1.0f);
    }
    if (x < 2.0f) {
        const float u1 = fmaf(-0.5f, x, 2.5f);  // a x - 5 a
        const float u2 = fmaf(u1, x, -4.0f);    // (...) x + 8 a
        return fmaf(u2, x, 2.0f);               // (...) x - 4 a
    }
    return 0.0f;
}

static inline void aa_scale(int in_size, int out_size, float& scale, float& support, float& invscale) {
    scale = (float)in_size / (float)out_size;  // area_pixel_compute_scale<float>(align_corners=false, no scale factor)
    support = scale >= 1.0f ? (float)(2.0 * (double)scale) : 2.0f;
    invscale = scale >= 1.0f ? (float)(1.0 / (double)scale) : 1.0f;
}

int resize_aa_taps(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return -1;
    float scale, support, invscale;
    aa_scale(in_size, out_size, scale, support, invscale);
    return (int)ceilf(support) * 2 + 1;
}

int resize_aa_table(int in_size, int out_size, int32_t* xmin_out, int32_t* xsize_out, float* weights) {
    const int taps = resize_aa_taps(in_size, out_size);
    if (taps < 0) return -1;
    float scale, support, invscale;
    aa_scale(in_size, out_size, scale, support, invscale);
    for (int i = 0; i < out_size; ++i) {
        const float center = (float)((double)scale * ((double)i + 0.5));
        const float lo = center - support, hi = center + support;
        int64_t xmin = (int64_t)((double)lo + 0.5);
        if (xmin < 0) xmin = 0;
        int64_t xmax = (int64_t)((double)hi + 0.5);
        if (xmax > in_size) xmax = in_size;
        int64_t xsize = xmax - xmin;
        if (xsize < 0) xsize = 0;
        if (xsize > taps) xsize = taps;
        float total = 0.0f;
        for (int j = 0; j < taps; ++j) weights[(size_t)j * out_size + i] = 0.0f;  // tap-major: [taps][out]
        for (int j = 0; j < (int)xsize; ++j) {
            const float d = (float)(j + xmin) - center;
            const float w = cubic_weight((float)(((double)d + 0.5) * (double)invscale));
            weights[(size_t)j * out_size + i] = w;
            total += w;
        }
        if (total != 0.0f)
            for (int j = 0; j < (int)xsize; ++j) weights[(size_t)j * out_size + i] /= total;
        xmin_out[i] = (int32_t)xmin;
        xsize_out[i] = (int32_t)xsize;
    }
    return 0;
}

// ---- device ----------------------------------------------------------------------------------------------------------
struct ResizeParams {
    const uint8_t* src;
    uint8_t* dst;
    int h, w, H, W;
    const int32_t* xmin_w; const int32_t* n_w; const float* wt_w;  // [W], [W], [taps_w][W]
    const int32_t* xmin_h; const int32_t* n_h; const float* wt_h;  // [H], [H], [taps_h][H]
    int th;        // output rows per CTA
    int max_rows;  // source rows a CTA may stage
};

// t = s0*w0; then 4*floor((n-1)/4) separate multiply/add steps; the remaining (n-1) mod 4 steps fused (oracle/resize_aa.py)
template <typename LoadS, typename LoadW>
__device__ __forceinline__ float tap_sum(int n, LoadS s, LoadW w) {
    if (n <= 0) return 0.f;
    float acc = __fmul_rn(s(0), w(0));
    const int unfused = ((n - 1) >> 2) << 2;
    int j = 1;
    for (; j <= unfused; ++j) acc = __fadd_rn(acc, __fmul_rn(s(j), w(j)));
    for (; j < n; ++j) acc = __fmaf_rn(s(j), w(j), acc);
    return acc;
}

__device__ __forceinline__ uint32_t to_u8(float v) {  // clamp(0, 255), round half to even (torch.round), cast
    return (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f));
}

__global__ void __launch_bounds__(256) resize_bicubic_aa_u8_kernel(const ResizeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int plane = blockIdx.y;
    const int y0 = blockIdx.x * p.th, y1 = min(y0 + p.th, p.H);
    const int r0 = p.xmin_h[y0];
    int r1 = r0;
    for (int y = y0; y < y1; ++y) r1 = max(r1, p.xmin_h[y] + p.n_h[y]);
    const int nr = min(r1 - r0, p.max_rows);
    float* tile = reinterpret_cast<float*>(smem);                                  // [max_rows][W]
    uint8_t* rows = smem + (((size_t)p.max_rows * p.W * sizeof(float) + 15) & ~(size_t)15);  // nr*w source bytes (+ <= 15 of skew)

    // ---- stage the source rows: one contiguous span of the plane; keep the global address's 16-byte phase in shared memory
    const uint8_t* g = p.src + ((size_t)plane * p.h + r0) * p.w;
    const size_t nbytes = (size_t)nr * p.w;
    const int skew = (int)(reinterpret_cast<uintptr_t>(g) & 15);
    uint8_t* srow = rows + skew;
    const size_t head = min(nbytes, (size_t)((16 - skew) & 15));
    for (size_t i = threadIdx.x; i < head; i += blockDim.x) srow[i] = g[i];
    const size_t nvec = (nbytes - head) >> 4;
    const uint4* gv = reinterpret_cast<const uint4*>(g + head);
    uint4* sv = reinterpret_cast<uint4*>(srow + head);
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) sv[i] = __ldg(gv + i);
    for (size_t i = head + (nvec << 4) + threadIdx.x; i < nbytes; i += blockDim.x) srow[i] = g[i];
    __syncthreads();

    // ---- width pass: tile[r][X] = sum_j row_r[xmin_w[X] + j] * wt_w[j][X]
    const int W = p.W;
    for (int idx = threadIdx.x; idx < nr * W; idx += blockDim.x) {
        const int r = idx / W, X = idx - r * W;
        const uint8_t* s = srow + (size_t)r * p.w + p.xmin_w[X];
        const float* wt = p.wt_w + X;
        tile[idx] = tap_sum(p.n_w[X], [&](int j) { return (float)s[j]; }, [&](int j) { return __ldg(wt + (size_t)j * W); });
    }
    __syncthreads();

    // ---- height pass + clamp/round/cast
    uint8_t* out = p.dst + (size_t)plane * p.H * W;
    if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        const int W4 = W >> 2;
        for (int idx = threadIdx.x; idx < (y1 - y0) * W4; idx += blockDim.x) {
            const int yy = idx / W4, X = (idx - yy * W4) << 2;
            const int y = y0 + yy, n = p.n_h[y];
            const float* t = tile + (size_t)(p.xmin_h[y] - r0) * W + X;
            const float* wt = p.wt_h + y;
            float a[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                a[c] = tap_sum(n, [&](int j) { return t[(size_t)j * W + c]; }, [&](int j) { return __ldg(wt + (size_t)j * p.H); });
            *reinterpret_cast<uint32_t*>(out + (size_t)y * W + X) =
                to_u8(a[0]) | (to_u8(a[1]) << 8) | (to_u8(a[2]) << 16) | (to_u8(a[3]) << 24);
        }
    } else {
        for (int idx = threadIdx.x; idx < (y1 - y0) * W; idx += blockDim.x) {
            const int yy = idx / W, X = idx - yy * W;
            const int y = y0 + yy;
            const float* t = tile + (size_t)(p.xmin_h[y] - r0) * W + X;
            const float* wt = p.wt_h + y;
            out[(size_t)y * W + X] = (uint8_t)to_u8(
                tap_sum(p.n_h[y], [&](int j) { return t[(size_t)j * W]; }, [&](int j) { return __ldg(wt + (size_t)j * p.H); }));
        }
    }
}

// ---- plan ------------------------------------------------------------------------------------------------------------
struct ResizePlan {
    int h, w, H, W;
    int taps_w, taps_h;
    int th, max_rows;
    size_t smem;
    void* dev = nullptr;  // one allocation: xmin_w, n_w, xmin_h, n_h, wt_w, wt_h
    const int32_t *xmin_w, *n_w, *xmin_h, *n_h;
    const float *wt_w, *wt_h;
};

static SmemAttrOnce g_resize_smem;

ResizePlan* resize_plan_create(int h, int w, int H, int W, int th_override) {
    if (h <= 0 || w <= 0 || H <= 0 || W <= 0 || w > (1 << 15) || h > (1 << 15) || W > (1 << 14) || H > (1 << 14)) return nullptr;
    const int tw = resize_aa_taps(w, W), thh = resize_aa_taps(h, H);
    std::vector<int32_t> xw(W), nw(W), xh(H), nh(H);
    std::vector<float> ww((size_t)tw * W), wh((size_t)thh * H);
    if (resize_aa_table(w, W, xw.data(), nw.data(), ww.data()) || resize_aa_table(h, H, xh.data(), nh.data(), wh.data()))
        return nullptr;
    // rows per CTA: the largest of 16/8/4/2/1 whose tile fits ~100 KB (two CTAs per SM), else whatever fits 227 KB
    auto rows_needed = [&](int th) {
        int m = 0;
        for (int y0 = 0; y0 < H; y0 += th) {
            int r1 = xh[y0];
            for (int y = y0; y < std::min(y0 + th, H); ++y) r1 = std::max(r1, xh[y] + nh[y]);
            m = std::max(m, r1 - xh[y0]);
        }
        return m;
    };
    auto smem_bytes = [&](int rows) { return (((size_t)rows * W * 4 + 15) & ~(size_t)15) + (size_t)rows * w + 32; };
    int best = 0;
    const int cands[5] = {16, 8, 4, 2, 1};
    if (th_override > 0) {
        if (smem_bytes(rows_needed(th_override)) <= 227 * 1024) best = th_override;
    } else {
        for (int c : cands)
            if (smem_bytes(rows_needed(c)) <= 100 * 1024) { best = c; break; }
        if (!best)
            for (int c : cands)
                if (smem_bytes(rows_needed(c)) <= 227 * 1024) { best = c; break; }
    }
    if (!best) return nullptr;  // a single output row's window does not fit in shared memory
    ResizePlan* pl = new ResizePlan();
    pl->h = h; pl->w = w; pl->H = H; pl->W = W;
    pl->taps_w = tw; pl->taps_h = thh;
    pl->th = best;
    pl->max_rows = rows_needed(best);
    pl->smem = (smem_bytes(pl->max_rows) + 15) & ~(size_t)15;
    const size_t ints = (size_t)2 * W + 2 * H;
    const size_t bytes = ints * 4 + ((size_t)tw * W + (size_t)thh * H) * 4;
    if (cudaMalloc(&pl->dev, bytes) != cudaSuccess) { delete pl; return nullptr; }
    std::vector<uint8_t> host(bytes);
    int32_t* ip = reinterpret_cast<int32_t*>(host.data());
    memcpy(ip, xw.data(), (size_t)W * 4); memcpy(ip + W, nw.data(), (size_t)W * 4);
    memcpy(ip + 2 * W, xh.data(), (size_t)H * 4); memcpy(ip + 2 * W + H, nh.data(), (size_t)H * 4);
    float* fp = reinterpret_cast<float*>(ip + ints);
    memcpy(fp, ww.data(), ww.size() * 4); memcpy(fp + ww.size(), wh.data(), wh.size() * 4);
    if (cudaMemcpy(pl->dev, host.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(pl->dev); delete pl; return nullptr; }
    const int32_t* dip = reinterpret_cast<const int32_t*>(pl->dev);
    pl->xmin_w = dip; pl->n_w = dip + W; pl->xmin_h = dip + 2 * W; pl->n_h = dip + 2 * W + H;
    const float* dfp = reinterpret_cast<const float*>(dip + ints);
    pl->wt_w = dfp; pl->wt_h = dfp + ww.size();
    return pl;
}

void resize_plan_destroy(ResizePlan* pl) {
    if (!pl) return;
    if (pl->dev) cudaFree(pl->dev);
    delete pl;
}

void resize_plan_info(const ResizePlan* pl, int* th, int* max_rows, int64_t* smem) {
    if (th) *th = pl->th;
    if (max_rows) *max_rows = pl->max_rows;
    if (smem) *smem = (int64_t)pl->smem;
}

int resize_bicubic_aa_u8(const ResizePlan* pl, const uint8_t* src, int planes, uint8_t* dst, cudaStream_t s) {
    if (!pl || planes <= 0 || planes > 65535) return -1;
    if (ensure_dyn_smem(g_resize_smem, resize_bicubic_aa_u8_kernel, 227 * 1024)) return -2;
    ResizeParams p;
    p.src = src; p.dst = dst;
    p.h = pl->h; p.w = pl->w; p.H = pl->H; p.W = pl->W;
    p.xmin_w = pl->xmin_w; p.n_w = pl->n_w; p.wt_w = pl->wt_w;
    p.xmin_h = pl->xmin_h; p.n_h = pl->n_h; p.wt_h = pl->wt_h;
    p.th = pl->th; p.max_rows = pl->max_rows;
    dim3 grid((pl->H + pl->th - 1) / pl->th, planes);
    count_launch();
    resize_bicubic_aa_u8_kernel<<<grid, 256, pl->smem, s>>>(p);
    return 0;
}

}  // namespace lcc
