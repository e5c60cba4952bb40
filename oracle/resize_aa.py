"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement (numpy, float32 with explicit rounding points) of the frame resize on LiveCC's ingest path:

    REF/livecc-utils/src/livecc_utils/video_process_patch.py:101-106 and :150-155
        transforms.functional.resize(clip_u8_TCHW, [H, W], interpolation=BICUBIC, antialias=True)

The arithmetic lives in third-party code that is not under /root/reference:
  * torchvision 0.26.0 `transforms/_functional_tensor.py::resize`: uint8 -> float32, `interpolate(mode="bicubic",
    align_corners=False, antialias=True)`, clamp(0, 255), round (half to even), cast to uint8;
  * ATen (torch 2.11.0) `native/cpu/UpSampleKernel.cpp`: `separable_upsample_generic_Nd_kernel_impl` — one pass per axis,
    LAST axis first (horizontal, then vertical), float32 intermediate; per output index the window and the weights come
    from `HelperInterpBase::_compute_indices_min_size_weights_aa` with `HelperInterpCubic::aa_filter` (Keys cubic,
    a = -0.5, the Pillow filter) evaluated in float32.

Rounding points that are NOT visible in the C++ source but are fixed by the compiled wheel (x86-64 kernels built with
FMA contraction), established by fitting this restatement to the wheel bit for bit (tests/test_resize_cpu.py pins it on
live torchvision and tests/golden/resize_aa_golden.json holds the wheel's outputs):
  * the cubic polynomials are evaluated with fused multiply-adds (`cubic_weight` below);
  * the tap accumulation `t = s0*w0; for j in 1..n-1: t += s_j*w_j` runs its first 4*floor((n-1)/4) iterations as
    separate multiply and add (a 4-wide unrolled body whose products come from a vector multiply) and the remaining
    (n-1) mod 4 iterations as fused multiply-adds (the scalar epilogue).
Domain: every size except "output width 1 with a height change" (there torchvision's result is not the separable filter: ATen's
[h, 1] intermediate is stride-ambiguous; the native plan refuses that case, tests/test_resize_cpu.py documents it). The hot path only
produces multiples of 28.
Parity status: pinned against torchvision 0.26.0 + torch 2.11.0 (CPU capability AVX512) on every size in the tests; a
differently compiled ATen may order these roundings differently (the uint8 results then differ on ~1e-4 of the pixels
by one level — the same spread as between two builds of the reference's own dependency).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

f32 = np.float32
f64 = np.float64


def _fma(a, b, c):
    """float32 fused multiply-add: the double product of two floats is exact, one rounding at the end (the double sum can
    round too; for |terms| < 2^11 with 24-bit inputs the double result is exact, so this IS fmaf here)."""
    return (np.asarray(a, f32).astype(f64) * np.asarray(b, f32).astype(f64) + np.asarray(c, f32).astype(f64)).astype(f32)


def cubic_weight(x: np.ndarray) -> np.ndarray:
    """HelperInterpCubic::aa_filter<float> (UpSampleKernel.cpp; Pillow's bicubic filter, a = -0.5), FMA-contracted:
    |x| < 1: ((a+2)|x| - (a+3)) |x|^2 + 1 ;  1 <= |x| < 2: ((a|x| - 5a)|x| + 8a)|x| - 4a ; else 0."""
    x = np.abs(np.asarray(x, f32))
    t1 = _fma(f32(1.5), x, f32(-2.5))
    near = _fma((t1 * x).astype(f32), x, f32(1.0))
    u1 = _fma(f32(-0.5), x, f32(2.5))
    u2 = _fma(u1, x, f32(-4.0))
    far = _fma(u2, x, f32(2.0))
    return np.where(x < 1, near, np.where(x < 2, far, f32(0))).astype(f32)


def aa_taps(in_size: int, out_size: int) -> int:
    """max_interp_size of _compute_index_ranges_weights: ceil(support) * 2 + 1."""
    scale = f32(f32(in_size) / f32(out_size))
    support = f32(f64(2.0) * f64(scale)) if scale >= 1 else f32(2.0)
    return int(np.ceil(support)) * 2 + 1


def aa_table(in_size: int, out_size: int) -> List[Tuple[int, np.ndarray]]:
    """Per output index: (first input index, float32 weights) — _compute_indices_min_size_weights_aa<float>.
    Mixed precision as C++ evaluates it: `scale * (i + 0.5)` and `(… + 0.5) * invscale` in double (0.5 is a double
    literal), `center - support` and `j + xmin - center` in float."""
    scale = f32(f32(in_size) / f32(out_size))  # area_pixel_compute_scale<float>, align_corners=False, no scale given
    support = f32(f64(2.0) * f64(scale)) if scale >= 1 else f32(2.0)
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1 else f32(1.0)
    max_n = int(np.ceil(support)) * 2 + 1
    table = []
    for i in range(out_size):
        center = f32(f64(scale) * (f64(i) + 0.5))
        xmin = max(int(f64(f32(center - support)) + 0.5), 0)
        xsize = min(int(f64(f32(center + support)) + 0.5), in_size) - xmin
        xsize = min(max(xsize, 0), max_n)
        j = np.arange(xsize)
        x = ((((j + xmin).astype(f32) - center).astype(f32).astype(f64) + 0.5) * f64(invscale)).astype(f32)
        w = cubic_weight(x)
        total = f32(0)
        for v in w:
            total = f32(total + v)
        if total != 0:
            w = (w / total).astype(f32)
        table.append((xmin, w))
    return table


def _pass(src: np.ndarray, table, axis: int) -> np.ndarray:
    """One separable pass along `axis` (interpolate_aa_single_dim[_zero_strides]) with the wheel's rounding order."""
    src = np.moveaxis(src, axis, -1)
    out = np.zeros(src.shape[:-1] + (len(table),), dtype=f32)
    for i, (xmin, ws) in enumerate(table):
        n = len(ws)
        if n == 0:
            continue
        acc = (src[..., xmin] * ws[0]).astype(f32)
        n_unfused = ((n - 1) // 4) * 4
        for j in range(1, n):
            if j <= n_unfused:
                acc = (acc + (src[..., xmin + j] * ws[j]).astype(f32)).astype(f32)
            else:
                acc = _fma(src[..., xmin + j], ws[j], acc)
        out[..., i] = acc
    return np.moveaxis(out, -1, axis)


def resize_bicubic_aa_f32(x: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """interpolate(x_f32[..., h, w], size, mode='bicubic', align_corners=False, antialias=True): width pass, then height."""
    x = np.asarray(x, f32)
    h, w = x.shape[-2:]
    y = _pass(x, aa_table(w, size[1]), x.ndim - 1)
    return _pass(y, aa_table(h, size[0]), x.ndim - 2)


def resize_bicubic_aa_u8(clip: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """torchvision F.resize on a uint8 [..., h, w] array: identity when the size already matches
    (_functional_tensor.py / functional.py early return), else float pass, clamp, round half to even, uint8."""
    clip = np.asarray(clip)
    assert clip.dtype == np.uint8
    if tuple(clip.shape[-2:]) == tuple(size):
        return clip
    y = resize_bicubic_aa_f32(clip.astype(f32), size)
    return np.rint(np.clip(y, 0, 255)).astype(np.uint8)
