"""GPU frame ingest: lcc_resize_bicubic_aa_u8 (one kernel: width pass + height pass + clamp/round) vs the reference's host call
transforms.functional.resize(clip, [H, W], BICUBIC, antialias=True) (video_process_patch.py:150-155) on this box's CPU.
Algorithmic bytes = source + destination uint8 planes; inputs rotate over > L2 of distinct clips; CUDA-event timing.
usage: python tools/bench_resize.py [--sweep]   (--sweep: rows-per-CTA tuning hook)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--iters", type=int, default=40)
args = ap.parse_args()
ctx = _cabi.Context(0)
peak = 6574.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SHAPES = [("1080p -> 448x796, 2 frames", 2, 1080, 1920, 448, 796), ("1080p -> 448x796, 6 frames", 6, 1080, 1920, 448, 796),
          ("720p -> 448x784, 2 frames", 2, 720, 1280, 448, 784), ("4K -> 448x796, 2 frames", 2, 2160, 3840, 448, 796),
          ("480p -> 448x798 (up), 2 frames", 2, 480, 854, 448, 798)]


def time_gpu(clips, size, th):
    outs = [ctx.resize_bicubic_aa_u8(c, size, rows_per_cta=th) for c in clips[:2]]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = outs[0]
    e0.record()
    for i in range(args.iters):
        ctx.resize_bicubic_aa_u8(clips[i % len(clips)], size, out=out, rows_per_cta=th)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters * 1e3  # us


print(f"| shape | rows/CTA (smem) | GPU us | GB/s (src+dst) | % of HBM peak {peak:.0f} | torchvision CPU ms ({torch.get_num_threads()} threads) | speed-up |")
print("|---|---|---|---|---|---|---|")
for name, T, h, w, H, W in SHAPES:
    nclip = max(2, int(160e6 // (T * 3 * h * w)) + 1)
    g = torch.Generator().manual_seed(0)
    clips = [torch.randint(0, 256, (T, 3, h, w), generator=g, dtype=torch.uint8).cuda() for _ in range(nclip)]
    bytes_alg = T * 3 * (h * w + H * W)
    host = clips[0].cpu()
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF
    TF.resize(host, [H, W], interpolation=InterpolationMode.BICUBIC, antialias=True)
    t0 = time.perf_counter()
    for _ in range(3):
        ref = TF.resize(host, [H, W], interpolation=InterpolationMode.BICUBIC, antialias=True)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    same = torch.equal(ctx.resize_bicubic_aa_u8(clips[0], (H, W)).cpu(), ref)
    for th in ([0, 4, 8, 16, 32] if args.sweep else [0]):
        try:
            info = ctx.resize_plan_info(ctx.resize_plan(h, w, H, W, th))
        except _cabi.LiveCCNativeError:
            continue
        us = time_gpu(clips, (H, W), th)
        gbs = bytes_alg / us / 1e3
        print(f"| {name}{'' if same else ' (MISMATCH vs torchvision)'} | {info['rows_per_cta']} ({info['smem_bytes'] // 1024} KB) | {us:.1f} | {gbs:.0f} | "
              f"{100 * gbs / peak:.1f} | {cpu_ms:.1f} | {cpu_ms * 1e3 / us:.0f}x |", flush=True)
    del clips
