"""Head-to-head with "the kernel to beat" (SURVEY.md §2.3 K4/K11, §8(d)): flash-attn 2 (the reference's attention,
REF/demo/infer.py:46; its sm_100 cubins are mma.sync + cp.async code) vs the tcgen05/TMEM kernels of this repo at the
§8 shapes, on the same B200, CUDA-event timed, L2 flushed between iterations. Writes a markdown table to stdout
(copy under profiles/). flash-attn is LIBRARY code used here only as the measured competitor."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi

ctx = _cabi.Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # 256 MB > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3  # us


def vit(nseg, heads=16, hd=80, seg=1024):
    from flash_attn import flash_attn_varlen_func

    N = nseg * seg
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn((N, 3 * heads * hd), device="cuda", generator=g).to(torch.bfloat16)
    cu = torch.arange(0, N + 1, seg, dtype=torch.int32, device="cuda")
    q, k, v = qkv.view(N, 3, heads, hd).unbind(1)
    ours = lambda: ctx.vit_attention(qkv, cu, seg, heads, hd, impl=2)
    fa2 = lambda: flash_attn_varlen_func(q, k, v, cu, cu, seg, seg, causal=False)
    a, b = ours().float(), fa2().reshape(N, heads * hd).float()
    flops = 4 * seg * seg * hd * heads * nseg
    t_o, t_f = timed(ours), timed(fa2)
    print(f"| ViT attention {nseg}x{seg} patches, 16 heads x 80 | {t_f:.1f} | {flops / t_f / 1e6:.0f} | {t_o:.1f} | "
          f"{flops / t_o / 1e6:.0f} | {t_f / t_o:.2f}x | {(a - b).abs().max().item():.3e} |", flush=True)


def prefill(S, past, Hq=28, Hkv=4):
    from flash_attn import flash_attn_func

    T = past + S
    pages = (T + 63) // 64
    g = torch.Generator(device="cuda").manual_seed(T)
    k = torch.randn((pages, Hkv, 64, 128), device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn((pages, Hkv, 64, 128), device="cuda", generator=g).to(torch.bfloat16)
    if pages * 64 > T:
        v[-1, :, T - (pages - 1) * 64:] = 0
    pt = torch.arange(pages, dtype=torch.int32, device="cuda")
    q = torch.randn((S, (Hq + 2 * Hkv) * 128), device="cuda", generator=g).to(torch.bfloat16)
    # FA2 sees the contiguous [1, T, Hkv, 128] cache the reference re-concatenates every step (cache_utils.py:119-120)
    kk = k.permute(0, 2, 1, 3).reshape(1, pages * 64, Hkv, 128)[:, :T].contiguous()
    vv = v.permute(0, 2, 1, 3).reshape(1, pages * 64, Hkv, 128)[:, :T].contiguous()
    qq = q[:, : Hq * 128].reshape(1, S, Hq, 128)
    ours = lambda: ctx.attn_prefill(q, k, v, pt, Hq, Hkv, past, impl=2, split=True)
    fa2 = lambda: flash_attn_func(qq, kk, vv, causal=True)
    a, b = ours().float(), fa2().reshape(S, Hq * 128).float()
    flops = 4 * 128 * Hq * (S * past + S * (S + 1) / 2)
    t_o, t_f = timed(ours), timed(fa2)
    print(f"| decoder prefill attention S={S}, past={past}, GQA 28:4, d=128 | {t_f:.1f} | {flops / t_f / 1e6:.0f} | "
          f"{t_o:.1f} | {flops / t_o / 1e6:.0f} | {t_f / t_o:.2f}x | {(a - b).abs().max().item():.3e} |", flush=True)


if __name__ == "__main__":
    print("| shape | flash-attn 2.8.3 us | TFLOP/s | this repo (tcgen05) us | TFLOP/s | speed-up | max abs diff |")
    print("|---|---|---|---|---|---|---|")
    for nseg in (1, 3, 8):
        vit(nseg)
    for past in (0, 1000, 9000, 17000, 70000):
        prefill(281, past)
    prefill(2084, 0)
