"""Event timing of the decoder prefill attention (S new tokens over past+S cached tokens, 28 q heads / 4 KV heads,
head_dim 128, paged cache) for both kernels: LCC_ATTN_MMA (mma.sync) and LCC_ATTN_TC (tcgen05/TMEM); also checks that
the two agree."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi

ctx = _cabi.Context(0)
Hq, Hkv = 28, 4


def run(S, past, iters=10):
    T = past + S
    pages = (T + 63) // 64
    g = torch.Generator(device="cuda").manual_seed(S + past)
    k = torch.randn((pages, Hkv, 64, 128), device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn((pages, Hkv, 64, 128), device="cuda", generator=g).to(torch.bfloat16)
    tail = pages * 64 - T
    if tail:
        v[-1, :, 64 - tail:] = 0
    pt = torch.randperm(pages, device="cuda", generator=g).to(torch.int32)
    inv = torch.empty_like(pt)
    inv[pt.long()] = torch.arange(pages, device="cuda", dtype=torch.int32)
    k, v = k[inv.long()].contiguous(), v[inv.long()].contiguous()  # logical page i lives at physical page pt[i]
    q = torch.randn((S, (Hq + 2 * Hkv) * 128), device="cuda", generator=g).to(torch.bfloat16)
    flops = 4 * 128 * Hq * (S * past + S * (S + 1) / 2)
    outs = {}
    for impl in (1, 2):
        outs[impl] = ctx.attn_prefill(q, k, v, pt, Hq, Hkv, past, impl=impl, split=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ctx.attn_prefill(q, k, v, pt, Hq, Hkv, past, impl=impl, split=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"S {S:4d} past {past:6d} impl {impl}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
    d = (outs[1].float() - outs[2].float()).abs().max().item()
    print(f"   max |mma - tc| = {d:.3e}", flush=True)


if __name__ == "__main__":
    pasts = os.environ.get("PASTS")  # e.g. PASTS=9000,17000 for a short run (LIVECC_B200_ATTN_PTMEM=0/1 picks the tc variant)
    for past in ([int(x) for x in pasts.split(",")] if pasts else (0, 1000, 3000, 9000, 17000, 70000)):
        run(281, past)
    if not pasts:
        run(2084, 0, iters=5)
