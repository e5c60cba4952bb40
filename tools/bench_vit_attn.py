"""Correctness (vs an fp32 reference) and event timing of the two ViT attention kernels
(LCC_VIT_ATTN_MMA = mma.sync flash kernels, LCC_VIT_ATTN_TC = tcgen05/TMEM kernel) on cu_seqlens segments."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi

ctx = _cabi.Context(0)
HD = 80


def ref_fp32(qkv, cu, heads):
    N = qkv.shape[0]
    q, k, v = qkv.reshape(N, 3, heads, HD).permute(1, 2, 0, 3).unbind(0)
    outs = []
    for s, e in zip(cu[:-1], cu[1:]):
        w = torch.softmax(torch.matmul(q[:, s:e].float(), k[:, s:e].float().transpose(1, 2)) * HD ** -0.5, dim=-1)
        outs.append(torch.matmul(w, v[:, s:e].float()))
    return torch.cat(outs, dim=1).transpose(0, 1).reshape(N, heads * HD)


def run(seglens, heads, impls=(1, 2), iters=20, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(sum(seglens) + heads)
    N = sum(seglens)
    qkv = (torch.randn((N, 3 * heads * HD), device="cuda", generator=g) * scale).to(torch.bfloat16)
    cu = [0]
    for n in seglens:
        cu.append(cu[-1] + n)
    cu_t = torch.tensor(cu, dtype=torch.int32, device="cuda")
    ref = ref_fp32(qkv, cu, heads)
    flops = sum(4 * n * n * HD * heads for n in seglens)
    for impl in impls:
        out = ctx.vit_attention(qkv, cu_t, max(seglens), heads, HD, impl=impl)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs()
        tol = 6e-3 + 1.5e-2 * ref.abs()
        bad = (err > tol).float().mean().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ctx.vit_attention(qkv, cu_t, max(seglens), heads, HD, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"segs {str(seglens[:3]):>18s}x{len(seglens):<2d} heads {heads:2d} impl {impl}: max err {err.max().item():.3e} "
              f"bad {bad:.1e}  {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


def main(impls):
    if os.environ.get("CASE") == "mcq":  # one shape only (for ncu)
        run([1024] * 8, 16, impls, iters=2)
        return
    run([128], 1, impls, iters=2)
    run([256], 2, impls, iters=2)
    run([1024], 16, impls)
    run([64, 80], 4, impls, iters=2)
    run([200, 1024, 328], 16, impls)
    run([1196] * 4, 16, impls)
    run([1024] * 8, 16, impls)
    run([1024], 16, impls, scale=4.0)


if __name__ == "__main__":
    impls = tuple(int(x) for x in os.environ.get("IMPLS", "1,2").split(","))
    for bn in os.environ.get("BNS", "64").split(","):
        os.environ["LIVECC_B200_VIT_TC_BN"] = bn
        print(f"--- tc kernel keys/tile = {bn}")
        main(impls)


