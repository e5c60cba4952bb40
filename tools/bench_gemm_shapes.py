"""Event timing of gemm_bf16_tn_kernel at the shapes of the streaming path: decoder prefill at M = 281 and
the ViT at M = 1024 / 3072, with the epilogues the model uses. CASE=<name> runs one shape twice (for `ncu -k regex:gemm`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi as A

ctx = A.Context(0)
SHAPES = {
    # name: (M, N, K, epilogue)
    "prefill_qkv": (281, 4608, 3584, A.EPI_BIAS),
    "prefill_o": (281, 3584, 3584, A.EPI_RESIDUAL),
    "prefill_gateup": (281, 37888, 3584, A.EPI_SWIGLU),
    "prefill_down": (281, 3584, 18944, A.EPI_RESIDUAL),
    "vit_qkv": (1024, 3840, 1280, A.EPI_BIAS),
    "vit_proj": (1024, 1280, 1280, A.EPI_BIAS_RESIDUAL),
    "vit_fc1": (1024, 5120, 1280, A.EPI_BIAS_QUICKGELU),
    "vit_fc2": (1024, 1280, 5120, A.EPI_BIAS_RESIDUAL),
    "vit_fc1_3072": (3072, 5120, 1280, A.EPI_BIAS_QUICKGELU),
    "vit_fc1_8192": (8192, 5120, 1280, A.EPI_BIAS_QUICKGELU),
    "mcq_gateup_2090": (2090, 37888, 3584, A.EPI_SWIGLU),
}


def run(name, iters=20, copies=8, block_n=0):
    M, N, K, epi = SHAPES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn((M, K), device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    ws = [(torch.randn((N, K), device="cuda", generator=g) * 0.02).to(torch.bfloat16) for _ in range(copies)]  # rotate weights: > L2
    bias = torch.zeros(N, dtype=torch.bfloat16, device="cuda")
    n_out = N // 2 if epi == A.EPI_SWIGLU else N
    res = torch.zeros((M, n_out), dtype=torch.bfloat16, device="cuda")
    out = torch.empty((M, n_out), dtype=torch.bfloat16, device="cuda")
    skw = torch.empty(8 * 384 * max(N, 8) * 4 // 4, dtype=torch.float32, device="cuda") if M <= 384 and N <= 8192 else None
    kw = dict(bias=bias if epi in (A.EPI_BIAS, A.EPI_BIAS_QUICKGELU, A.EPI_BIAS_GELU, A.EPI_BIAS_RESIDUAL) else None,
              residual=res if epi in (A.EPI_RESIDUAL, A.EPI_BIAS_RESIDUAL) else None, epilogue=epi, splitk_ws=skw, block_n=block_n)
    for w in ws:
        ctx.gemm(a, w, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ctx.gemm(a, ws[i % copies], out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    flops = 2.0 * M * N * K
    wbytes = N * K * 2
    ref = ""
    if os.environ.get("CUBLAS"):  # the library GEMM (torch.matmul -> cuBLASLt, no fused epilogue) on the same operands
        for w in ws:
            torch.matmul(a, w.t())
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            torch.matmul(a, ws[i % copies].t())
        e1.record()
        torch.cuda.synchronize()
        cu = e0.elapsed_time(e1) * 1e3 / iters
        ref = f"  | cuBLAS plain GEMM {cu:7.1f} us {flops / cu / 1e6:7.1f} TFLOP/s"
    print(f"{name:16s} bn={block_n:3d} M={M:5d} N={N:6d} K={K:6d}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  weights {wbytes / us / 1e3:7.1f} GB/s{ref}",
          flush=True)


if __name__ == "__main__":
    case = os.environ.get("CASE")
    if case:
        run(case, iters=2, copies=2, block_n=int(os.environ.get('BN', '0')))
    elif os.environ.get("SWEEP_BN"):
        for n in SHAPES:
            for bn in [int(x) for x in os.environ.get('BNS', '0,128,256,-128,-160,-192,-224,-256').split(',')]:
                try:
                    run(n, block_n=bn)
                except A.LiveCCNativeError as ex:
                    print(f"{n:16s} bn={bn:3d}: rejected ({str(ex)[:60]})", flush=True)
    else:
        for n in SHAPES:
            run(n)
