"""Event-timed microbench of the decode GEMVs over all 28 layers' weights (inputs >> L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
t = cfg.text_config
H, I = t.hidden_size, t.intermediate_size
x = torch.randn(H, device="cuda").to(torch.bfloat16)
a = torch.randn(I, device="cuda").to(torch.bfloat16)
h = x.clone()
def timeit(fn, nbytes, name, iters=5):
    for lw in eng.weights.layers: fn(lw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for lw in eng.weights.layers: fn(lw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * len(eng.weights.layers))
    print(f"{name:10s} {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s  ({nbytes / us / 1e3 / 6574.1 * 100:.1f}% of measured HBM peak)")
timeit(lambda lw: eng.ctx.gemv_norm_swiglu(lw.gate_up_w, x, lw.ln2_w, 1e-6), 2 * I * H * 2, "gate_up")
timeit(lambda lw: eng.ctx.gemv_residual(lw.down_w, a, h), I * H * 2, "down")
timeit(lambda lw: eng.ctx.gemv_residual(lw.o_w, x, h), H * H * 2, "o_proj")
timeit(lambda lw: eng.ctx.gemv_norm_bias(lw.qkv_w, x, lw.ln1_w, 1e-6, lw.qkv_b), lw_bytes := (t.num_attention_heads + 2 * t.num_key_value_heads) * 128 * H * 2, "qkv")
