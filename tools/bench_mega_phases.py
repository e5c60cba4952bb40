"""Per-phase streaming efficiency of the persistent decode kernel (LiveCC-7B dims): the sub-range hook runs one phase
(over all 28 layers) at a time; bytes = the weights (and KV) that phase streams; CUDA-event timed, 10 launches."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration

cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
t = cfg.text_config
H, I, L, V = t.hidden_size, t.intermediate_size, t.num_hidden_layers, t.vocab_size
qkv_dim = (t.num_attention_heads + 2 * t.num_key_value_heads) * 128
peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6574.1
only = os.environ.get("PHASES")  # e.g. "full" -> only the lines whose name contains it
cases = [(1, 1000), (1, 17000), (4, 8000), (8, 8000)] if not os.environ.get("CASES") else \
    [tuple(int(x) for x in c.split(":")) for c in os.environ["CASES"].split(",")]
print("look-ahead tiles:", os.environ.get("LIVECC_B200_MEGA_LOOKAHEAD", "default"))
for B, kv in cases:
    g = torch.Generator().manual_seed(0)
    reqs = [dict(input_ids=torch.randint(1000, 9000, (1, kv), generator=g).cuda()) for _ in range(B)]
    outs = eng.generate_batch(reqs, max_new_tokens=1)
    caches = [o.past_key_values for o in outs]
    with torch.inference_mode():
        for c in caches:
            c.scalars[_cabi.SC_FINISHED] = 0
    sts = [c.stream_state() for c in caches]
    kvb = 2 * t.num_key_value_heads * 128 * 2 * kv * B
    phases = [("qkv", 1, 0, qkv_dim * H * 2), ("attention", 2, 0, kvb), ("o_proj", 4, 0, H * H * 2), ("gate_up", 8, 0, 2 * I * H * 2),
              ("down", 16, 0, H * I * 2), ("layers", 31, 0, (qkv_dim * H + H * H + 3 * I * H) * 2 + kvb), ("lm_head", 0, 1, 0),
              ("full step", 31, 1, (qkv_dim * H + H * H + 3 * I * H) * 2 + kvb)]
    print(f"B={B} kv_len={kv}")
    for name, mask, head, per_layer in phases:
        if only and only not in name:
            continue
        nl = L if mask else 0
        nbytes = per_layer * nl + (V * H * 2 if head else 0)
        for _ in range(2):
            eng._native.decode_mega_debug(sts, 0, nl, mask, head)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng._native.decode_mega_debug(sts, 0, nl, mask, head)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  {name:10s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / ms / 1e6:7.0f} GB/s  {100 * nbytes / ms / 1e6 / peak:5.1f} % of {peak:.0f}"
              f"  err={eng._native.mega_error()}", flush=True)
    for c in caches:
        c.release()
