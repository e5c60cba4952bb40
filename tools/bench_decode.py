"""Decode-step microbenchmark (7B dims): ms per decode step (CUDA-graph replay) at a given KV length.
usage: python tools/bench_decode.py [--kv 8192] [--steps 64]   (env LIVECC_B200_NO_PDL=1 / LIVECC_B200_NO_GRAPH=1)"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration

ap = argparse.ArgumentParser()
ap.add_argument("--kv", type=int, default=8192)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--model", default="7b")
args = ap.parse_args()
cfg = LiveCCConfig.livecc_7b() if args.model == "7b" else LiveCCConfig.small()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
dev = eng.device
# one text-only prefill of `kv` random ids (chunked to the workspace capacity by generate's own growth)
ids = torch.randint(1000, 5000, (1, args.kv), device=dev)
t0 = time.time()
out = eng.generate(input_ids=ids, max_new_tokens=2, repetition_penalty=1.05)
torch.cuda.synchronize()
print(f"prefill {args.kv} tokens: {time.time() - t0:.2f}s, phases {eng.last_stats}")
cache = out.past_key_values
for rep in range(3):
    ids2 = torch.cat([out.sequences[:, :-1], torch.randint(1000, 5000, (1, 8), device=dev)], 1)
    torch.cuda.synchronize()
    out = eng.generate(input_ids=ids2, past_key_values=cache, max_new_tokens=args.steps + 1, repetition_penalty=1.05)
    st = eng.last_stats
    print(f"rep {rep}: generated {st['generated']}, decode {st['decode_ms']:.2f} ms -> {st['decode_ms'] / max(st['generated'] - 1, 1):.3f} ms/step "
          f"(kv {st['kv_len']}), prefill(8 tok) {st['prefill_ms']:.2f} ms")
t = cfg.text_config
wbytes = (t.num_hidden_layers * ((t.num_attention_heads + 2 * t.num_key_value_heads) * 128 * t.hidden_size + t.hidden_size ** 2
                                 + 3 * t.intermediate_size * t.hidden_size) + t.vocab_size * t.hidden_size) * 2
kvb = st["kv_len"] * 2 * t.num_hidden_layers * t.num_key_value_heads * 128 * 2
ms = st["decode_ms"] / max(st["generated"] - 1, 1)
print(f"bytes/step {(wbytes + kvb) / 1e9:.2f} GB -> {(wbytes + kvb) / ms / 1e6:.0f} GB/s")

# ---- launch-overhead probe: replay the captured decode-step graph with the finished flag set, so every kernel
#      exits right after its flag check. What remains is the per-node launch/scheduling cost of the 170-node graph.
g = list(eng._graphs.values())[-1]
with torch.inference_mode():
    cache.scalars[_cabi.SC_FINISHED] = 1
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"no-op graph replay (all kernels early-exit): {e0.elapsed_time(e1) / 50:.3f} ms per step "
      f"for {cfg.text_config.num_hidden_layers * 5 + 2} kernel nodes")
