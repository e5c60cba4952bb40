"""BASELINE config #5 shape (offline MCQ: 16 frames at 448x448 per sample -> grid [8,32,32], 8192 patches, 2048
vision tokens + prompt; ViT + prefill only, no decode loop; REF/evaluation/distributed_mcq_predictor.py:75-105).
The engine is one-stream-per-call, so a "batch" is a loop over samples. Reports ViT and prefill time per sample and
the achieved tensor throughput of the ViT (1.49 TFLOP per 1024-patch segment)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
from livecc_b200.processing import StubProcessor

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
args = ap.parse_args()
cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
proc = StubProcessor(cfg, emit_frames=True)
g = torch.Generator().manual_seed(0)
low = torch.rand((args.frames, 3, 28, 28), generator=g)
clip = (torch.nn.functional.interpolate(low, size=(448, 448), mode="bilinear") * 255).to(torch.uint8)
content = [{"type": "video", "video": clip}, {"type": "text", "text": "Which option is correct? A. B. C. D. Answer:"}]
text = proc.apply_chat_template([{"role": "user", "content": content}], tokenize=False, add_generation_prompt=True)
inp = proc(text=text, videos=[clip], return_attention_mask=False).to("cuda")
S = inp.input_ids.shape[1]
for _ in range(2):  # warm-up (workspace growth, graph-free path)
    eng.generate(**inp, max_new_tokens=1).past_key_values.release()
torch.cuda.synchronize()
vit = pre = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.batch):
    out = eng.generate(**inp, max_new_tokens=1)
    vit += eng.last_stats["vit_ms"]
    pre += eng.last_stats["prefill_ms"]
    out.past_key_values.release()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1)
segs = (args.frames + 1) // 2
vit_tflop = segs * 1.49
t = cfg.text_config
pre_tflop = (2 * 6_525_288_448 * S + 4 * S * S / 2 * t.hidden_size * t.num_hidden_layers) / 1e12
print(f"samples {args.batch}, {args.frames} frames -> {segs * 1024} patches, {S} prompt tokens")
print(f"ViT      {vit / args.batch:8.2f} ms/sample  -> {vit_tflop / (vit / args.batch / 1e3):7.1f} TFLOP/s "
      f"({vit_tflop / (vit / args.batch / 1e3) / 1441.0 * 100:.1f}% of sustained bf16 peak)")
print(f"prefill  {pre / args.batch:8.2f} ms/sample  -> {pre_tflop / (pre / args.batch / 1e3):7.1f} TFLOP/s "
      f"({pre_tflop / (pre / args.batch / 1e3) / 1441.0 * 100:.1f}% of sustained bf16 peak)")
print(f"total    {total / args.batch:8.2f} ms/sample  -> {args.batch / (total / 1e3):.2f} samples/s, "
      f"{args.batch * args.frames / (total / 1e3):.1f} frames/s")
