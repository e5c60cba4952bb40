"""BASELINE config #5 (offline batch-32 16-frame MCQ, VideoMME shape, synthetic): ViT + prefill only, no decode loop,
through `LiveCCB200ForConditionalGeneration.forward_mcq` = the reference's single-forward scoring
(REF/evaluation/distributed_mcq_predictor.py:72-105): left-padded batch, last-position logits over the letter ids.
Sweeps the batch size 1..32; per batch the vision tower runs once over all B*8 temporal segments (B*8192 patches) and the
decoder prefill once per sample (2090 tokens each). Reports samples/s, frames/s and the achieved tensor throughput
against the sustained bf16 peak (algorithmic FLOPs: SURVEY.md §8(d): 1.49 TFLOP per 1024-patch segment;
2*6.525e9*S + 4*S*(S/2)*3584*28 per prefill)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
from livecc_b200.processing import StubProcessor

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,2,4,8,16,32")
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
proc = StubProcessor(cfg)
letters = [proc.tokenizer(f": {c}").input_ids[-1] for c in "ABCD"]
peak = 1441.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    pass


def sample(i):
    g = torch.Generator().manual_seed(i)
    low = torch.rand((args.frames, 3, 28, 28), generator=g)
    clip = (torch.nn.functional.interpolate(low, size=(448, 448), mode="bilinear") * 255).to(torch.uint8)
    q = "Which option is correct? " + " ".join(f"word{j}" for j in range(i % 7)) + " A. one B. two C. three D. four"
    content = [{"type": "video", "video": clip}, {"type": "text", "text": q + "\nPlease select the correct answer."}]
    text = proc.apply_chat_template([{"role": "user", "content": content}], tokenize=False, add_generation_prompt=True) + "Answer:"
    return proc(text=text, videos=[clip], return_attention_mask=False)


pool = [sample(i) for i in range(8)]
t = cfg.text_config
print("| batch | ms / batch | samples/s | frames/s | ViT+prefill TFLOP/s | % of sustained bf16 peak |")
print("|---|---|---|---|---|---|")
for B in [int(x) for x in args.batches.split(",")]:
    sub = [pool[i % len(pool)] for i in range(B)]
    Lmax = max(s.input_ids.shape[1] for s in sub)
    ids = torch.full((B, Lmax), cfg.pad_token_id, dtype=torch.int64)
    mask = torch.zeros((B, Lmax), dtype=torch.int64)
    flops = 0.0
    for b, s in enumerate(sub):
        n = s.input_ids.shape[1]
        ids[b, Lmax - n:] = s.input_ids[0]
        mask[b, Lmax - n:] = 1
        flops += ((args.frames + 1) // 2) * 1.49e12 + 2 * 6_525_288_448 * n + 4 * n * (n / 2) * t.hidden_size * t.num_hidden_layers
    px = torch.cat([s.pixel_values_videos for s in sub]).to("cuda")
    grid = torch.cat([s.video_grid_thw for s in sub])
    ids, mask = ids.to("cuda"), mask.to("cuda")
    eng.forward_mcq(ids, mask, letters, pixel_values_videos=px, video_grid_thw=grid)   # warm-up (workspace growth)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        pred, _ = eng.forward_mcq(ids, mask, letters, pixel_values_videos=px, video_grid_thw=grid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    tf = flops / (ms / 1e3) / 1e12
    print(f"| {B} | {ms:.1f} | {B / (ms / 1e3):.2f} | {B * args.frames / (ms / 1e3):.0f} | {tf:.0f} | {100 * tf / peak:.1f} |", flush=True)
