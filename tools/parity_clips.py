"""Greedy-id parity over many synthetic clips: engine vs the installed HF model (bf16, same GPU).
Writes a JSON summary: exact-match rate of free-running ids, worst teacher-forced |dlogit|, and for every
diverging step the oracle's top-1/top-2 margin (divergences must sit on sub-tolerance margins).
usage: python tools/parity_clips.py [--clips 100] [--model small] [--out profiles/r01_parity_clips.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200.checkpoint import synthetic_state_dict
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
from livecc_b200.processing import StubProcessor
from oracle.hf_oracle import build_hf_model, hf_generate_chunk

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=100)
ap.add_argument("--model", default="small")
ap.add_argument("--size", type=int, default=112)
ap.add_argument("--tokens", type=int, default=8)
ap.add_argument("--out", default="profiles/r01_parity_clips.json")
args = ap.parse_args()
DEV = "cuda"
cfg = LiveCCConfig.livecc_7b() if args.model == "7b" else LiveCCConfig.small()
sd = synthetic_state_dict(cfg, dtype=torch.bfloat16, device=DEV, gen_device=DEV)
eng = LiveCCB200ForConditionalGeneration.from_state_dict(cfg, sd, DEV)
hf = build_hf_model(cfg, sd, dtype=torch.bfloat16, device=DEV, attn_implementation="sdpa")
proc = StubProcessor(cfg)


def turn_inputs(clip_id, turn, frames):
    g = torch.Generator().manual_seed(clip_id * 10 + turn)
    low = torch.rand((frames, 3, max(2, args.size // 16), max(2, args.size // 16)), generator=g)
    clip = (torch.nn.functional.interpolate(low, size=(args.size, args.size), mode="bilinear") * 255).to(torch.uint8)
    t0 = 0.0 if turn == 0 else 3.0 + (turn - 1)
    content = [{"type": "text", "text": f"Time={t0:.1f}-{3.0 + turn:.1f}s"}, {"type": "video", "video": clip}]
    if turn == 0:
        content.append({"type": "text", "text": "Please describe the video."})
    text = proc.apply_chat_template([{"role": "user", "content": content}], tokenize=False, add_generation_prompt=True)
    if turn > 0:
        text = "<|im_end|>\n" + text[text.index("<|im_start|>user"):]
    return proc(text=text, videos=[clip], return_attention_mask=False)


exact_clips = exact_turns = total_turns = steps = 0
worst = 0.0
divergences = []
for cid in range(args.clips):
    hf.model.rope_deltas = None
    kv = past = cache_tf = past_tf = cache_fr = past_fr = None
    clip_ok = True
    for turn, frames in enumerate([6, 2]):
        inp = turn_inputs(cid, turn, frames)
        new_ids = inp.input_ids.to(DEV)
        px, grid = inp.pixel_values_videos.to(DEV), inp.video_grid_thw
        o, L = hf_generate_chunk(hf, inp, kv, past, max_new_tokens=args.tokens, output_logits=True)
        kv, past = o.past_key_values, o.sequences[:, :-1]
        gen = o.sequences[0, L:].tolist()
        ids_tf = new_ids if past_tf is None else torch.cat([past_tf, new_ids], 1)
        out = eng.generate(input_ids=ids_tf, pixel_values_videos=px, video_grid_thw=grid, past_key_values=cache_tf,
                           repetition_penalty=1.05, max_new_tokens=len(gen), output_logits=True, _forced_ids=gen)
        cache_tf, past_tf = out.past_key_values, out.sequences[:, :-1]
        for lo, le in zip(o.logits, out.logits):
            worst = max(worst, (lo[0].float() - le.float()).abs().max().item())
            steps += 1
        if clip_ok:
            ids_fr = new_ids if past_fr is None else torch.cat([past_fr, new_ids], 1)
            fr = eng.generate(input_ids=ids_fr, pixel_values_videos=px, video_grid_thw=grid, past_key_values=cache_fr,
                              repetition_penalty=1.05, max_new_tokens=args.tokens)
            cache_fr, past_fr = fr.past_key_values, fr.sequences[:, :-1]
            gen_fr = fr.sequences[0, ids_fr.shape[1]:].tolist()
            total_turns += 1
            if gen_fr == gen:
                exact_turns += 1
            else:
                clip_ok = False
                k = next(i for i, (a, b) in enumerate(zip(gen, gen_fr)) if a != b)
                # oracle margin at the first diverging step, after the repetition penalty
                lg = o.logits[k][0].float().clone()
                hist = o.sequences[0, : L + k]
                sc = lg[hist]
                lg[hist] = torch.where(sc < 0, sc * 1.05, sc / 1.05)
                top2 = lg.topk(2).values
                divergences.append({"clip": cid, "turn": turn, "step": k, "oracle": gen[k], "engine": gen_fr[k],
                                    "oracle_margin": float(top2[0] - top2[1])})
    for c in (cache_tf, cache_fr):
        if c is not None:
            c.release()
    exact_clips += int(clip_ok)
summary = {"model": cfg.name, "oracle": "transformers 5.5.0 Qwen2VLForConditionalGeneration bf16 sdpa (same GPU)",
           "clips": args.clips, "turns_per_clip": 2, "tokens_per_turn": args.tokens, "frame_size": args.size,
           "clips_exact": exact_clips, "turns_compared": total_turns, "turns_exact": exact_turns,
           "teacher_forced_steps": steps, "worst_abs_dlogit": worst, "divergences": divergences}
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(summary, open(args.out, "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "divergences"}))
print("divergences:", divergences[:10])
