"""Generates the golden vectors under tests/golden/ from the INSTALLED reference implementation
(transformers Qwen2VLForConditionalGeneration + Qwen2VLVideoProcessor, fp32 on CPU).

Run here (no GPU needed):  python tests/golden/make_golden.py
The reference repo holds no tests or golden vectors for this path, so these files pin the
oracle restatement (oracle/restated.py) and the host logic (positions, patchify) against outputs of the
third-party code the reference actually executes. Version: transformers 5.5.0.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from livecc_b200.checkpoint import synthetic_state_dict  # noqa: E402
from livecc_b200.config import LiveCCConfig  # noqa: E402
from livecc_b200.processing import StubProcessor  # noqa: E402
from oracle.hf_oracle import build_hf_model, hf_generate_chunk  # noqa: E402


def turn_inputs(proc, turn, frames, hw, seed):
    g = torch.Generator().manual_seed(seed)
    clip = torch.randint(0, 256, (frames, 3, hw[0], hw[1]), generator=g, dtype=torch.uint8)
    t0 = 0.0 if turn == 0 else 3.0 + (turn - 1)
    content = [{"type": "text", "text": f"Time={t0:.1f}-{3.0 + turn:.1f}s"}, {"type": "video", "video": clip}]
    if turn == 0:
        content.append({"type": "text", "text": "Please describe the video."})
    text = proc.apply_chat_template([{"role": "user", "content": content}], tokenize=False, add_generation_prompt=True)
    if turn > 0:
        text = "<|im_end|>\n" + text[text.index("<|im_start|>user"):]
    return proc(text=text, videos=[clip], return_attention_mask=False), clip


def main():
    import transformers
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor

    out = {"transformers": transformers.__version__, "torch": torch.__version__}

    # 1. video processor: patch rows of uint8 clips (checksum + a few probes)
    vp = Qwen2VLVideoProcessor(min_pixels=3136, max_pixels=12845056)
    cases = []
    for (T, H, W, seed) in [(1, 224, 224, 0), (2, 448, 448, 1), (6, 112, 140, 2), (3, 56, 84, 3)]:
        g = torch.Generator().manual_seed(seed)
        clip = torch.randint(0, 256, (T, 3, H, W), generator=g, dtype=torch.uint8)
        r = vp(videos=[clip], return_tensors="pt", do_sample_frames=False)
        px = r["pixel_values_videos"]
        cases.append({"T": T, "H": H, "W": W, "seed": seed, "grid": r["video_grid_thw"].tolist(),
                      "shape": list(px.shape), "sum": float(px.double().sum()), "abs_sum": float(px.double().abs().sum()),
                      "probe": [float(px[i % px.shape[0], (i * 37) % px.shape[1]]) for i in range(0, 400, 13)]})
    out["video_processor"] = cases

    # 2. get_rope_index tables (5.5.0 semantics)
    cfg = LiveCCConfig.small()
    sd = synthetic_state_dict(cfg, dtype=torch.float32)
    model = build_hf_model(cfg, sd, dtype=torch.float32)
    rope = []
    V, VS, VE = cfg.video_token_id, cfg.vision_start_token_id, cfg.vision_end_token_id
    for ids, grids in [([1, 2, VS] + [V] * 18 + [VE, 3, 4], [[3, 4, 6]]),
                       ([5, VS] + [V] * 64 + [VE, 7, 8, 9], [[1, 16, 16]]),
                       ([5, VS] + [V] * 16 + [VE, 7, VS] + [V] * 12 + [VE, 9], [[1, 8, 8], [2, 4, 6]])]:
        t = torch.tensor([ids])
        mm = (t == V).to(torch.int32) * 2
        pos, delta = model.model.get_rope_index(t, mm_token_type_ids=mm, video_grid_thw=torch.tensor(grids))
        rope.append({"ids": ids, "grids": grids, "pos": pos[:, 0].tolist(), "delta": int(delta)})
    out["rope_index"] = rope

    # 3. streaming generate: 3 turns (6, 2, 2 frames @112x140), 6 greedy tokens each, fp32 eager on CPU
    proc = StubProcessor(cfg)
    past_kv = past_ids = None
    turns = []
    for turn, frames in enumerate([6, 2, 2]):
        inp, _ = turn_inputs(proc, turn, frames, (112, 140), 100 + turn)
        o, L = hf_generate_chunk(model, inp, past_kv, past_ids, max_new_tokens=6, output_logits=True)
        past_kv, past_ids = o.past_key_values, o.sequences[:, :-1]
        lg = [x[0] for x in o.logits]
        turns.append({"frames": frames, "new_ids": inp.input_ids[0].tolist(), "generated": o.sequences[0, L:].tolist(),
                      "kv_len": past_kv.get_seq_length(), "rope_delta": int(model.model.rope_deltas),
                      "top5": [[[int(i), float(v)] for v, i in zip(*l.topk(5))] for l in lg],
                      "logit_sum": [float(l.double().sum()) for l in lg]})
    out["streaming"] = {"config": "small", "seed": 1234, "hw": [112, 140], "turns": turns}
    json.dump(out, open(os.path.join(HERE, "hf_golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "hf_golden.json"))


if __name__ == "__main__":
    main()
