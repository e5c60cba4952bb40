"""Does CUDA-graph replay of the ViT pass (263 launches per 1024-patch chunk) beat eager launches? 7B dims, synthetic weights.
Prints ms per pass for eager and graph replay, for the 2-frame chunk (1024 patches) and the 6-frame opening chunk (3072)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
from livecc_b200.processing import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD

cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
mean255 = (torch.tensor(OPENAI_CLIP_MEAN) * (1.0 / (1 / 255))).tolist()
std255 = (torch.tensor(OPENAI_CLIP_STD) * (1.0 / (1 / 255))).tolist()
for T in (2, 6):
    fr = torch.randint(0, 256, (T, 3, 448, 448), dtype=torch.uint8, device="cuda")
    n = (T // 2) * 32 * 32
    eng._ensure_workspace(n, 0)
    out = torch.empty((n // 4, cfg.vision_config.hidden_size), dtype=torch.bfloat16, device="cuda")
    with torch.inference_mode():
        for _ in range(3):
            eng._native.vit_forward_frames(fr, mean255, std255, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng._native.vit_forward_frames(fr, mean255, std255, out)
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 20
        ref = out.clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eng._native.vit_forward_frames(fr, mean255, std255, out)
            with torch.cuda.graph(g, stream=s):
                eng._native.vit_forward_frames(fr, mean255, std255, out)
        torch.cuda.current_stream().wait_stream(s)
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        same = torch.equal(out, ref)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / 20
    print(f"ViT pass, {T} frames ({n} patches): eager {eager:.3f} ms, graph replay {graph:.3f} ms, identical output: {same}", flush=True)
