#!/bin/bash
O=gpurun_out/c29; mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 800 -k "vit_forward" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
python - > $O/vit_stats.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, "tests")
import test_engine_gpu as T
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
from oracle.restated import RestatedLiveCC
from livecc_b200.config import LiveCCConfig
from livecc_b200.checkpoint import synthetic_state_dict
from livecc_b200.processing import StubProcessor
cfg = LiveCCConfig.small()
sd = synthetic_state_dict(cfg, dtype=torch.bfloat16, device="cuda", gen_device="cuda")
eng = LiveCCB200ForConditionalGeneration.from_state_dict(cfg, sd, "cuda")
rs = RestatedLiveCC(cfg, sd)
proc = StubProcessor(cfg)
for frames, hw in [(2, (112, 112)), (6, (224, 140)), (2, (448, 448))]:
    inp = T.make_turn_inputs(proc, 0, frames, hw, 3)
    px = inp.pixel_values_videos.to("cuda")
    o = eng.get_video_features(px, inp.video_grid_thw).float(); r = rs.vit_forward(px, inp.video_grid_thw).float()
    rms = r.pow(2).mean(dim=1, keepdim=True).sqrt()
    ulps = (o - r).abs() / (2.0 ** -8 * torch.maximum(r.abs(), rms))
    cos = torch.nn.functional.cosine_similarity(o, r, dim=1)
    f = ulps.flatten()
    print(frames, hw, "p50", f.median().item(), "p99", f.kthvalue(int(0.99 * f.numel())).values.item(), "p999", f.kthvalue(int(0.999 * f.numel())).values.item(), "max", f.max().item(), "min_cos", cos.min().item())
PY
cat $O/rc.txt; tail -n 12 $O/tests.log | cut -c1-400; cat $O/vit_stats.txt | tail -5
