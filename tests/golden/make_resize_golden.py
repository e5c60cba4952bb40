"""Golden vectors for the frame resize: outputs of the reference's own call
`transforms.functional.resize(clip, [H, W], BICUBIC, antialias=True)` (video_process_patch.py:150-155) from the
torchvision / torch wheels in this image, stored as SHA-256 digests of the uint8 result plus a few probe pixels.
Run from the repo root:  python tests/golden/make_resize_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torchvision

from resize_cases import CASES, digest, make_clip, torchvision_resize

out = {"generator": "tests/golden/make_resize_golden.py", "torch": torch.__version__, "torchvision": torchvision.__version__,
       "cpu_capability": torch.backends.cpu.get_cpu_capability(), "cases": {}}
for name, T, h, w, H, W in CASES:
    clip = make_clip(name, T, h, w)
    ref = torchvision_resize(clip, (H, W))
    flat = ref.flatten()
    probes = torch.linspace(0, flat.numel() - 1, 16).long()
    out["cases"][name] = {"in": [T, 3, h, w], "out": [H, W], "input_sha256": digest(clip), "sha256": digest(ref),
                          "probe_index": probes.tolist(), "probe_value": flat[probes].tolist()}
    print(name, out["cases"][name]["sha256"][:16])
with open(os.path.join(os.path.dirname(__file__), "resize_aa_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
