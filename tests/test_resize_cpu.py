"""Frame-resize oracle and host logic (no GPU): oracle/resize_aa.py against the golden digests generated from the
reference's own torchvision call, against live torchvision, and the C-ABI's host table builder against the oracle."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from resize_cases import CASES, digest, make_clip, torchvision_resize  # noqa: E402

from oracle import resize_aa  # noqa: E402

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "resize_aa_golden.json")))
SMALL = [c for c in CASES if c[1] * c[2] * c[3] <= 2 * 720 * 1280]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(case):
    name, T, h, w, H, W = case
    clip = make_clip(name, T, h, w)
    g = GOLDEN["cases"][name]
    assert digest(clip) == g["input_sha256"], "test input generator drifted from the golden generator"
    out = resize_aa.resize_bicubic_aa_u8(clip.numpy(), (H, W))
    flat = out.reshape(-1)
    assert [int(flat[i]) for i in g["probe_index"]] == g["probe_value"]
    assert digest(out) == g["sha256"]


@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_oracle_matches_live_torchvision(case):
    """Bit-exact where the ATen build matches the one the rounding order was fitted on (AVX512 kernels of the pinned
    wheel); on another build the reference's own result moves by at most one level on a handful of pixels."""
    name, T, h, w, H, W = case
    clip = make_clip(name, T, h, w)
    ref = torchvision_resize(clip, (H, W)).numpy()
    out = resize_aa.resize_bicubic_aa_u8(clip.numpy(), (H, W))
    if torch.backends.cpu.get_cpu_capability() == GOLDEN["cpu_capability"] and torch.__version__ == GOLDEN["torch"]:
        assert np.array_equal(out, ref)
    else:
        d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3


def test_oracle_float_pass_is_bit_exact():
    """The pre-rounding float32 image, both scaling directions (checks the fitted rounding order directly)."""
    if torch.backends.cpu.get_cpu_capability() != GOLDEN["cpu_capability"] or torch.__version__ != GOLDEN["torch"]:
        pytest.skip("rounding order pinned on a different ATen build")
    for (h, w, H, W) in [(100, 160, 56, 84), (120, 90, 173, 201), (250, 333, 28, 56)]:
        x = make_clip(f"f{h}x{w}", 1, h, w).float()
        ref = torch.nn.functional.interpolate(x, size=(H, W), mode="bicubic", antialias=True, align_corners=False).numpy()
        out = resize_aa.resize_bicubic_aa_f32(x.numpy(), (H, W))
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_identity_size_returns_input():
    clip = make_clip("id", 1, 28, 56)
    assert resize_aa.resize_bicubic_aa_u8(clip.numpy(), (28, 56)) is not None
    assert np.array_equal(resize_aa.resize_bicubic_aa_u8(clip.numpy(), (28, 56)), clip.numpy())


@pytest.mark.parametrize("in_size,out_size", [(160, 84), (300, 448), (1920, 448), (1080, 252), (131, 28), (448, 448),
                                              (854, 798), (7, 3), (3, 7), (1, 1), (2, 5), (4096, 28), (28, 1036)])
def test_cabi_host_table_matches_oracle(in_size, out_size):
    """lcc_resize_aa_table (C++ host code behind the C ABI) == the oracle's window/weight table, bit for bit."""
    from livecc_b200 import _cabi

    xmin, xsize, w = _cabi.resize_aa_table(in_size, out_size)
    table = resize_aa.aa_table(in_size, out_size)
    assert w.shape == (resize_aa.aa_taps(in_size, out_size), out_size)
    for k, (x0, ws) in enumerate(table):
        assert xmin[k] == x0 and xsize[k] == len(ws)
        assert np.array_equal(w[: len(ws), k].view(np.uint32), ws.view(np.uint32))
        assert not w[len(ws):, k].any()


def test_oracle_matches_torchvision_on_random_sizes():
    """60 seeded random (h, w) -> (H, W) pairs, 1..260 -> 1..200 plus a few large sources: bit-exact, except the one degenerate
    family the oracle header names (output width 1 with a height change), where torchvision's own result is not the separable
    filter; that family is asserted to differ so that a change in the dependency is noticed."""
    if torch.backends.cpu.get_cpu_capability() != GOLDEN["cpu_capability"] or torch.__version__ != GOLDEN["torch"]:
        pytest.skip("rounding order pinned on a different ATen build")
    rng = np.random.default_rng(7)
    for it in range(60):
        h, w = int(rng.integers(1, 260)), int(rng.integers(1, 260))
        H, W = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        if it % 10 == 0:
            h, w = int(rng.integers(300, 900)), int(rng.integers(300, 900))
        if W == 1 and H != h:
            W = 2
        clip = torch.randint(0, 256, (1, 2, h, w), generator=torch.Generator().manual_seed(it), dtype=torch.uint8)
        assert np.array_equal(resize_aa.resize_bicubic_aa_u8(clip.numpy(), (H, W)), torchvision_resize(clip, (H, W)).numpy()), (h, w, H, W)
    clip = torch.randint(0, 256, (1, 1, 60, 20), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    assert not np.array_equal(resize_aa.resize_bicubic_aa_u8(clip.numpy(), (30, 1)), torchvision_resize(clip, (30, 1)).numpy())
    assert np.array_equal(resize_aa.resize_bicubic_aa_u8(clip.numpy(), (60, 1)), torchvision_resize(clip, (60, 1)).numpy())  # width only


REF_SOURCES = "/root/reference/demo/sources"


@pytest.mark.skipif(not os.path.isdir(REF_SOURCES), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name", ["cvpr25_vlog.mp4", "howto_fix_laptop_mute_1080p.mp4",
                                  "warriors_vs_rockets_2025wcr1_mute_1080p.mp4"])
def test_oracle_matches_torchvision_on_the_reference_videos(name):
    """Real decoded frames of the three demo videos (REF/demo/sources), resized to the size the demo's reader picks
    (get_smart_resized_video_reader, max_pixels = 384*28*28 as REF/demo/infer.py:63): oracle == the reference's
    torchvision call, bit for bit. The GPU kernel equals the oracle bit for bit (tests/test_resize_gpu.py), which closes
    the chain kernel == torchvision on these videos without shipping the videos to the GPU box."""
    from livecc_b200.livecc_utils import video_process_patch as vpp

    reader, H, W = vpp.get_smart_resized_video_reader(os.path.join(REF_SOURCES, name), 384 * 28 * 28)
    n = len(reader)
    frames = torch.from_numpy(reader.get_batch([n // 3, (2 * n) // 3]).asnumpy()).permute(0, 3, 1, 2).contiguous()
    assert frames.shape[1] == 3 and frames.dtype == torch.uint8
    ref = torchvision_resize(frames, (H, W)).numpy()
    out = resize_aa.resize_bicubic_aa_u8(frames.numpy(), (H, W))
    if torch.backends.cpu.get_cpu_capability() == GOLDEN["cpu_capability"] and torch.__version__ == GOLDEN["torch"]:
        assert np.array_equal(out, ref)
    else:
        d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
