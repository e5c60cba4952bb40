"""GPU frame ingest: lcc_resize_bicubic_aa_u8 through the C ABI against the oracle, the golden
digests of the reference's torchvision call, and live torchvision. Bar: bit-exact (uint8)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from resize_cases import CASES, digest, make_clip, torchvision_resize  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "resize_aa_golden.json")))


@pytest.fixture(scope="module")
def ctx():
    from livecc_b200 import _cabi

    c = _cabi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_matches_golden_and_oracle(ctx, case):
    from oracle import resize_aa

    name, T, h, w, H, W = case
    clip = make_clip(name, T, h, w)
    out = ctx.resize_bicubic_aa_u8(clip.cuda(), (H, W))
    torch.cuda.synchronize()
    assert out.shape == (T, 3, H, W) and out.dtype == torch.uint8
    assert digest(out) == GOLDEN["cases"][name]["sha256"]
    if T * h * w <= 2 * 720 * 1280:
        assert np.array_equal(out.cpu().numpy(), resize_aa.resize_bicubic_aa_u8(clip.numpy(), (H, W)))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_matches_live_torchvision(ctx, case):
    """The reference's own call on this box's CPU. Bit-exact on the ATen build the rounding order was pinned on; another
    build of the reference's dependency may move a handful of pixels by one level (oracle/resize_aa.py)."""
    name, T, h, w, H, W = case
    clip = make_clip(name, T, h, w)
    ref = torchvision_resize(clip, (H, W))
    out = ctx.resize_bicubic_aa_u8(clip.cuda(), (H, W)).cpu()
    if torch.backends.cpu.get_cpu_capability() == GOLDEN["cpu_capability"] and torch.__version__ == GOLDEN["torch"]:
        assert torch.equal(out, ref)
    else:
        d = (out.to(torch.int16) - ref.to(torch.int16)).abs()
        assert int(d.max()) <= 1 and float((d != 0).float().mean()) < 1e-3


def test_rows_per_cta_and_alignment_do_not_change_the_result(ctx):
    """Every tiling (rows per CTA) and every 16-byte phase of the source rows gives the same bytes; W % 4 != 0 takes the
    scalar store path."""
    from oracle import resize_aa

    for (h, w, H, W) in [(211, 173, 56, 84), (120, 90, 173, 201), (64, 257, 30, 61)]:
        clip = make_clip(f"t{h}x{w}", 2, h, w)
        want = torch.from_numpy(resize_aa.resize_bicubic_aa_u8(clip.numpy(), (H, W)))
        for th in (0, 1, 2, 4, 8, 16):
            assert torch.equal(ctx.resize_bicubic_aa_u8(clip.cuda(), (H, W), rows_per_cta=th).cpu(), want), (h, w, th)
        # shift the whole clip by 1..15 bytes inside a larger buffer
        flat = clip.flatten()
        for shift in (1, 3, 8, 15):
            buf = torch.zeros(flat.numel() + 32, dtype=torch.uint8, device="cuda")
            buf[shift:shift + flat.numel()] = flat.cuda()
            view = buf[shift:shift + flat.numel()].view(clip.shape)
            assert torch.equal(ctx.resize_bicubic_aa_u8(view, (H, W)).cpu(), want), (h, w, shift)


def test_full_size_properties(ctx):
    """1080p -> 448x796 (the ingest shape of a 16:9 source): constant frames stay constant (weights sum to 1 within
    rounding -> exact after rounding to uint8), a horizontal flip commutes with the resize up to the window asymmetry
    (<= 1 level), and planes are independent (a batched call equals per-plane calls)."""
    h, w, H, W = 1080, 1920, 448, 796
    for v in (0, 1, 127, 255):
        c = torch.full((1, 3, h, w), v, dtype=torch.uint8, device="cuda")
        assert bool((ctx.resize_bicubic_aa_u8(c, (H, W)) == v).all())
    clip = make_clip("full", 2, h, w).cuda()
    out = ctx.resize_bicubic_aa_u8(clip, (H, W))
    for t in range(2):
        for ch in range(3):
            one = ctx.resize_bicubic_aa_u8(clip[t, ch].contiguous(), (H, W))
            assert torch.equal(one, out[t, ch])
    flipped = ctx.resize_bicubic_aa_u8(clip.flip(-1).contiguous(), (H, W)).flip(-1)
    assert int((flipped.to(torch.int16) - out.to(torch.int16)).abs().max()) <= 1
    assert digest(out.cpu()) == digest(torchvision_resize(clip.cpu(), (H, W))) or \
        torch.backends.cpu.get_cpu_capability() != GOLDEN["cpu_capability"]


def test_mirror_routes_cuda_clips_through_the_kernel(ctx):
    """livecc_utils.get_smart_resized_clip(device=cuda) == the host path, and _spatial_resize_video on a CUDA clip."""
    from livecc_b200 import _cabi
    from livecc_b200.livecc_utils import video_process_patch as vpp

    reader = vpp.SyntheticVideoReader("synthetic://40x360x640@30?seed=3")
    reader.get_frame_timestamp(0)
    pts = reader._frame_pts[:, 1]
    stamps = torch.arange(0.0, 3.0, 0.5)
    host, ts_h, idx_h = vpp.get_smart_resized_clip(reader, 252, 448, stamps, pts)
    before = _cabi.launch_count()
    dev, ts_d, idx_d = vpp.get_smart_resized_clip(reader, 252, 448, stamps, pts, device="cuda")
    assert _cabi.launch_count() == before + 1
    assert dev.is_cuda and idx_h == idx_d and torch.equal(ts_h, ts_d)
    if torch.backends.cpu.get_cpu_capability() == GOLDEN["cpu_capability"]:
        assert torch.equal(dev.cpu(), host)
    else:
        assert int((dev.cpu().to(torch.int16) - host.to(torch.int16)).abs().max()) <= 1


def test_plan_rejects_windows_larger_than_shared_memory(ctx):
    from livecc_b200 import _cabi

    with pytest.raises(_cabi.LiveCCNativeError):
        ctx.resize_plan(30000, 30000, 28, 28)
    with pytest.raises(_cabi.LiveCCNativeError):  # output width 1 with a height change: outside the oracle's domain
        ctx.resize_plan(60, 20, 30, 1)
    ctx.resize_plan(60, 20, 60, 1)  # width only: fine
