"""Shared inputs for the resize tests (CPU and GPU) and the golden generator: seeded uint8 clips [T, 3, h, w] mixing
smooth structure (what video frames look like; exercises the filter's overshoot -> clamp) and full-range noise."""
import hashlib

import numpy as np
import torch

# (name, T, h, w, H, W): down/up-scaling, one identical axis, odd sizes (unaligned rows, W % 4 != 0), windows of 2..21 taps
CASES = [
    ("720p_to_448", 2, 720, 1280, 448, 784),
    ("1080p_to_448sq", 2, 1080, 1920, 448, 448),
    ("360p_to_252", 2, 360, 640, 252, 448),
    ("upscale", 2, 300, 300, 448, 448),
    ("width_only", 1, 448, 500, 448, 448),
    ("height_only", 1, 500, 448, 448, 448),
    ("odd_sizes", 3, 97, 131, 28, 30),
    ("odd_up", 1, 33, 47, 57, 91),
    ("tiny_src", 1, 3, 5, 28, 28),
    ("one_px", 1, 1, 1, 4, 4),
    ("strong_down", 1, 480, 854, 28, 56),
]


def make_clip(name: str, T: int, h: int, w: int) -> torch.Tensor:
    seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")
    g = torch.Generator().manual_seed(seed)
    noise = torch.randint(0, 256, (T, 3, h, w), generator=g, dtype=torch.uint8)
    yy = torch.arange(h, dtype=torch.float64).view(1, 1, h, 1)
    xx = torch.arange(w, dtype=torch.float64).view(1, 1, 1, w)
    ph = torch.rand((T, 3, 1, 1), generator=g, dtype=torch.float64) * 6.283
    smooth = 127.5 + 127.5 * torch.sin(yy * 0.11 + ph) * torch.cos(xx * 0.07 - ph)
    edges = ((xx.long() // 9 + yy.long() // 7) % 2) * 255.0   # hard edges: ringing beyond [0, 255] before the clamp
    third = h // 3
    out = noise.clone()
    out[:, :, third:2 * third] = smooth[:, :, third:2 * third].round().clamp(0, 255).to(torch.uint8)
    out[:, :, 2 * third:] = edges.expand(T, 3, h, w)[:, :, 2 * third:].to(torch.uint8)
    return out.contiguous()


def digest(a) -> str:
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def torchvision_resize(clip: torch.Tensor, size) -> torch.Tensor:
    """The reference's call (video_process_patch.py:150-155)."""
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF

    return TF.resize(clip, list(size), interpolation=InterpolationMode.BICUBIC, antialias=True)
