"""REF/livecc-utils/src/livecc_utils/video_process_patch.py restated without its two missing
dependencies.

* `decord` (absent here): the reader objects only need the subset the reference touches —
  `VideoReader(path, num_threads=)`, `len()`, `.next()`, `.get_avg_fps()`, `.get_frame_timestamp(0)`,
  `._frame_pts[:, 1]`, `.get_batch(idxs).asnumpy()` (video_process_patch.py:40,47,50-51,79,110-112,123,146).
  `SyntheticVideoReader` provides it for `synthetic://` paths (BASELINE configs are synthetic clips) and
  `Cv2VideoReader` for real files (OpenCV is installed).
* `qwen_vl_utils.vision_process` (absent, unpinned in REF/livecc-utils/pyproject.toml:9): the constants
  and the two integer helpers `smart_resize` / `smart_nframes` are restated from the published package
  [unverifiable offline]; env overrides keep the reference's names (video_process_patch.py:10-14).
"""
from __future__ import annotations

import math
import os
import re

import numpy as np
import torch

from ..processing import smart_resize as _hf_smart_resize

# ---- qwen_vl_utils.vision_process constants, with LiveCC's overrides (video_process_patch.py:10-14) ----
os.environ["FORCE_QWENVL_VIDEO_READER"] = "decord+"
os.environ["VIDEO_MAX_PIXELS"] = str(int(os.environ.get("VIDEO_MAX_PIXELS", 24576 * 28 * 28)))
FORCE_QWENVL_VIDEO_READER = "decord+"
IMAGE_FACTOR = 28
FRAME_FACTOR = 2
FPS = 2.0
FPS_MIN_FRAMES = 4
VIDEO_MAX_PIXELS = 768 * 28 * 28
VIDEO_TOTAL_PIXELS = int(float(os.environ["VIDEO_MAX_PIXELS"]))
VIDEO_MIN_PIXELS = int(os.environ.get("VIDEO_MIN_PIXELS", 100 * 28 * 28))
FPS_MAX_FRAMES = int(os.environ.get("FPS_MAX_FRAMES", 480))


def _round_by_factor(x, f):
    return round(x / f) * f


def _ceil_by_factor(x, f):
    return math.ceil(x / f) * f


def _floor_by_factor(x, f):
    return math.floor(x / f) * f


def smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=56 * 56, max_pixels=14 * 14 * 4 * 1280):
    return _hf_smart_resize(height, width, factor=factor, min_pixels=min_pixels, max_pixels=max_pixels)


def smart_nframes(ele: dict, total_frames: int, video_fps: float) -> int:
    if "nframes" in ele:
        nframes = _round_by_factor(ele["nframes"], FRAME_FACTOR)
    else:
        fps = ele.get("fps", FPS)
        min_frames = _ceil_by_factor(ele.get("min_frames", FPS_MIN_FRAMES), FRAME_FACTOR)
        max_frames = _floor_by_factor(ele.get("max_frames", min(FPS_MAX_FRAMES, total_frames)), FRAME_FACTOR)
        nframes = total_frames / video_fps * fps
        nframes = min(min(max(nframes, min_frames), max_frames), total_frames)
        nframes = _floor_by_factor(nframes, FRAME_FACTOR)
    if not (FRAME_FACTOR <= nframes <= total_frames):
        raise ValueError(f"nframes should in interval [{FRAME_FACTOR}, {total_frames}], but got {nframes}.")
    return int(nframes)


# ---- readers (decord subset) -------------------------------------------------------------------
class _Batch:
    def __init__(self, arr):
        self._arr = arr

    def asnumpy(self):
        return self._arr


class SyntheticVideoReader:
    """`synthetic://<frames>x<height>x<width>@<fps>?seed=<n>`: deterministic uint8 THWC frames.
    Frames are moving crops of a low-pass noise canvas, so that attention is not degenerate."""

    def __init__(self, path: str, num_threads: int = 0):
        m = re.match(r"synthetic://(\d+)x(\d+)x(\d+)@([\d.]+)(?:\?seed=(\d+))?$", path)
        if not m:
            raise ValueError(f"bad synthetic video spec {path!r}")
        self.n, self.h, self.w = int(m.group(1)), int(m.group(2)), int(m.group(3))
        self.fps = float(m.group(4))
        self.seed = int(m.group(5) or 0)
        self._frame_pts = None
        self._cursor = 0

    def __len__(self):
        return self.n

    def get_avg_fps(self):
        return self.fps

    def get_frame_timestamp(self, idx):
        if self._frame_pts is None:
            start = np.arange(self.n, dtype=np.float64) / self.fps
            self._frame_pts = np.stack([start, start + 1.0 / self.fps], axis=1)
        return self._frame_pts[idx]

    _BASE_CACHE = {}

    def _base(self) -> torch.Tensor:
        """One smooth noise canvas per (size, seed), cached per process; frames are moving crops of it, so
        producing a frame costs a 0.6 MB copy (the synthetic source must not dominate the ingest time)."""
        key = (self.h, self.w, self.seed)
        base = self._BASE_CACHE.get(key)
        if base is None:
            g = torch.Generator().manual_seed(self.seed * 1000003 + 17)
            hh, ww = self.h + 64, self.w + 64
            low = torch.rand((1, 3, max(2, hh // 16), max(2, ww // 16)), generator=g)
            img = torch.nn.functional.interpolate(low, size=(hh, ww), mode="bilinear", align_corners=False)[0]
            base = (img * 200 + 25).clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()  # HWC
            if len(self._BASE_CACHE) > 64:
                self._BASE_CACHE.clear()
            self._BASE_CACHE[key] = base
        return base

    def _frame(self, i: int) -> np.ndarray:
        base = self._base()
        dy = (int(i) * 7) % 64
        dx = (int(i) * 13) % 64
        return base[dy:dy + self.h, dx:dx + self.w].numpy().copy()

    def next(self):
        f = self._frame(self._cursor)
        self._cursor += 1
        return f

    def get_batch(self, idxs):
        return _Batch(np.stack([self._frame(i) for i in idxs], axis=0) if len(idxs) else
                      np.zeros((0, self.h, self.w, 3), np.uint8))


class Cv2VideoReader:
    """OpenCV-backed stand-in for decord.VideoReader (real mp4 files, e.g. REF/demo/sources/*.mp4)."""

    def __init__(self, path: str, num_threads: int = 0):
        import cv2

        self._cv2 = cv2
        self.cap = cv2.VideoCapture(path)
        if not self.cap.isOpened():
            raise ValueError(f"video_path {path} not found")
        self.n = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.fps = float(self.cap.get(cv2.CAP_PROP_FPS))
        self._frame_pts = None

    def __len__(self):
        return self.n

    def get_avg_fps(self):
        return self.fps

    def get_frame_timestamp(self, idx):
        if self._frame_pts is None:
            start = np.arange(self.n, dtype=np.float64) / self.fps
            self._frame_pts = np.stack([start, start + 1.0 / self.fps], axis=1)
        return self._frame_pts[idx]

    def _read_at(self, i):
        self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, int(i))
        ok, bgr = self.cap.read()
        if not ok:
            raise IndexError(i)
        return self._cv2.cvtColor(bgr, self._cv2.COLOR_BGR2RGB)

    def next(self):
        ok, bgr = self.cap.read()
        if not ok:
            raise StopIteration
        return self._cv2.cvtColor(bgr, self._cv2.COLOR_BGR2RGB)

    def get_batch(self, idxs):
        return _Batch(np.stack([self._read_at(i) for i in idxs], axis=0))


def _open_reader(video_path: str, num_threads: int = 0):
    if isinstance(video_path, str) and video_path.startswith("synthetic://"):
        return SyntheticVideoReader(video_path, num_threads)
    return Cv2VideoReader(video_path, num_threads)


def _resize_bicubic_antialias(video: torch.Tensor, size):
    """transforms.functional.resize(video, size, BICUBIC, antialias=True) (video_process_patch.py:101-106,150-155)."""
    if tuple(video.shape[-2:]) == tuple(size):
        return video
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF

    return TF.resize(video, list(size), interpolation=InterpolationMode.BICUBIC, antialias=True)


# ---- the four public functions ------------------------------------------------------------------
def _read_video_decord_plus(ele: dict, strict_fps: bool = False, drop_last: bool = True, return_pts: bool = False):
    """video_process_patch.py:24-83."""
    video_path = ele["video"]
    if isinstance(video_path, str) and (video_path.startswith("synthetic://") or os.path.exists(video_path)):
        vr = _open_reader(video_path, num_threads=2)
    elif ele.get("remote_loader") is not None:
        vr = _open_reader(ele["remote_loader"](video_path), num_threads=2)
    else:
        raise ValueError(f"video_path {video_path} not found")
    video_start = ele.get("video_start", None)
    video_end = ele.get("video_end", None)
    video_fps = vr.get_avg_fps()
    clip_idxs, clip_pts = None, None
    if video_start is not None or video_end is not None:
        vr.get_frame_timestamp(0)
        video_pts = vr._frame_pts[:, 1]
        video_start = video_pts[0] if not video_start else video_start
        video_end = video_pts[-1] if not video_end else video_end
        clip_idxs = ((video_start <= video_pts) & (video_pts <= video_end)).nonzero()[0]
        clip_pts = video_pts[clip_idxs]
        total_frames = len(clip_idxs)
    else:
        total_frames = len(vr)
    if not strict_fps:
        nframes = smart_nframes(ele, total_frames=total_frames, video_fps=video_fps)
        nframes_idxs = np.linspace(0, total_frames - 1, nframes).round().astype(int)
        clip_idxs = nframes_idxs if clip_idxs is None else clip_idxs[nframes_idxs]
    else:
        if clip_pts is None:
            vr.get_frame_timestamp(0)
            clip_pts = vr._frame_pts[:, 1]
            clip_idxs = np.arange(len(clip_pts))
        expected_timestamps = np.arange(clip_pts[0], clip_pts[-1] + 1e-6, 1 / FPS)
        if len(expected_timestamps) > FPS_MAX_FRAMES:
            if drop_last:
                expected_timestamps = expected_timestamps[:FPS_MAX_FRAMES]
            else:
                expected_timestamps = expected_timestamps[
                    np.linspace(0, len(expected_timestamps) - 1, FPS_MAX_FRAMES).round().astype(int)]
        expected_idxs_for_clip_pts = (expected_timestamps[:, None] <= clip_pts).argmax(axis=1)
        clip_pts = clip_pts[expected_idxs_for_clip_pts].tolist()
        clip_idxs = clip_idxs[expected_idxs_for_clip_pts].tolist()
        while len(clip_idxs) % FRAME_FACTOR != 0:
            clip_idxs.append(clip_idxs[-1])
            clip_pts.append(clip_pts[-1])
    clip = torch.from_numpy(vr.get_batch(list(clip_idxs)).asnumpy()).permute(0, 3, 1, 2)
    sample_fps = len(clip_idxs) / max(total_frames, 1e-6) * video_fps
    if return_pts:
        return clip, sample_fps, clip_pts
    return clip, sample_fps


def _spatial_resize_video(video: torch.Tensor, nframes: int = None):
    """video_process_patch.py:88-107."""
    if not nframes:
        nframes, _, height, width = video.shape
    else:
        height, width = video.shape[2:]
    max_pixels = max(min(VIDEO_MAX_PIXELS, VIDEO_TOTAL_PIXELS / nframes * FRAME_FACTOR), int(VIDEO_MIN_PIXELS * 1.05))
    resized_height, resized_width = smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=VIDEO_MIN_PIXELS,
                                                 max_pixels=max_pixels)
    return _resize_bicubic_antialias(video, (resized_height, resized_width)).float()


def get_smart_resized_video_reader(video_path: str, max_pixels: int = None):
    """video_process_patch.py:109-124."""
    video_reader = _open_reader(video_path)
    nframes = min(len(video_reader), FPS_MAX_FRAMES)
    height, width, _ = video_reader.next().shape
    if max_pixels is None:
        max_pixels = max(min(VIDEO_MAX_PIXELS, VIDEO_TOTAL_PIXELS / nframes * FRAME_FACTOR), int(VIDEO_MIN_PIXELS * 1.05))
    resized_height, resized_width = smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=VIDEO_MIN_PIXELS,
                                                 max_pixels=max_pixels)
    video_reader = _open_reader(video_path, num_threads=2)
    return video_reader, resized_height, resized_width


def get_smart_resized_clip(video_reader, resized_height: int, resized_width: int, timestamps: torch.Tensor,
                           video_pts: np.ndarray, video_pts_index_from: int = 0):
    """video_process_patch.py:126-156."""
    while len(timestamps) % FRAME_FACTOR != 0:
        timestamps = torch.cat([timestamps, timestamps[-1:] + 1 / FPS])
    clip_idxs = []
    for timestamp in timestamps:
        while video_pts_index_from < len(video_pts) and video_pts[video_pts_index_from] < timestamp:
            video_pts_index_from += 1
        if video_pts_index_from >= len(video_pts):
            break
        clip_idxs.append(video_pts_index_from)
    while len(clip_idxs) % FRAME_FACTOR != 0:
        clip_idxs = clip_idxs[:-1]
        timestamps = timestamps[:-1]
    clip = torch.from_numpy(video_reader.get_batch(clip_idxs).asnumpy()).permute(0, 3, 1, 2)
    if (clip.shape[0] == 3) and (clip.shape[1] == len(clip_idxs)):
        clip = clip.transpose(0, 1)
    clip = _resize_bicubic_antialias(clip, (resized_height, resized_width))
    return clip, timestamps, clip_idxs
