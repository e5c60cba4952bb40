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

    def _crop(self, i: int):
        base = self._base()
        dy = (int(i) * 7) % 64
        dx = (int(i) * 13) % 64
        return base[dy:dy + self.h, dx:dx + self.w]

    def _frame(self, i: int) -> np.ndarray:
        return self._crop(i).numpy().copy()

    def next(self):
        f = self._frame(self._cursor)
        self._cursor += 1
        return f

    def get_batch(self, idxs):
        out = np.empty((len(idxs), self.h, self.w, 3), np.uint8)  # one copy per frame, straight into the batch
        for k, i in enumerate(idxs):
            out[k] = self._crop(i).numpy()
        return _Batch(out)


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


_NATIVE_CTX = {}


def _native_ctx(device: torch.device):
    """One C-ABI context per CUDA device for the ingest kernels (created on first use)."""
    from .. import _cabi

    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _NATIVE_CTX:
        _NATIVE_CTX[idx] = _cabi.Context(idx)
    return _NATIVE_CTX[idx]


def _resize_bicubic_antialias(video: torch.Tensor, size):
    """transforms.functional.resize(video, size, BICUBIC, antialias=True) (video_process_patch.py:101-106,150-155).
    A uint8 CUDA clip is resized by the native kernel (csrc/resize.cu: bit-identical to torchvision's CPU result, both
    passes in one kernel); a host tensor goes through torchvision exactly as in the reference."""
    if tuple(video.shape[-2:]) == tuple(size):
        return video
    if video.is_cuda:
        if video.dtype != torch.uint8:
            raise ValueError("GPU frame ingest resizes uint8 clips (decoded frames); got " + str(video.dtype))
        with torch.cuda.device(video.device):
            return _native_ctx(video.device).resize_bicubic_aa_u8(video.contiguous(), size)
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF

    return TF.resize(video, list(size), interpolation=InterpolationMode.BICUBIC, antialias=True)


# ---- the four public functions ------------------------------------------------------------------
def _first_frame_at_or_after(pts: np.ndarray, stamps: np.ndarray, start: int = 0) -> np.ndarray:
    """For every (ascending) timestamp the index of the first frame, at or after `start`, whose pts is >= the
    timestamp; indices past the end are len(pts). This is the forward scan of video_process_patch.py:137-142
    (the cursor never moves backwards and may stay on a frame for several timestamps) done with one
    searchsorted."""
    pts = np.asarray(pts, dtype=np.float64)
    idx = np.searchsorted(pts, np.asarray(stamps, dtype=np.float64), side="left")
    return np.maximum(idx, start)


def _open_for_element(ele: dict):
    path = ele["video"]
    if isinstance(path, str) and (path.startswith("synthetic://") or os.path.exists(path)):
        return _open_reader(path, num_threads=2)
    if ele.get("remote_loader") is not None:
        return _open_reader(ele["remote_loader"](path), num_threads=2)
    raise ValueError(f"video_path {path} not found")


def _read_video_decord_plus(ele: dict, strict_fps: bool = False, drop_last: bool = True, return_pts: bool = False):
    """Reads (a window of) a video as uint8 TCHW (video_process_patch.py:24-83).
    ele: {"video": path, optional "video_start"/"video_end" seconds, "remote_loader", nframes/fps hints}.
    Default: `smart_nframes` frames, evenly spread over the window. strict_fps: one frame per 1/FPS s (first
    frame whose pts reaches the tick), capped at FPS_MAX_FRAMES, padded to a multiple of FRAME_FACTOR by
    repeating the last frame. Returns (clip, sample_fps[, pts])."""
    vr = _open_for_element(ele)
    start, end = ele.get("video_start"), ele.get("video_end")
    native_fps = vr.get_avg_fps()
    window_idx = window_pts = None
    if start is not None or end is not None:
        vr.get_frame_timestamp(0)
        all_pts = vr._frame_pts[:, 1]
        lo = all_pts[0] if not start else start
        hi = all_pts[-1] if not end else end
        window_idx = np.flatnonzero((lo <= all_pts) & (all_pts <= hi))
        window_pts = all_pts[window_idx]
        n_available = len(window_idx)
    else:
        n_available = len(vr)
    if strict_fps:
        if window_pts is None:
            vr.get_frame_timestamp(0)
            window_pts = vr._frame_pts[:, 1]
            window_idx = np.arange(len(window_pts))
        ticks = np.arange(window_pts[0], window_pts[-1] + 1e-6, 1 / FPS)
        if len(ticks) > FPS_MAX_FRAMES:
            ticks = ticks[:FPS_MAX_FRAMES] if drop_last else \
                ticks[np.linspace(0, len(ticks) - 1, FPS_MAX_FRAMES).round().astype(int)]
        pick = np.minimum(_first_frame_at_or_after(window_pts, ticks), len(window_pts) - 1)
        chosen_idx, chosen_pts = window_idx[pick].tolist(), window_pts[pick].tolist()
        pad = -len(chosen_idx) % FRAME_FACTOR
        chosen_idx += chosen_idx[-1:] * pad
        chosen_pts += chosen_pts[-1:] * pad
    else:
        n_wanted = smart_nframes(ele, total_frames=n_available, video_fps=native_fps)
        spread = np.linspace(0, n_available - 1, n_wanted).round().astype(int)
        chosen_idx = spread if window_idx is None else window_idx[spread]
        chosen_pts = window_pts
    clip = torch.from_numpy(vr.get_batch(list(chosen_idx)).asnumpy()).permute(0, 3, 1, 2)  # THWC -> TCHW
    sample_fps = len(chosen_idx) / max(n_available, 1e-6) * native_fps
    return (clip, sample_fps, chosen_pts) if return_pts else (clip, sample_fps)


def _budgeted_max_pixels(nframes: int) -> float:
    """Per-frame pixel budget: total budget spread over frame pairs, clamped (video_process_patch.py:93,115)."""
    return max(min(VIDEO_MAX_PIXELS, VIDEO_TOTAL_PIXELS / nframes * FRAME_FACTOR), int(VIDEO_MIN_PIXELS * 1.05))


def _spatial_resize_video(video: torch.Tensor, nframes: int = None):
    """Bicubic antialiased resize of a TCHW clip to the smart-resized size under the per-frame pixel budget;
    returns float frames like the reference (video_process_patch.py:88-107)."""
    height, width = video.shape[2:]
    target = smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=VIDEO_MIN_PIXELS,
                          max_pixels=_budgeted_max_pixels(nframes or video.shape[0]))
    return _resize_bicubic_antialias(video, target).float()


def get_smart_resized_video_reader(video_path: str, max_pixels: int = None):
    """Opens a reader and fixes the streaming frame size once from the first frame
    (video_process_patch.py:109-124). Returns (reader, resized_height, resized_width)."""
    probe = _open_reader(video_path)
    height, width, _ = probe.next().shape
    if max_pixels is None:
        max_pixels = _budgeted_max_pixels(min(len(probe), FPS_MAX_FRAMES))
    target = smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=VIDEO_MIN_PIXELS, max_pixels=max_pixels)
    return _open_reader(video_path, num_threads=2), target[0], target[1]


def get_smart_resized_clip(video_reader, resized_height: int, resized_width: int, timestamps: torch.Tensor,
                           video_pts: np.ndarray, video_pts_index_from: int = 0, device=None):
    """Frames for the given timestamps (video_process_patch.py:126-156): timestamps are padded to a multiple of
    FRAME_FACTOR by extrapolation, each maps to the first not-yet-passed frame whose pts reaches it, timestamps
    beyond the last frame are dropped and the result is trimmed back to a multiple of FRAME_FACTOR.
    Returns (uint8 TCHW clip resized to (resized_height, resized_width), timestamps, frame indices).
    `device` (extension): a CUDA device moves the DECODED frames there (pinned staging, async copy)
    and resizes them on the GPU; the returned clip lives on that device and is bit-identical to the host result."""
    extra = -len(timestamps) % FRAME_FACTOR
    if extra:
        tail = timestamps[-1] + torch.arange(1, extra + 1, dtype=timestamps.dtype) / FPS
        timestamps = torch.cat([timestamps, tail])
    pts = video_pts.numpy() if isinstance(video_pts, torch.Tensor) else np.asarray(video_pts)
    idx = _first_frame_at_or_after(pts, timestamps.numpy(), video_pts_index_from)
    overrun = idx >= len(pts)
    n_found = int(np.argmax(overrun)) if overrun.any() else len(idx)  # the scan stops at the first overrun
    odd = n_found % FRAME_FACTOR
    clip_idxs = idx[: n_found - odd].tolist()
    if odd:  # the reference trims the timestamps only by the frames dropped for evenness (not to the clip length)
        timestamps = timestamps[:-odd]
    clip = torch.from_numpy(video_reader.get_batch(clip_idxs).asnumpy())
    if device is not None and torch.device(device).type == "cuda" and clip.numel():
        # decoded THWC bytes go up as they are (pinned staging, async copy); the layout change to TCHW and the resize run
        # on the device (on the host the strided TCHW copy alone costs as much as the upload)
        clip = clip.pin_memory().to(device, non_blocking=True)
    clip = clip.permute(0, 3, 1, 2)  # THWC -> TCHW
    if clip.shape[0] == 3 and clip.shape[1] == len(clip_idxs):  # a reader that returns channel-first batches
        clip = clip.transpose(0, 1)
    return _resize_bicubic_antialias(clip, (resized_height, resized_width)), timestamps, clip_idxs
