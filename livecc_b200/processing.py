"""Host-side input preparation for the hot path: frame -> patch rows, prompt -> token ids.

* `smart_resize`, `patchify_video` restate the arithmetic of transformers' Qwen2-VL video processor
  (image_processing_qwen2_vl.py:62-88, video_processing_qwen2_vl.py:192-286,
  image_processing_backends.py:292-331) so that `pixel_values_videos` is bit-identical to what
  REF/demo/infer.py:151-157 feeds the model.
* `StubProcessor` mirrors the call surface of the HF processor that REF/demo/infer.py:147-157,175 uses
  (`apply_chat_template`, `__call__`, `decode`, `.tokenizer`). The real tokenizer/chat-template files
  are not available offline, so text is tokenised by a deterministic word-hash into ids [1000, 100000)
  while every *special* token keeps its real Qwen2-VL id; the `<|video_pad|>` expansion
  (processing_qwen2_vl.py:111-119) and `mm_token_type_ids` (:126-127) follow the reference.
"""
from __future__ import annotations

import math
import re
from typing import List, Optional, Sequence

import torch

OPENAI_CLIP_MEAN = [0.48145466, 0.4578275, 0.40821073]
OPENAI_CLIP_STD = [0.26862954, 0.26130258, 0.27577711]

# Real Qwen2-VL ids; a LiveCCConfig may relocate them (LiveCCConfig.special_token_ids()).
SPECIAL_TOKENS = {
    "<|endoftext|>": 151643,
    "<|im_start|>": 151644,
    "<|im_end|>": 151645,
    "<|vision_start|>": 151652,
    "<|vision_end|>": 151653,
    "<|image_pad|>": 151655,
    "<|video_pad|>": 151656,
}
NEWLINE_ID = 198  # '\n' in the Qwen2 vocabulary


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                 max_pixels: int = 14 * 14 * 4 * 1280):
    """image_processing_qwen2_vl.py:62-88."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(
            f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def patchify_video(clip: torch.Tensor, patch_size: int = 14, temporal_patch_size: int = 2, merge_size: int = 2):
    """uint8 (or float) clip [T,3,H,W] with H,W multiples of 28 -> (f32 [N, 3*2*14*14], grid [[t,h,w]]).

    Row order (t, h/2, w/2, 2, 2), column order (c, tp, 14, 14): video_processing_qwen2_vl.py:240-272.
    Rescale+normalize is the processor's fused form (image_processing_backends.py:301-304,327)."""
    assert clip.dim() == 4 and clip.shape[1] == 3
    T, C, H, W = clip.shape
    if H % (patch_size * merge_size) or W % (patch_size * merge_size):
        raise ValueError("frame size must be a multiple of 28 (apply smart_resize first)")
    rescale_factor = 1 / 255
    mean = torch.tensor(OPENAI_CLIP_MEAN) * (1.0 / rescale_factor)
    std = torch.tensor(OPENAI_CLIP_STD) * (1.0 / rescale_factor)
    x = clip.to(torch.float32)
    x = (x - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    if pad := -T % temporal_patch_size:
        x = torch.cat((x, x[-1:].expand(pad, -1, -1, -1)), dim=0)
    grid_t = x.shape[0] // temporal_patch_size
    grid_h, grid_w = H // patch_size, W // patch_size
    x = x.view(grid_t, temporal_patch_size, C, grid_h // merge_size, merge_size, patch_size,
               grid_w // merge_size, merge_size, patch_size)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    flat = x.reshape(grid_t * grid_h * grid_w, C * temporal_patch_size * patch_size * patch_size)
    return flat.contiguous(), torch.tensor([[grid_t, grid_h, grid_w]], dtype=torch.int64)


class BatchFeature(dict):
    """Minimal stand-in for transformers.BatchFeature: attribute access + .to(device)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        for k, v in list(self.items()):
            if isinstance(v, torch.Tensor):
                self[k] = v.to(device, non_blocking=True)
        return self


class StubTokenizer:
    _split = re.compile("(" + "|".join(re.escape(s) for s in SPECIAL_TOKENS) + r"|\n|\w+|[^\w\s]| +)")

    def __init__(self, special: Optional[dict] = None, newline_id: int = NEWLINE_ID, vocab_size: int = 152064):
        self.special = dict(special or SPECIAL_TOKENS)
        self.newline_id = newline_id
        self.word_lo = 1000
        self.word_n = max(1, min(99000, vocab_size - 1064 - self.word_lo))

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self._split.findall(text):
            if piece in self.special:
                ids.append(self.special[piece])
            elif piece == "\n":
                ids.append(self.newline_id)
            elif piece.strip() == "":
                continue  # spaces attach to the following word in BPE; the stub drops them
            else:
                h = 2166136261
                for b in piece.encode("utf-8"):
                    h = ((h ^ b) * 16777619) & 0xFFFFFFFF
                ids.append(self.word_lo + h % self.word_n)
        return ids

    def __call__(self, text: str):
        return BatchFeature(input_ids=self.encode(text))

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str:
        special = set(self.special.values())
        out = []
        for i in ids:
            i = int(i)
            if i in special:
                if not skip_special_tokens:
                    out.append(f"<|{i}|>")
            elif i == self.newline_id:
                out.append("\n")
            else:
                out.append(f"<{i}>")
        return "".join(out)


class StubProcessor:
    """The subset of AutoProcessor that LiveCCDemoInfer touches (REF/demo/infer.py:48-59,147-160,175)."""

    DEFAULT_SYSTEM = "You are a helpful assistant."

    def __init__(self, config=None, merge_size: int = 2, patch_size: int = 14, temporal_patch_size: int = 2,
                 emit_frames: bool = False):
        """`config`: optional LiveCCConfig supplying (possibly relocated) special ids and vocab size.
        `emit_frames=True` (GPU frame ingest): uint8 clips are passed through as `video_frames`
        (4x fewer H2D bytes than f32 patch rows) and the engine normalises + patchifies them on the device."""
        self.emit_frames = emit_frames
        if config is not None:
            self.tokenizer = StubTokenizer(config.special_token_ids(), config.newline_token_id,
                                           config.text_config.vocab_size)
            merge_size = config.vision_config.spatial_merge_size
            patch_size = config.vision_config.patch_size
            temporal_patch_size = config.vision_config.temporal_patch_size
        else:
            self.tokenizer = StubTokenizer()
        self.merge_size = merge_size
        self.patch_size = patch_size
        self.temporal_patch_size = temporal_patch_size
        self.video_token = "<|video_pad|>"
        self.video_token_id = self.tokenizer.special[self.video_token]
        self.image_token_id = self.tokenizer.special["<|image_pad|>"]

    def apply_chat_template(self, conversation, tokenize: bool = False, add_generation_prompt: bool = False, **_):
        """Qwen2-VL chat template layout: system header (when the first message is not a system one),
        `<|im_start|>role\\n ... <|im_end|>\\n`, videos rendered as vision_start/video_pad/vision_end."""
        if tokenize:
            raise NotImplementedError("StubProcessor.apply_chat_template supports tokenize=False only")
        out = []
        if conversation and conversation[0]["role"] != "system":
            out.append(f"<|im_start|>system\n{self.DEFAULT_SYSTEM}<|im_end|>\n")
        for msg in conversation:
            out.append(f"<|im_start|>{msg['role']}\n")
            content = msg["content"]
            if isinstance(content, str):
                out.append(content)
            else:
                for item in content:
                    if item["type"] == "text":
                        out.append(item["text"])
                    elif item["type"] == "video":
                        out.append("<|vision_start|><|video_pad|><|vision_end|>")
                    elif item["type"] == "image":
                        out.append("<|vision_start|><|image_pad|><|vision_end|>")
            out.append("<|im_end|>\n")
        if add_generation_prompt:
            out.append("<|im_start|>assistant\n")
        return "".join(out)

    def __call__(self, text, images=None, videos: Optional[list] = None, return_tensors: str = "pt",
                 return_attention_mask: bool = True, **_):
        if images is not None:
            raise NotImplementedError("the LiveCC streaming path is video-only")
        if isinstance(text, (list, tuple)):
            if len(text) != 1:
                raise NotImplementedError("batch size 1 only (one stream per call)")
            text = text[0]
        data = {}
        if videos:
            use_frames = self.emit_frames and len(videos) == 1 and videos[0].dtype == torch.uint8
            if use_frames:
                clip = videos[0]
                T, _, H, W = clip.shape
                if H % (self.patch_size * self.merge_size) or W % (self.patch_size * self.merge_size):
                    raise ValueError("frame size must be a multiple of 28 (apply smart_resize first)")
                data["video_frames"] = clip.contiguous()
                data["video_grid_thw"] = torch.tensor([[(T + self.temporal_patch_size - 1) // self.temporal_patch_size,
                                                        H // self.patch_size, W // self.patch_size]], dtype=torch.int64)
            else:
                flats, grids = [], []
                for clip in videos:
                    f, g = patchify_video(clip, self.patch_size, self.temporal_patch_size, self.merge_size)
                    flats.append(f)
                    grids.append(g)
                data["pixel_values_videos"] = torch.cat(flats, dim=0)
                data["video_grid_thw"] = torch.cat(grids, dim=0)
            # processing_qwen2_vl.py:111-119: one <|video_pad|> per merged token. The reference expands the placeholder in the
            # TEXT and tokenizes the result; splicing n pad ids between the tokenized pieces gives the same ids (the pad is a
            # special token: the tokenizer splits on it before anything else) without a regex pass over ~3 KB of pads per chunk.
            merge_len = self.merge_size ** 2
            pieces = text.split(self.video_token)
            if len(pieces) - 1 > len(data["video_grid_thw"]):
                raise ValueError("more <|video_pad|> placeholders in the text than videos")
            id_list = self.tokenizer.encode(pieces[0])
            for idx, piece in enumerate(pieces[1:]):
                n = int(data["video_grid_thw"][idx].prod().item()) // merge_len
                id_list.extend([self.video_token_id] * n)
                id_list.extend(self.tokenizer.encode(piece))
            ids = torch.tensor([id_list], dtype=torch.int64)
        else:
            ids = torch.tensor([self.tokenizer.encode(text)], dtype=torch.int64)
        data["input_ids"] = ids
        if return_attention_mask:
            data["attention_mask"] = torch.ones_like(ids)
        # processing_qwen2_vl.py:126-127 (transformers >= 5): 0 text, 1 image, 2 video
        mm = torch.zeros_like(ids, dtype=torch.int32)
        mm[ids == self.image_token_id] = 1
        mm[ids == self.video_token_id] = 2
        data["mm_token_type_ids"] = mm
        return BatchFeature(data)

    def decode(self, ids, skip_special_tokens: bool = True) -> str:
        if isinstance(ids, torch.Tensor):
            ids = ids.tolist()
        return self.tokenizer.decode(ids, skip_special_tokens=skip_special_tokens)
