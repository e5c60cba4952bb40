"""Model hyper-parameters of the LiveCC hot path (Qwen2-VL architecture).

Field names follow transformers' Qwen2VLConfig / Qwen2VLTextConfig / Qwen2VLVisionConfig
(SP/transformers/models/qwen2_vl/configuration_qwen2_vl.py:31-41,100-101,159-163) so that the
reference-facing code reads `model.config.eos_token_id`, `.video_token_id` exactly as
REF/demo/infer.py:171 and REF/livecc-utils/src/livecc_utils/generation_patch.py:37 do.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class VisionConfig:
    depth: int = 32
    embed_dim: int = 1280
    hidden_size: int = 3584  # merger output = text hidden size
    mlp_ratio: int = 4
    num_heads: int = 16
    in_channels: int = 3
    patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def mlp_dim(self) -> int:
        return self.embed_dim * self.mlp_ratio


@dataclass
class TextConfig:
    vocab_size: int = 152064
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: tuple = (16, 24, 24)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class LiveCCConfig:
    text_config: TextConfig = field(default_factory=TextConfig)
    vision_config: VisionConfig = field(default_factory=VisionConfig)
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653
    bos_token_id: int = 151643
    eos_token_id: int = 151645
    pad_token_id: int = 151643
    im_start_token_id: int = 151644
    newline_token_id: int = 198
    name: str = "livecc-7b"

    def special_token_ids(self) -> dict:
        """Token string -> id for the special tokens the chat template emits."""
        return {
            "<|endoftext|>": self.bos_token_id,
            "<|im_start|>": self.im_start_token_id,
            "<|im_end|>": self.eos_token_id,
            "<|vision_start|>": self.vision_start_token_id,
            "<|vision_end|>": self.vision_end_token_id,
            "<|image_pad|>": self.image_token_id,
            "<|video_pad|>": self.video_token_id,
        }

    # -- factories --------------------------------------------------------------------------
    @staticmethod
    def livecc_7b() -> "LiveCCConfig":
        """LiveCC-7B = Qwen2-VL-7B dims (SURVEY.md §8 header)."""
        return LiveCCConfig()

    @staticmethod
    def small(layers: int = 2, vit_depth: int = 2) -> "LiveCCConfig":
        """Parity-test config: the 7B head geometry (decoder head_dim 128 with GQA 7:1, ViT
        head_dim 80) at a fraction of the width/depth, so that the CPU oracle finishes in seconds."""
        return LiveCCConfig(
            text_config=TextConfig(hidden_size=1792, intermediate_size=4864, num_hidden_layers=layers,
                                   num_attention_heads=14, num_key_value_heads=2),
            vision_config=VisionConfig(depth=vit_depth, embed_dim=320, hidden_size=1792, num_heads=4),
            name=f"livecc-small-l{layers}v{vit_depth}",
        ).with_vocab(16384)

    def with_vocab(self, vocab_size: int) -> "LiveCCConfig":
        """Shrinks the vocabulary (CPU-oracle speed) and relocates the special ids to its top."""
        self.text_config.vocab_size = vocab_size
        top = vocab_size - 32
        self.bos_token_id = self.pad_token_id = top + 0
        self.im_start_token_id = top + 1
        self.eos_token_id = top + 2
        self.vision_start_token_id = top + 9
        self.vision_end_token_id = top + 10
        self.image_token_id = top + 12
        self.video_token_id = top + 13
        return self

    def validate(self) -> None:
        t, v = self.text_config, self.vision_config
        if t.head_dim != 128:
            raise ValueError("decoder kernels are specialised for head_dim 128")
        if v.head_dim != 80:
            raise ValueError("ViT attention kernel is specialised for head_dim 80")
        if sum(t.mrope_section) * 2 != t.head_dim:
            raise ValueError("mrope_section must sum to head_dim/2")
        if v.hidden_size != t.hidden_size:
            raise ValueError("merger output must equal text hidden size")
        if t.intermediate_size % 32 or t.hidden_size % 8 or v.embed_dim % 8:
            raise ValueError("unsupported dims")
