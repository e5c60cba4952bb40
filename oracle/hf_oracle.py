"""ORACLE (test infrastructure — never imported by the product path).

Drives the *installed* third-party implementation that carries the reference's arithmetic:
`transformers.Qwen2VLForConditionalGeneration` (transformers 5.5.0, un-vendored dependency of
showlab/livecc: REF/demo/infer.py:4,43-47; REF/README.md:25). The reference repo holds no model code and
no tests for this path, so parity is pinned against this model object, built with the
same synthetic checkpoint the engine loads (livecc_b200/checkpoint.py), through a restated
`LiveCCDemoInfer.live_cc` chunk loop (REF/demo/infer.py:105-180).

Version caveat recorded next to every parity number: oracle = transformers 5.5.0 semantics (first-turn
M-RoPE of t>1 grids differs from the 4.5x stack LiveCC shipped with).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch


def to_hf_config(cfg, attn_implementation: str = "eager"):
    """livecc_b200.config.LiveCCConfig -> transformers.Qwen2VLConfig."""
    from transformers import Qwen2VLConfig

    t, v = cfg.text_config, cfg.vision_config
    hf = Qwen2VLConfig(
        text_config=dict(
            vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
            num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
            num_key_value_heads=t.num_key_value_heads, rms_norm_eps=t.rms_norm_eps,
            rope_parameters=dict(rope_type="default", rope_theta=t.rope_theta, mrope_section=list(t.mrope_section)),
            max_position_embeddings=131072, tie_word_embeddings=False,
            bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id, pad_token_id=cfg.pad_token_id,
        ),
        vision_config=dict(
            depth=v.depth, embed_dim=v.embed_dim, hidden_size=v.hidden_size, mlp_ratio=v.mlp_ratio,
            num_heads=v.num_heads, in_channels=v.in_channels, patch_size=v.patch_size,
            spatial_merge_size=v.spatial_merge_size, temporal_patch_size=v.temporal_patch_size,
        ),
        image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
        vision_start_token_id=cfg.vision_start_token_id, vision_end_token_id=cfg.vision_end_token_id,
        tie_word_embeddings=False,
    )
    hf._attn_implementation = attn_implementation
    return hf


_LIGER_APPLIED = [False]


def apply_liger():
    """Oracle variant (B) of SURVEY.md §7: the reference's real GPU path calls
    `liger_kernel.transformers.apply_liger_kernel_to_qwen2_vl()` before the model is built (REF/demo/infer.py:2-3,
    REF/inference.md:14). The patch swaps module-level symbols of transformers for the rest of the process, so a
    test that uses it must run after every variant-(A) test of the same process."""
    from liger_kernel.transformers import apply_liger_kernel_to_qwen2_vl

    apply_liger_kernel_to_qwen2_vl()
    _LIGER_APPLIED[0] = True


def oracle_variant(attn_implementation: str) -> str:
    return f"HF transformers bf16 + {attn_implementation}" + (" + liger_kernel (variant B)" if _LIGER_APPLIED[0] else " (variant A)")


def build_hf_model(cfg, tensors: Iterable[Tuple[str, torch.Tensor]] | Dict[str, torch.Tensor],
                   dtype=torch.float32, device="cpu", attn_implementation: str = "eager", liger: bool = False):
    """Instantiates HF Qwen2-VL the way `from_pretrained(torch_dtype=...)` does (parameters created in
    `dtype`, the fp32 rotary `inv_freq` buffers left in fp32) and loads the given HF-named tensors.
    liger=True: apply_liger() first (process-wide!)."""
    if liger:
        apply_liger()
    from transformers import Qwen2VLForConditionalGeneration

    hf_cfg = to_hf_config(cfg, attn_implementation)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device("meta"):  # skip HF's random init; every parameter is overwritten below
            model = Qwen2VLForConditionalGeneration._from_config(hf_cfg, attn_implementation=attn_implementation)
    finally:
        torch.set_default_dtype(prev)
    model = model.to_empty(device=device)
    params = dict(model.named_parameters())
    items = tensors.items() if isinstance(tensors, dict) else tensors
    seen = set()
    with torch.no_grad():
        for name, ten in items:
            p = params[name]
            assert p.shape == ten.shape, (name, p.shape, ten.shape)
            p.copy_(ten.to(device=p.device, dtype=p.dtype))
            seen.add(name)
    missing = set(params) - seen
    assert not missing, f"synthetic checkpoint misses {sorted(missing)[:5]}"
    # from_pretrained keeps these non-persistent buffers in fp32 (mq2vl.py:274-279,681-683)
    hd = cfg.vision_config.head_dim // 2
    ref = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float) / hd))
    model.model.visual.rotary_pos_emb.inv_freq = ref.to(device)
    rot = model.model.language_model.rotary_emb
    inv_freq, rot.attention_scaling = rot.compute_default_rope_parameters(rot.config, device)
    rot.inv_freq = inv_freq.to(device)
    rot.original_inv_freq = inv_freq.clone().to(device)
    assert model.model.visual.rotary_pos_emb.inv_freq.dtype == torch.float32
    assert model.model.language_model.rotary_emb.inv_freq.dtype == torch.float32
    model.eval()
    model.generation_config.eos_token_id = cfg.eos_token_id
    model.generation_config.pad_token_id = cfg.eos_token_id
    model.generation_config.do_sample = False
    return model


@torch.inference_mode()
def hf_generate_chunk(model, inputs: dict, past_key_values, past_ids: Optional[torch.Tensor],
                      max_new_tokens: int = 16, repetition_penalty: float = 1.05,
                      logits_processor=None, output_logits: bool = False):
    """One iteration of the per-chunk body of live_cc (REF/demo/infer.py:158-174) on the HF model.
    `inputs`: processor output for the new turn (input_ids = new tokens only)."""
    dev = model.device
    new_ids = inputs["input_ids"].to(dev)
    kw = {}
    if inputs.get("pixel_values_videos") is not None:  # text-only turns (history building) carry no pixels
        kw = dict(pixel_values_videos=inputs["pixel_values_videos"].to(dev),
                  video_grid_thw=inputs["video_grid_thw"].to(dev))
    if past_ids is not None:
        input_ids = torch.cat([past_ids, new_ids], dim=1)
    else:
        input_ids = new_ids
    # transformers 5.x wants mm_token_type_ids aligned with the ids it is given (full history)
    mm = torch.zeros_like(input_ids, dtype=torch.int32)
    mm[input_ids == model.config.image_token_id] = 1
    mm[input_ids == model.config.video_token_id] = 2
    kw["mm_token_type_ids"] = mm
    out = model.generate(
        input_ids=input_ids, **kw, past_key_values=past_key_values, return_dict_in_generate=True,
        do_sample=False, repetition_penalty=repetition_penalty, logits_processor=logits_processor,
        max_new_tokens=max_new_tokens, pad_token_id=model.config.eos_token_id if hasattr(model.config, "eos_token_id") and model.config.eos_token_id is not None else model.generation_config.pad_token_id,
        output_logits=output_logits,
    )
    return out, input_ids.size(1)
