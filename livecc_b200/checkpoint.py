"""Checkpoints: HF-named state dicts -> engine weight layout, plus the synthetic checkpoint.

There is no network and no LiveCC checkpoint in the build environment, so parity and benchmarks use a
*synthetic checkpoint*: every tensor of the HF parameter list is filled by a counter-based hash
(`hash_uniform`) that is bit-identical on CPU and CUDA, so the HF oracle and this engine can be given
exactly the same weights on any box without shipping 16 GB of fixtures.

Parameter names follow transformers 5.5.0 `Qwen2VLForConditionalGeneration.state_dict()`
(`model.visual.*`, `model.language_model.*`, `lm_head.weight`; mq2vl.py:287-337,508-545,828-860,1310-1330).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Tuple

import torch

from .config import LiveCCConfig

# --------------------------------------------------------------------------------------------
# Counter-based uniform generator (fmix32 of the element index), identical on every device.
# --------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _mul32(x: torch.Tensor, c: int) -> torch.Tensor:
    """(x * c) mod 2^32 for 0 <= x < 2^32 without ever overflowing int64."""
    lo, hi = c & 0xFFFF, c >> 16
    return (x * lo + (((x * hi) & 0xFFFF) << 16)) & _M32


def hash_uniform(numel: int, seed: int, device, chunk: int = 1 << 24) -> torch.Tensor:
    """float32 tensor of `numel` values in [-0.5, 0.5), a pure function of (index, seed)."""
    out = torch.empty(numel, dtype=torch.float32, device=device)
    seed_mix = (seed * 0x9E3779B1 + 0x7F4A7C15) & _M32
    for start in range(0, numel, chunk):
        n = min(chunk, numel - start)
        i = torch.arange(start, start + n, dtype=torch.int64, device=device)
        x = ((i & _M32) ^ (i >> 32) ^ seed_mix) & _M32
        x = x ^ (x >> 16)
        x = _mul32(x, 0x85EBCA6B)
        x = x ^ (x >> 13)
        x = _mul32(x, 0xC2B2AE35)
        x = x ^ (x >> 16)
        out[start:start + n] = (x >> 8).to(torch.float32) * (1.0 / 16777216.0) - 0.5
    return out


# --------------------------------------------------------------------------------------------
# HF parameter list
# --------------------------------------------------------------------------------------------
def hf_param_specs(cfg: LiveCCConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) in a fixed order; kind in {'w','b','norm_w','norm_b'}."""
    t, v = cfg.text_config, cfg.vision_config
    specs: List[Tuple[str, Tuple[int, ...], str]] = []
    V = "model.visual."
    specs.append((V + "patch_embed.proj.weight",
                  (v.embed_dim, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size), "w"))
    for i in range(v.depth):
        B = f"{V}blocks.{i}."
        specs += [
            (B + "norm1.weight", (v.embed_dim,), "norm_w"), (B + "norm1.bias", (v.embed_dim,), "norm_b"),
            (B + "norm2.weight", (v.embed_dim,), "norm_w"), (B + "norm2.bias", (v.embed_dim,), "norm_b"),
            (B + "attn.qkv.weight", (3 * v.embed_dim, v.embed_dim), "w"), (B + "attn.qkv.bias", (3 * v.embed_dim,), "b"),
            (B + "attn.proj.weight", (v.embed_dim, v.embed_dim), "w"), (B + "attn.proj.bias", (v.embed_dim,), "b"),
            (B + "mlp.fc1.weight", (v.mlp_dim, v.embed_dim), "w"), (B + "mlp.fc1.bias", (v.mlp_dim,), "b"),
            (B + "mlp.fc2.weight", (v.embed_dim, v.mlp_dim), "w"), (B + "mlp.fc2.bias", (v.embed_dim,), "b"),
        ]
    md = v.embed_dim * v.spatial_merge_size ** 2
    specs += [
        (V + "merger.ln_q.weight", (v.embed_dim,), "norm_w"), (V + "merger.ln_q.bias", (v.embed_dim,), "norm_b"),
        (V + "merger.mlp.0.weight", (md, md), "w"), (V + "merger.mlp.0.bias", (md,), "b"),
        (V + "merger.mlp.2.weight", (v.hidden_size, md), "w"), (V + "merger.mlp.2.bias", (v.hidden_size,), "b"),
    ]
    L = "model.language_model."
    H, I = t.hidden_size, t.intermediate_size
    kvd = t.num_key_value_heads * t.head_dim
    specs.append((L + "embed_tokens.weight", (t.vocab_size, H), "w"))
    for i in range(t.num_hidden_layers):
        B = f"{L}layers.{i}."
        specs += [
            (B + "input_layernorm.weight", (H,), "norm_w"),
            (B + "self_attn.q_proj.weight", (H, H), "w"), (B + "self_attn.q_proj.bias", (H,), "b"),
            (B + "self_attn.k_proj.weight", (kvd, H), "w"), (B + "self_attn.k_proj.bias", (kvd,), "b"),
            (B + "self_attn.v_proj.weight", (kvd, H), "w"), (B + "self_attn.v_proj.bias", (kvd,), "b"),
            (B + "self_attn.o_proj.weight", (H, H), "w"),
            (B + "post_attention_layernorm.weight", (H,), "norm_w"),
            (B + "mlp.gate_proj.weight", (I, H), "w"), (B + "mlp.up_proj.weight", (I, H), "w"),
            (B + "mlp.down_proj.weight", (H, I), "w"),
        ]
    specs.append((L + "norm.weight", (H,), "norm_w"))
    specs.append(("lm_head.weight", (t.vocab_size, H), "w"))
    return specs


def synthetic_tensors(cfg: LiveCCConfig, seed: int = 1234, dtype=torch.bfloat16, device="cpu",
                      gen_device=None) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yields (hf_name, tensor). Weights ~ U(-a, a) with a = sqrt(3)*0.02*sqrt(1024/fan_in) capped
    at 0.035 (std ~0.02 at fan_in <= 1024, variance-preserving beyond); biases U(-0.02, 0.02);
    norm weights 1 + U(-0.1, 0.1); norm biases U(-0.02, 0.02). Values are rounded once to `dtype`."""
    gen_device = gen_device or device
    for idx, (name, shape, kind) in enumerate(hf_param_specs(cfg)):
        numel = 1
        for s in shape:
            numel *= s
        u = hash_uniform(numel, seed * 1000003 + idx, gen_device)  # [-0.5, 0.5)
        if kind == "w":
            fan_in = numel // shape[0]
            a = min(0.0346, 0.0346 * (1024.0 / fan_in) ** 0.5)
            x = u * (2.0 * a)
        elif kind == "b":
            x = u * 0.04
        elif kind == "norm_w":
            x = u * 0.2 + 1.0
        else:
            x = u * 0.04
        yield name, x.to(dtype).view(shape).to(device)


def synthetic_state_dict(cfg: LiveCCConfig, seed: int = 1234, dtype=torch.bfloat16, device="cpu",
                         gen_device=None) -> Dict[str, torch.Tensor]:
    return dict(synthetic_tensors(cfg, seed, dtype, device, gen_device))


# --------------------------------------------------------------------------------------------
# Engine layout
# --------------------------------------------------------------------------------------------
def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I,K],[I,K] -> [2I,K] in 32-row groups: 16 gate rows followed by the 16 matching up rows.
    This is the layout LCC_EPI_SWIGLU and the decode gate/up kernel expect."""
    I, K = gate.shape
    assert I % 16 == 0 and up.shape == gate.shape
    g = gate.view(I // 16, 16, K)
    u = up.view(I // 16, 16, K)
    return torch.stack([g, u], dim=1).reshape(2 * I, K).contiguous()


@dataclass
class VitBlockWeights:
    norm1_w: torch.Tensor
    norm1_b: torch.Tensor
    norm2_w: torch.Tensor
    norm2_b: torch.Tensor
    qkv_w: torch.Tensor
    qkv_b: torch.Tensor
    proj_w: torch.Tensor
    proj_b: torch.Tensor
    fc1_w: torch.Tensor
    fc1_b: torch.Tensor
    fc2_w: torch.Tensor
    fc2_b: torch.Tensor


@dataclass
class DecoderLayerWeights:
    ln1_w: torch.Tensor
    qkv_w: torch.Tensor   # [Hq*D + 2*Hkv*D, H]  (q rows, then k rows, then v rows)
    qkv_b: torch.Tensor
    o_w: torch.Tensor
    ln2_w: torch.Tensor
    gate_up_w: torch.Tensor  # [2I, H] interleaved (interleave_gate_up)
    down_w: torch.Tensor


@dataclass
class EngineWeights:
    """bf16 device tensors in the layout the kernels consume."""
    patch_w: torch.Tensor  # [embed, patch_dim_padded]
    vit_blocks: List[VitBlockWeights] = field(default_factory=list)
    merger_ln_w: torch.Tensor = None
    merger_ln_b: torch.Tensor = None
    merger_fc1_w: torch.Tensor = None
    merger_fc1_b: torch.Tensor = None
    merger_fc2_w: torch.Tensor = None
    merger_fc2_b: torch.Tensor = None
    embed: torch.Tensor = None
    layers: List[DecoderLayerWeights] = field(default_factory=list)
    final_norm_w: torch.Tensor = None
    lm_head: torch.Tensor = None

    def nbytes(self) -> int:
        n = 0

        def add(t):
            nonlocal n
            if isinstance(t, torch.Tensor):
                n += t.numel() * t.element_size()

        for k, val in self.__dict__.items():
            if isinstance(val, list):
                for item in val:
                    for t in item.__dict__.values():
                        add(t)
            else:
                add(val)
        return n


def load_engine_weights(cfg: LiveCCConfig, tensors, device, dtype=torch.bfloat16) -> EngineWeights:
    """Builds EngineWeights from an iterable/dict of (hf_name, tensor). Fuses q/k/v and gate/up.
    Streaming: tensors are consumed one at a time so that a 16 GB checkpoint never exists twice."""
    t, v = cfg.text_config, cfg.vision_config
    it = tensors.items() if isinstance(tensors, dict) else tensors
    pending: Dict[str, torch.Tensor] = {}
    w = EngineWeights(patch_w=None)
    vit = [dict() for _ in range(v.depth)]
    dec = [dict() for _ in range(t.num_hidden_layers)]
    misc: Dict[str, torch.Tensor] = {}

    def dev(x):
        return x.to(device=device, dtype=dtype).contiguous()

    for name, ten in it:
        if name.startswith("model.visual.blocks."):
            parts = name.split(".")
            vit[int(parts[3])][".".join(parts[4:])] = dev(ten)
        elif name.startswith("model.language_model.layers."):
            parts = name.split(".")
            li = int(parts[3])
            key = ".".join(parts[4:])
            d = dec[li]
            d[key] = dev(ten)
            # fuse eagerly to bound peak memory
            if all(k in d for k in ("self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight")):
                d["qkv_w"] = torch.cat([d.pop("self_attn.q_proj.weight"), d.pop("self_attn.k_proj.weight"),
                                        d.pop("self_attn.v_proj.weight")], dim=0).contiguous()
            if all(k in d for k in ("self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias")):
                d["qkv_b"] = torch.cat([d.pop("self_attn.q_proj.bias"), d.pop("self_attn.k_proj.bias"),
                                        d.pop("self_attn.v_proj.bias")], dim=0).contiguous()
            if all(k in d for k in ("mlp.gate_proj.weight", "mlp.up_proj.weight")):
                d["gate_up_w"] = interleave_gate_up(d.pop("mlp.gate_proj.weight"), d.pop("mlp.up_proj.weight"))
        else:
            misc[name] = dev(ten)

    V = "model.visual."
    w.patch_w = misc[V + "patch_embed.proj.weight"].reshape(v.embed_dim, v.patch_dim).contiguous()
    for d in vit:
        w.vit_blocks.append(VitBlockWeights(
            norm1_w=d["norm1.weight"], norm1_b=d["norm1.bias"], norm2_w=d["norm2.weight"], norm2_b=d["norm2.bias"],
            qkv_w=d["attn.qkv.weight"], qkv_b=d["attn.qkv.bias"], proj_w=d["attn.proj.weight"],
            proj_b=d["attn.proj.bias"], fc1_w=d["mlp.fc1.weight"], fc1_b=d["mlp.fc1.bias"],
            fc2_w=d["mlp.fc2.weight"], fc2_b=d["mlp.fc2.bias"]))
    w.merger_ln_w, w.merger_ln_b = misc[V + "merger.ln_q.weight"], misc[V + "merger.ln_q.bias"]
    w.merger_fc1_w, w.merger_fc1_b = misc[V + "merger.mlp.0.weight"], misc[V + "merger.mlp.0.bias"]
    w.merger_fc2_w, w.merger_fc2_b = misc[V + "merger.mlp.2.weight"], misc[V + "merger.mlp.2.bias"]
    L = "model.language_model."
    w.embed = misc[L + "embed_tokens.weight"]
    for d in dec:
        w.layers.append(DecoderLayerWeights(
            ln1_w=d["input_layernorm.weight"], qkv_w=d["qkv_w"], qkv_b=d["qkv_b"],
            o_w=d["self_attn.o_proj.weight"], ln2_w=d["post_attention_layernorm.weight"],
            gate_up_w=d["gate_up_w"], down_w=d["mlp.down_proj.weight"]))
    w.final_norm_w = misc[L + "norm.weight"]
    w.lm_head = misc["lm_head.weight"]
    return w
