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


def _sharp_permutation(cfg: LiveCCConfig, device) -> torch.Tensor:
    """pi over the text ids [0, P) (P = first special id): pi(i) = (A*i + B) mod P, A coprime with P."""
    import math

    P = min(cfg.bos_token_id, cfg.eos_token_id, cfg.im_start_token_id, cfg.text_config.vocab_size)
    A = 48271
    while math.gcd(A, P) != 1:
        A += 2
    i = torch.arange(P, dtype=torch.int64, device=device)
    return (i * A + 12345) % P


def sharp_chain(cfg: LiveCCConfig, start: int, n: int) -> List[int]:
    """The id chain a `sharp` checkpoint generates after `start` (ignoring the EOS swap): pi(start), pi^2(start), ..."""
    pi = _sharp_permutation(cfg, "cpu")
    out, t = [], int(start)
    for _ in range(n):
        t = int(pi[t])
        out.append(t)
    return out


def synthetic_tensors(cfg: LiveCCConfig, seed: int = 1234, dtype=torch.bfloat16, device="cpu",
                      gen_device=None, sharp: bool = False, sharp_eos_after: int = 0,
                      only=None) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yields (hf_name, tensor). Weights ~ U(-a, a) with a = sqrt(3)*0.02*sqrt(1024/fan_in) capped
    at 0.035 (std ~0.02 at fan_in <= 1024, variance-preserving beyond); biases U(-0.02, 0.02);
    norm weights 1 + U(-0.1, 0.1); norm biases U(-0.02, 0.02). Values are rounded once to `dtype`.

    sharp=True builds the *sharp* variant used by the id-exactness tests. Random weights give nearly flat logits
    (top-1/top-2 gap of a few bf16 ulps), so greedy ids flip between any two correct bf16 implementations. The sharp
    variant gives the model a confident next-token distribution the way a trained LM has one: the embedding table is
    scaled by max(8, 2*layers) (the token direction then carries about half of the final hidden state's norm at 7B
    depth, the 28 layers the rest) and lm_head[pi(i)] = embed[i] for a fixed permutation pi of the text ids, so the top-1 logit
    leads by several logit units (>> 10x the bf16 tolerance) while every other logit is still produced by the full
    network. The generated ids then follow pi from the last prompt token -- which is the point: id-exactness over
    hundreds of steps checks positions, cache, penalty and stop bookkeeping end to end, while numerics are checked by
    the teacher-forced logit tolerance on the flat checkpoint. sharp_eos_after=k > 0 additionally routes the k-th
    token of the chain that starts at the newline id (the last prompt token of the chat template) to EOS, so streams
    stop early under CUDA-graph replay. `only`: optional set of names to generate (the others are skipped)."""
    gen_device = gen_device or device
    specs = hf_param_specs(cfg)
    embed_idx = next(i for i, sp in enumerate(specs) if sp[0].endswith("embed_tokens.weight"))

    def fill(idx, shape, kind):
        numel = 1
        for s in shape:
            numel *= s
        u = hash_uniform(numel, seed * 1000003 + idx, gen_device)  # [-0.5, 0.5)
        if kind == "w":
            fan_in = numel // shape[0]
            a = min(0.0346, 0.0346 * (1024.0 / fan_in) ** 0.5)
            return u * (2.0 * a)
        if kind == "b":
            return u * 0.04
        if kind == "norm_w":
            return u * 0.2 + 1.0
        return u * 0.04

    for idx, (name, shape, kind) in enumerate(specs):
        if only is not None and name not in only:
            continue
        if sharp and name == "lm_head.weight":
            emb = fill(embed_idx, specs[embed_idx][1], "w").view(specs[embed_idx][1])
            pi = _sharp_permutation(cfg, emb.device)
            x = fill(idx, shape, kind).view(shape)      # special-id rows keep their random fill
            x[pi] = emb[: pi.numel()]                    # lm_head[pi(i)] = embed[i]
            if sharp_eos_after > 0:
                chain = sharp_chain(cfg, cfg.newline_token_id, sharp_eos_after)
                src = chain[-2] if sharp_eos_after > 1 else cfg.newline_token_id   # token whose successor becomes EOS
                nxt = chain[-1]
                x[cfg.eos_token_id] = emb[src]
                x[nxt] = emb[cfg.eos_token_id]
            x = x.reshape(-1)
        else:
            x = fill(idx, shape, kind)
            if sharp and idx == embed_idx:
                x = x * max(8.0, 2.0 * cfg.text_config.num_hidden_layers)
        yield name, x.to(dtype).view(shape).to(device)


def synthetic_state_dict(cfg: LiveCCConfig, seed: int = 1234, dtype=torch.bfloat16, device="cpu",
                         gen_device=None, sharp: bool = False, sharp_eos_after: int = 0) -> Dict[str, torch.Tensor]:
    return dict(synthetic_tensors(cfg, seed, dtype, device, gen_device, sharp, sharp_eos_after))


def sharp_overrides(cfg: LiveCCConfig, seed: int = 1234, dtype=torch.bfloat16, device="cpu", gen_device=None,
                    sharp_eos_after: int = 0) -> Dict[str, torch.Tensor]:
    """The two tensors in which the sharp checkpoint differs from the flat one (embed_tokens, lm_head):
    `dict(flat_state_dict, **sharp_overrides(...))` is the sharp state dict without regenerating 8 B parameters."""
    names = {"model.language_model.embed_tokens.weight", "lm_head.weight"}
    return dict(synthetic_tensors(cfg, seed, dtype, device, gen_device, True, sharp_eos_after, only=names))


# --------------------------------------------------------------------------------------------
# HF checkpoint directories (config.json, generation_config.json, *.safetensors)
# --------------------------------------------------------------------------------------------
def config_from_hf_json(d: dict) -> LiveCCConfig:
    """config.json of a Qwen2-VL / LiveCC checkpoint (flat 4.x layout or nested 5.x `text_config`) -> LiveCCConfig."""
    from .config import TextConfig, VisionConfig

    tc = d.get("text_config", d)
    vc = d.get("vision_config", {})
    rp = tc.get("rope_parameters") or tc.get("rope_scaling") or d.get("rope_scaling") or {}
    text = TextConfig(
        vocab_size=tc.get("vocab_size", 152064), hidden_size=tc.get("hidden_size", 3584),
        intermediate_size=tc.get("intermediate_size", 18944), num_hidden_layers=tc.get("num_hidden_layers", 28),
        num_attention_heads=tc.get("num_attention_heads", 28), num_key_value_heads=tc.get("num_key_value_heads", 4),
        rms_norm_eps=tc.get("rms_norm_eps", 1e-6), rope_theta=rp.get("rope_theta", tc.get("rope_theta", 1e6)),
        mrope_section=tuple(rp.get("mrope_section", (16, 24, 24))))
    vis = VisionConfig(
        depth=vc.get("depth", 32), embed_dim=vc.get("embed_dim", 1280), hidden_size=vc.get("hidden_size", text.hidden_size),
        mlp_ratio=vc.get("mlp_ratio", 4), num_heads=vc.get("num_heads", 16), in_channels=vc.get("in_channels", 3),
        patch_size=vc.get("patch_size", 14), spatial_merge_size=vc.get("spatial_merge_size", 2),
        temporal_patch_size=vc.get("temporal_patch_size", 2))
    eos = d.get("eos_token_id", tc.get("eos_token_id", 151645))
    cfg = LiveCCConfig(
        text_config=text, vision_config=vis, image_token_id=d.get("image_token_id", 151655),
        video_token_id=d.get("video_token_id", 151656), vision_start_token_id=d.get("vision_start_token_id", 151652),
        vision_end_token_id=d.get("vision_end_token_id", 151653), bos_token_id=d.get("bos_token_id", tc.get("bos_token_id", 151643)),
        eos_token_id=eos[0] if isinstance(eos, (list, tuple)) else eos, name=d.get("_name_or_path", "livecc"))
    cfg.tie_word_embeddings = bool(d.get("tie_word_embeddings", tc.get("tie_word_embeddings", False)))
    return cfg


def read_hf_configs(model_path: str) -> Tuple[LiveCCConfig, dict]:
    """(LiveCCConfig, generation_config dict) of a checkpoint directory; LiveCC-7B defaults if a file is absent."""
    import json
    import os

    cfg = LiveCCConfig.livecc_7b()
    p = os.path.join(model_path, "config.json")
    if os.path.exists(p):
        with open(p) as f:
            cfg = config_from_hf_json(json.load(f))
    gen = {}
    p = os.path.join(model_path, "generation_config.json")
    if os.path.exists(p):
        with open(p) as f:
            gen = json.load(f)
    return cfg, gen


def canonical_hf_name(name: str) -> str:
    """Maps pre-5.x parameter names (`visual.*`, `model.layers.*`, `model.embed_tokens.*`, `model.norm.*`) to the
    5.x names (`model.visual.*`, `model.language_model.*`) that hf_param_specs() uses."""
    if name.startswith("visual."):
        return "model." + name
    if name.startswith("model.") and not name.startswith(("model.visual.", "model.language_model.")):
        return "model.language_model." + name[len("model."):]
    return name


def iter_hf_checkpoint(model_path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    """Streams (canonical_name, tensor) from every *.safetensors file of a checkpoint directory."""
    import glob
    import os

    from safetensors import safe_open

    files = sorted(glob.glob(os.path.join(model_path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {model_path!r} (checkpoints cannot be downloaded offline)")
    for f in files:
        with safe_open(f, framework="pt") as sf:
            for name in sf.keys():
                yield canonical_hf_name(name), sf.get_tensor(name)


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

    _seen: List[str] = []
    for name, ten in it:
        _seen.append(name)
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

    # explicit report instead of an opaque KeyError further down
    expected = {n for n, _, _ in hf_param_specs(cfg)}
    seen_names = set(_seen)
    if getattr(cfg, "tie_word_embeddings", False) or ("lm_head.weight" not in seen_names
                                                       and "model.language_model.embed_tokens.weight" in seen_names):
        expected.discard("lm_head.weight")  # tied checkpoints (e.g. the 2B family) carry no lm_head tensor
    missing, unexpected = sorted(expected - seen_names), sorted(seen_names - expected - {"lm_head.weight"})
    if missing:
        raise KeyError(f"checkpoint misses {len(missing)} tensors, e.g. {missing[:4]}"
                       + (f"; unexpected names, e.g. {unexpected[:4]}" if unexpected else ""))
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
    w.lm_head = misc["lm_head.weight"] if "lm_head.weight" in misc else w.embed  # tie_word_embeddings
    return w
