"""Checkpoint plumbing on the CPU: an HF-named safetensors directory in the pre-5.x and the
5.x naming round-trips into the engine weight layout (q/k/v fused, gate/up interleaved), config.json /
generation_config.json are honoured, tied checkpoints fall back to the embedding table and broken ones fail with a
readable message. The GPU half (from_pretrained -> identical ids) is tests/test_engine_gpu.py::test_from_pretrained_round_trip."""
import json
import os

import pytest
import torch

from livecc_b200.checkpoint import (canonical_hf_name, config_from_hf_json, interleave_gate_up, iter_hf_checkpoint,
                                    load_engine_weights, read_hf_configs, sharp_chain, sharp_overrides, synthetic_state_dict)
from livecc_b200.config import LiveCCConfig, TextConfig, VisionConfig


def tiny_cfg():
    return LiveCCConfig(text_config=TextConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                                               num_attention_heads=2, num_key_value_heads=1),
                        vision_config=VisionConfig(depth=1, embed_dim=160, hidden_size=256, num_heads=2)).with_vocab(512)


def hf_config_json(cfg, nested=True, tie=False):
    t, v = cfg.text_config, cfg.vision_config
    text = dict(vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
                num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
                num_key_value_heads=t.num_key_value_heads, rms_norm_eps=t.rms_norm_eps, tie_word_embeddings=tie)
    rope = dict(rope_theta=t.rope_theta, mrope_section=list(t.mrope_section))
    vis = dict(depth=v.depth, embed_dim=v.embed_dim, hidden_size=v.hidden_size, mlp_ratio=v.mlp_ratio, num_heads=v.num_heads,
               in_channels=3, patch_size=14, spatial_merge_size=2, temporal_patch_size=2)
    top = dict(image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id, vision_start_token_id=cfg.vision_start_token_id,
               vision_end_token_id=cfg.vision_end_token_id, bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id,
               tie_word_embeddings=tie, vision_config=vis)
    if nested:   # transformers 5.x
        top["text_config"] = dict(text, rope_parameters=dict(rope, rope_type="default"))
    else:        # 4.x: flat text fields, rope_scaling
        top.update(text, rope_theta=t.rope_theta, rope_scaling=dict(type="mrope", mrope_section=list(t.mrope_section)))
    return top


def old_name(n):
    if n.startswith("model.visual."):
        return n[len("model."):]
    if n.startswith("model.language_model."):
        return "model." + n[len("model.language_model."):]
    return n


def write_dir(path, cfg, sd, old=False, nested=True, tie=False, shards=2, gen=None):
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    items = [(old_name(k) if old else k, v.contiguous()) for k, v in sd.items() if not (tie and k == "lm_head.weight")]
    per = (len(items) + shards - 1) // shards
    for i in range(shards):
        save_file(dict(items[i * per:(i + 1) * per]), os.path.join(path, f"model-{i + 1:05d}-of-{shards:05d}.safetensors"))
    json.dump(hf_config_json(cfg, nested, tie), open(os.path.join(path, "config.json"), "w"))
    if gen is not None:
        json.dump(gen, open(os.path.join(path, "generation_config.json"), "w"))


def same_weights(a, b):
    for k, va in a.__dict__.items():
        vb = b.__dict__[k]
        if isinstance(va, list):
            assert len(va) == len(vb)
            for x, y in zip(va, vb):
                for kk in x.__dict__:
                    assert torch.equal(x.__dict__[kk], y.__dict__[kk]), (k, kk)
        elif isinstance(va, torch.Tensor):
            assert torch.equal(va, vb), k


@pytest.mark.parametrize("old,nested", [(False, True), (True, False)])
def test_safetensors_round_trip_both_namings(tmp_path, old, nested):
    cfg = tiny_cfg()
    sd = synthetic_state_dict(cfg, dtype=torch.bfloat16)
    ref = load_engine_weights(cfg, sd, "cpu")
    d = str(tmp_path / "ckpt")
    write_dir(d, cfg, sd, old=old, nested=nested, gen=dict(do_sample=True, top_k=1, eos_token_id=[cfg.eos_token_id, cfg.bos_token_id]))
    cfg2, gen = read_hf_configs(d)
    assert cfg2.text_config == cfg.text_config and cfg2.vision_config == cfg.vision_config
    assert (cfg2.video_token_id, cfg2.eos_token_id, cfg2.bos_token_id) == (cfg.video_token_id, cfg.eos_token_id, cfg.bos_token_id)
    assert gen["eos_token_id"] == [cfg.eos_token_id, cfg.bos_token_id]
    names = [n for n, _ in iter_hf_checkpoint(d)]
    assert sorted(names) == sorted(sd)          # canonical 5.x names whatever the file used
    got = load_engine_weights(cfg2, iter_hf_checkpoint(d), "cpu")
    same_weights(ref, got)
    # the engine layout itself: q|k|v rows fused in that order, gate/up interleaved 16/16
    l0 = "model.language_model.layers.0."
    assert torch.equal(got.layers[0].qkv_w, torch.cat([sd[l0 + "self_attn.q_proj.weight"], sd[l0 + "self_attn.k_proj.weight"],
                                                       sd[l0 + "self_attn.v_proj.weight"]]))
    assert torch.equal(got.layers[0].gate_up_w, interleave_gate_up(sd[l0 + "mlp.gate_proj.weight"], sd[l0 + "mlp.up_proj.weight"]))
    assert torch.equal(got.layers[0].gate_up_w[16:32], sd[l0 + "mlp.up_proj.weight"][:16])


def test_tied_embeddings_and_broken_checkpoints(tmp_path):
    cfg = tiny_cfg()
    sd = synthetic_state_dict(cfg, dtype=torch.bfloat16)
    d = str(tmp_path / "tied")
    write_dir(d, cfg, sd, tie=True)
    cfg2, _ = read_hf_configs(d)
    w = load_engine_weights(cfg2, iter_hf_checkpoint(d), "cpu")
    assert w.lm_head is w.embed                                   # tie_word_embeddings: lm_head = embed_tokens
    broken = {k: v for k, v in sd.items() if "layers.1.mlp.down_proj" not in k}
    with pytest.raises(KeyError, match="misses 1 tensors.*down_proj"):
        load_engine_weights(cfg, broken, "cpu")
    with pytest.raises(FileNotFoundError):
        list(iter_hf_checkpoint(str(tmp_path / "nothing-here")))
    assert canonical_hf_name("visual.blocks.0.norm1.weight") == "model.visual.blocks.0.norm1.weight"
    assert canonical_hf_name("model.layers.3.mlp.up_proj.weight") == "model.language_model.layers.3.mlp.up_proj.weight"
    assert canonical_hf_name("model.language_model.norm.weight") == "model.language_model.norm.weight"
    assert canonical_hf_name("lm_head.weight") == "lm_head.weight"
    assert config_from_hf_json({}).text_config == LiveCCConfig.livecc_7b().text_config   # defaults = LiveCC-7B


def test_sharp_checkpoint_construction():
    """sharp variant: lm_head[pi(i)] = embed[i] (before the embedding gain), EOS routed after k chain steps."""
    cfg = LiveCCConfig.small()
    flat = synthetic_state_dict(cfg, dtype=torch.float32)
    ov = sharp_overrides(cfg, dtype=torch.float32, sharp_eos_after=5)
    assert set(ov) == {"model.language_model.embed_tokens.weight", "lm_head.weight"}
    emb = flat["model.language_model.embed_tokens.weight"]
    gain = max(8.0, 2.0 * cfg.text_config.num_hidden_layers)
    assert torch.equal(ov["model.language_model.embed_tokens.weight"], emb * gain)
    chain = sharp_chain(cfg, cfg.newline_token_id, 6)
    assert len(set(chain)) == 6 and all(0 <= c < cfg.bos_token_id for c in chain)
    lm = ov["lm_head.weight"]
    assert torch.equal(lm[chain[0]], emb[cfg.newline_token_id]) and torch.equal(lm[chain[1]], emb[chain[0]])
    assert torch.equal(lm[cfg.eos_token_id], emb[chain[3]])       # the 5th token of the chain from '\n' is EOS
    assert torch.equal(lm[chain[4]], emb[cfg.eos_token_id])
