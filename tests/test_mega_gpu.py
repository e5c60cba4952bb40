"""The persistent decode-step kernel (csrc/decode_mega.cu) against the per-op kernels it replaces, phase by phase
through the sub-range hook of the C ABI (lcc_decode_mega_debug), then whole steps: batched decoding must give each
stream exactly the ids / logits / cache it gets when decoded alone, and the per-op path (LIVECC_B200_MEGA=0) must agree
within bf16 tolerance."""
import os

import pytest
import torch

from livecc_b200 import _cabi as A
from livecc_b200.checkpoint import synthetic_state_dict
from livecc_b200.config import LiveCCConfig
from livecc_b200.processing import StubProcessor

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _turn_inputs(proc, turn, frames, seed, hw=(112, 112)):
    g = torch.Generator().manual_seed(seed)
    clip = torch.randint(0, 256, (frames, 3, hw[0], hw[1]), generator=g, dtype=torch.uint8)
    t0 = 0.0 if turn == 0 else 3.0 + (turn - 1)
    content = [{"type": "text", "text": f"Time={t0:.1f}-{3.0 + turn:.1f}s"}, {"type": "video", "video": clip}]
    if turn == 0:
        content.append({"type": "text", "text": "Please describe the video."})
    text = proc.apply_chat_template([{"role": "user", "content": content}], tokenize=False, add_generation_prompt=True)
    if turn > 0:
        text = "<|im_end|>\n" + text[text.index("<|im_start|>user"):]
    return proc(text=text, videos=[clip], return_attention_mask=False)


def _mega_engine(cfg, sd):
    """Engine whose one-stream decode also runs the persistent kernel, so that "alone" and "batched" are the same code."""
    from livecc_b200.engine import LiveCCB200ForConditionalGeneration

    os.environ["LIVECC_B200_MEGA"] = "1"
    try:
        return LiveCCB200ForConditionalGeneration.from_state_dict(cfg, sd, DEV)
    finally:
        del os.environ["LIVECC_B200_MEGA"]


@pytest.fixture(scope="module")
def small():
    cfg = LiveCCConfig.small()
    sd = synthetic_state_dict(cfg, dtype=torch.bfloat16, device=DEV, gen_device=DEV)
    return cfg, sd, _mega_engine(cfg, sd)


@pytest.fixture(scope="module")
def wide():
    """LiveCC-7B widths (hidden 3584, inter 18944, vocab 152064, 28:4 heads) with 2 decoder layers and a 1-block ViT: the
    tile counts, K splits and row-block rotation of the real model at a fraction of the memory."""
    from livecc_b200.config import TextConfig, VisionConfig

    cfg = LiveCCConfig(text_config=TextConfig(num_hidden_layers=2), vision_config=VisionConfig(depth=1), name="livecc-7b-wide-l2")
    sd = synthetic_state_dict(cfg, dtype=torch.bfloat16, device=DEV, gen_device=DEV)
    return cfg, sd, _mega_engine(cfg, sd)


def _ws(eng, off, rows, cols, dtype=torch.bfloat16):
    n = rows * cols * (2 if dtype == torch.bfloat16 else 4)
    return eng._native.workspace[off:off + n].view(dtype).view(rows, cols)


REPORT = os.environ.get("MEGA_DEBUG_REPORT") == "1"  # print every comparison and keep going (bring-up aid)
_failed = []


def _close(a, b, ulps=2.0, frac=1e-3, what="", **_):
    """|a - b| <= ulps bf16 ulps of max(|b|, rms(b)): the two paths accumulate in different orders, so elements that
    are small sums of large terms (residual adds, silu(g)*u near zero) differ by an ulp of the TERMS, not of the element."""
    a, b = a.float(), b.float()
    tol = ulps * 2.0 ** -7 * torch.maximum(b.abs(), b.pow(2).mean().sqrt())   # one bf16 ulp is 2^-8..2^-7 of the value
    bad = ((a - b).abs() > tol).float().mean().item()
    if REPORT:
        nan = int(torch.isnan(a).sum())
        print(f"[mega] {what}: bad {bad:.3e} max|d| {(a - b).abs().nan_to_num(1e9).max().item():.4e} nan {nan} "
              f"ref|max| {b.abs().max().item():.3f}", flush=True)
        if bad > frac or nan:
            _failed.append(what)
            wrong = ((a - b).abs() > tol).nonzero().flatten()
            blocks = sorted(set((wrong // 32).tolist()))
            print(f"[mega]    wrong 32-row blocks ({len(blocks)}): {blocks[:40]}", flush=True)
        return
    assert bad <= frac, f"{what}: {bad:.3e} of elements beyond {ulps} bf16 ulps, max abs err {(a - b).abs().max().item():.4e}"


def _prefilled_streams(eng, cfg, specs):
    """specs: per stream a list of frame counts (turns). Returns caches whose last turn was prefilled with slot = b and
    whose first decode input (embedding of the first generated token) sits in row b of the decode buffers."""
    proc = StubProcessor(cfg)
    caches, pasts = [], []
    for b, turns in enumerate(specs):
        cache = past = None
        for t, frames in enumerate(turns[:-1]):
            inp = _turn_inputs(proc, t, frames, 50 * b + t)
            ids = inp.input_ids.to(DEV) if past is None else torch.cat([past, inp.input_ids.to(DEV)], 1)
            o = eng.generate(input_ids=ids, pixel_values_videos=inp.pixel_values_videos.to(DEV), video_grid_thw=inp.video_grid_thw,
                             past_key_values=cache, repetition_penalty=1.05, max_new_tokens=3)
            cache, past = o.past_key_values, o.sequences[:, :-1]
        caches.append(cache)
        pasts.append(past)
    reqs = []
    for b, turns in enumerate(specs):
        t = len(turns) - 1
        inp = _turn_inputs(proc, t, turns[-1], 50 * b + t)
        ids = inp.input_ids.to(DEV) if pasts[b] is None else torch.cat([pasts[b], inp.input_ids.to(DEV)], 1)
        reqs.append(dict(input_ids=ids, pixel_values_videos=inp.pixel_values_videos.to(DEV), video_grid_thw=inp.video_grid_thw,
                         past_key_values=caches[b]))
    outs = eng.generate_batch(reqs, repetition_penalty=1.05, max_new_tokens=1)   # prefill only, slot = b
    caches = [o.past_key_values for o in outs]
    with torch.inference_mode():
        for c in caches:
            c.scalars[A.SC_FINISHED] = 0   # max_new_tokens = 1 finished the call; re-open the stream for the hook below
    torch.cuda.synchronize()
    return caches, outs


@pytest.mark.parametrize("which", ["small", "wide"])
@torch.inference_mode()
def test_phases_match_the_per_op_kernels(which, request, ctx):
    cfg, sd, eng = request.getfixturevalue(which)
    t = cfg.text_config
    H, I, Hq, Hkv, qkv_dim = t.hidden_size, t.intermediate_size, t.num_attention_heads, t.num_key_value_heads, \
        (t.num_attention_heads + 2 * t.num_key_value_heads) * 128
    nm = eng._native
    for specs in ([[6]], [[6, 2], [2], [6, 2, 2]]):
        caches, _ = _prefilled_streams(eng, cfg, specs)
        B = len(caches)
        sts = [c.stream_state() for c in caches]
        h_rows = _ws(eng, nm.decode_hidden_offset, 8, H)
        h0 = h_rows[:B].clone()
        for layer in range(t.num_hidden_layers):
            lw = eng.weights.layers[layer]
            # reference chain with the per-op kernels, one stream at a time
            ref = {k: [] for k in ("qkv", "attn", "h_o", "act", "h_out")}
            for b in range(B):
                c = caches[b]
                hb = h0[b].clone()
                qkv = ctx.gemv_norm_bias(lw.qkv_w, hb, lw.ln1_w, t.rms_norm_eps, lw.qkv_b)
                ref["qkv"].append(qkv.clone())
                attn = ctx.attn_decode(qkv, eng.pool.k[layer], eng.pool.v[layer], c.page_table, c.scalars, eng.text_inv_freq, Hq, Hkv, 4)
                ref["attn"].append(attn.clone())
                ctx.gemv_residual(lw.o_w, attn, hb)
                ref["h_o"].append(hb.clone())
                act = ctx.gemv_norm_swiglu(lw.gate_up_w, hb, lw.ln2_w, t.rms_norm_eps)
                ref["act"].append(act.clone())
                ctx.gemv_residual(lw.down_w, act, hb)
                ref["h_out"].append(hb.clone())
            torch.cuda.synchronize()
            checks = [(1, "qkv", nm.decode_qkv_offset, qkv_dim), (3, "attn", nm.decode_attn_offset, Hq * 128),
                      (7, "h_o", nm.decode_hidden_offset, H), (15, "act", nm.decode_act_offset, I),
                      (31, "h_out", nm.decode_hidden_offset, H)]
            for mask, key, off, cols in checks:
                h_rows[:B].copy_(h0)
                nm.decode_mega_debug(sts, layer, layer + 1, mask, 0)
                torch.cuda.synchronize()
                assert nm.mega_error() == 0, f"persistent kernel flagged error {nm.mega_error()} (layer {layer} mask {mask})"
                got = _ws(eng, off, 8, cols)[:B]
                for b in range(B):
                    # gate rounding flips (1 ulp of the pre-activation) move silu(g)*u by a few ulps of the tensor scale
                    _close(got[b], ref[key][b], ulps=4.0 if key in ("act", "h_out") else 2.0,
                           what=f"B={B} layer {layer} phase {key} stream {b}")
            h0 = h_rows[:B].clone()   # output of the full layer feeds the next one
        # lm_head
        lg_ref = [ctx.gemv_norm_logits(eng.weights.lm_head, h0[b].clone(), eng.weights.final_norm_w, t.rms_norm_eps)[0] for b in range(B)]
        nm.decode_mega_debug(sts, 0, 0, 0, 1)
        torch.cuda.synchronize()
        got = _ws(eng, nm.logits_offset, 8, t.vocab_size, torch.float32)[:B]
        for b in range(B):
            _close(got[b], lg_ref[b], ulps=2.0, what=f"B={B} lm_head stream {b}")
        for c in caches:
            c.release()
    assert not _failed, _failed


def _decode_alone(eng, cfg, specs, max_new):
    proc = StubProcessor(cfg)
    res = []
    for b, turns in enumerate(specs):
        cache = past = None
        for t, frames in enumerate(turns):
            inp = _turn_inputs(proc, t, frames, 50 * b + t)
            ids = inp.input_ids.to(DEV) if past is None else torch.cat([past, inp.input_ids.to(DEV)], 1)
            last = t == len(turns) - 1
            o = eng.generate(input_ids=ids, pixel_values_videos=inp.pixel_values_videos.to(DEV), video_grid_thw=inp.video_grid_thw,
                             past_key_values=cache, repetition_penalty=1.05, max_new_tokens=max_new if last else 3,
                             output_logits=last)
            cache, past = o.past_key_values, o.sequences[:, :-1]
        kv = [tuple(x.clone() for x in cache.gather(l)) for l in range(cfg.text_config.num_hidden_layers)]
        res.append((o.sequences.clone(), torch.stack(o.logits), kv))
        cache.release()
    return res


@pytest.mark.parametrize("which,specs", [("small", [[6, 2], [2], [6, 2, 2], [6]]), ("wide", [[6, 2], [2]]),
                                         ("wide", [[2], [6], [2, 2], [6, 2], [2], [2, 2, 2], [6], [2, 2]])])
def test_batched_decode_equals_sequential(which, specs, request):
    """B streams with different histories decoded together (one persistent kernel per step, CUDA-graph replay) give
    bit-identical ids, logits and KV caches to decoding each stream alone (B = 4 small config; B = 2 and B = 8 at the 7B
    widths)."""
    cfg, sd, eng = request.getfixturevalue(which)
    alone = _decode_alone(eng, cfg, specs, max_new=6)
    proc = StubProcessor(cfg)
    caches, pasts = [], []
    for b, turns in enumerate(specs):
        cache = past = None
        for t, frames in enumerate(turns[:-1]):
            inp = _turn_inputs(proc, t, frames, 50 * b + t)
            ids = inp.input_ids.to(DEV) if past is None else torch.cat([past, inp.input_ids.to(DEV)], 1)
            o = eng.generate(input_ids=ids, pixel_values_videos=inp.pixel_values_videos.to(DEV), video_grid_thw=inp.video_grid_thw,
                             past_key_values=cache, repetition_penalty=1.05, max_new_tokens=3)
            cache, past = o.past_key_values, o.sequences[:, :-1]
        caches.append(cache)
        pasts.append(past)
    reqs = []
    for b, turns in enumerate(specs):
        t = len(turns) - 1
        inp = _turn_inputs(proc, t, turns[-1], 50 * b + t)
        ids = inp.input_ids.to(DEV) if pasts[b] is None else torch.cat([pasts[b], inp.input_ids.to(DEV)], 1)
        reqs.append(dict(input_ids=ids, pixel_values_videos=inp.pixel_values_videos.to(DEV), video_grid_thw=inp.video_grid_thw,
                         past_key_values=caches[b]))
    outs = eng.generate_batch(reqs, repetition_penalty=1.05, max_new_tokens=6, output_logits=True)
    for b, (o, (seq, logits, kv)) in enumerate(zip(outs, alone)):
        assert torch.equal(o.sequences, seq), f"stream {b}: ids differ between batched and single-stream decoding"
        assert torch.equal(torch.stack(o.logits), logits), f"stream {b}: logits differ bitwise"
        for l in range(cfg.text_config.num_hidden_layers):
            k, v = o.past_key_values.gather(l)
            assert torch.equal(k, kv[l][0]) and torch.equal(v, kv[l][1]), f"stream {b} layer {l}: cache differs"
        o.past_key_values.release()


def test_persistent_kernel_vs_per_op_path(small):
    """Same stream through the persistent kernel (`small` engines are built with LIVECC_B200_MEGA=1; every batched step
    uses it) and through the per-op decode kernels (the one-stream default): logits within bf16 tolerance, cache lengths
    equal."""
    from livecc_b200.engine import LiveCCB200ForConditionalGeneration

    cfg, sd, mega = small
    perop = LiveCCB200ForConditionalGeneration.from_state_dict(cfg, sd, DEV)
    proc = StubProcessor(cfg)
    inp = _turn_inputs(proc, 0, 6, 77).to(DEV)
    a = mega.generate(**inp, repetition_penalty=1.05, max_new_tokens=6, output_logits=True)
    gen = a.sequences[0, inp.input_ids.shape[1]:].tolist()
    b = perop.generate(**inp, repetition_penalty=1.05, max_new_tokens=len(gen), output_logits=True, _forced_ids=gen)
    assert a.past_key_values.get_seq_length() == b.past_key_values.get_seq_length()
    worst = max((x - y).abs().max().item() for x, y in zip(a.logits, b.logits))
    assert worst < 0.06, worst
