"""Executes the UNMODIFIED reference orchestrator (REF/demo/infer.py) under the import swap INTEGRATION.md §A
describes and diffs everything it does against livecc_b200.streaming.LiveCCDemoInfer: every generate() call
(ids, cache length, pixel rows, kwargs incl. the ThresholdLogitsProcessor the canonical caller REF/demo/cli.py:16-19
asks for), every yielded ((start, stop), text) and the `state` contract. The model is a deterministic recording mock
(this is the host-side boundary test for SURVEY.md §8 B1-B3; the arithmetic behind generate() is covered by the
`-m gpu` parity tests). Runs only where /root/reference exists (the build container); the GPU box has no reference."""
import importlib
import importlib.util
import os
import sys
import types
from unittest import mock

import pytest
import torch

from livecc_b200.config import LiveCCConfig
from livecc_b200.processing import StubProcessor

REF_INFER = "/root/reference/demo/infer.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_INFER), reason="reference checkout not present on this box")


class _Cache:
    def __init__(self):
        self.n = 0

    def get_seq_length(self):
        return self.n


class RecordingModel:
    """generate() is a pure function of (call index, max_new_tokens); every argument is recorded."""

    def __init__(self, cfg):
        self.config, self.device, self.calls = cfg, torch.device("cpu"), []

    def generate(self, input_ids=None, past_key_values=None, max_new_tokens=16, **kw):
        cache = past_key_values or _Cache()
        k = len(self.calls)
        procs = kw.get("logits_processor")
        rec = dict(L=input_ids.shape[1], past=cache.n, new=input_ids[0, cache.n:].tolist(), max_new_tokens=max_new_tokens,
                   rows=kw["pixel_values_videos"].shape[0] if kw.get("pixel_values_videos") is not None else 0,
                   px_sum=float(kw["pixel_values_videos"].double().sum()) if kw.get("pixel_values_videos") is not None else 0.0,
                   grid=kw["video_grid_thw"].tolist() if kw.get("video_grid_thw") is not None else None,
                   do_sample=kw.get("do_sample"), repetition_penalty=kw.get("repetition_penalty"),
                   pad_token_id=kw.get("pad_token_id"), return_dict=kw.get("return_dict_in_generate"),
                   procs=None if procs is None else [(type(p).__name__, p.token_id, p.base_threshold, p.step) for p in procs],
                   other=sorted(set(kw) - {"pixel_values_videos", "video_grid_thw", "do_sample", "repetition_penalty",
                                           "pad_token_id", "return_dict_in_generate", "logits_processor"}))
        self.calls.append(rec)
        n = 3 + k % 4  # 3..6 tokens, the last one EOS on even calls (early stop) else a plain id (token budget)
        gen = [2000 + 7 * k + j for j in range(n)]
        if k % 2 == 0:
            gen[-1] = self.config.eos_token_id
        seq = torch.cat([input_ids, torch.tensor([gen], dtype=input_ids.dtype)], 1)
        cache.n = seq.shape[1] - 1
        return types.SimpleNamespace(sequences=seq, past_key_values=cache)


def _load_reference_class(cfg, model):
    """REF/demo/infer.py imported verbatim with the module swap of INTEGRATION.md §A."""
    real_tf = importlib.import_module("transformers")
    ours = importlib.import_module("livecc_b200.livecc_utils")

    class _Model:
        @classmethod
        def from_pretrained(cls, path, **kw):
            model.from_pretrained_kwargs = dict(path=path, **kw)
            return model

    class _Processor:
        @classmethod
        def from_pretrained(cls, path, **kw):
            return StubProcessor(cfg)

    tf = types.ModuleType("transformers")
    tf.Qwen2VLForConditionalGeneration, tf.AutoProcessor = _Model, _Processor
    tf.LogitsProcessor, tf.logging = real_tf.LogitsProcessor, real_tf.logging
    liger, liger_t = types.ModuleType("liger_kernel"), types.ModuleType("liger_kernel.transformers")
    liger_t.apply_liger_kernel_to_qwen2_vl = lambda *a, **k: None
    liger.transformers = liger_t
    qvu = types.ModuleType("qwen_vl_utils")
    qvu.process_vision_info = lambda *a, **k: (None, None)
    stubs = {"transformers": tf, "liger_kernel": liger, "liger_kernel.transformers": liger_t, "qwen_vl_utils": qvu,
             "livecc_utils": ours}
    with mock.patch.dict(sys.modules, stubs):
        spec = importlib.util.spec_from_file_location("ref_demo_infer", REF_INFER)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        infer = mod.LiveCCDemoInfer(model_path="synthetic-checkpoint", device="cpu")
    return mod, infer


def _drive(infer, path, seconds, **kw):
    """REF/demo/cli.py:13-24."""
    state = {"video_path": path}
    outs = []
    for t in range(seconds + 1):
        state["video_timestamp"] = t
        for (s, e), resp, state in infer.live_cc(message="Please describe the video.", state=state,
                                                  max_pixels=384 * 28 * 28, repetition_penalty=1.05, **kw):
            outs.append((s, e, resp, float(state["last_timestamp"]), int(state["last_video_pts_index"]),
                         state["past_ids"][0].tolist(), state["past_key_values"].get_seq_length()))
        if state.get("video_end", False):
            break
    return outs, state


@pytest.mark.parametrize("thr", [dict(streaming_eos_base_threshold=0.0, streaming_eos_threshold_step=0), dict()])
def test_unmodified_reference_orchestrator_equals_streaming_mirror(thr):
    from livecc_b200.streaming import LiveCCDemoInfer

    cfg = LiveCCConfig.small()
    path = "synthetic://330x140x112@30?seed=9"   # 11 s of video, non-square frames
    m_ref, m_own = RecordingModel(cfg), RecordingModel(cfg)
    mod, ref = _load_reference_class(cfg, m_ref)
    assert m_ref.from_pretrained_kwargs["torch_dtype"] == "auto" and m_ref.from_pretrained_kwargs["device_map"] == "cpu"
    assert callable(m_ref.prepare_inputs_for_generation)          # REF/demo/infer.py:50 assigned it on our object
    own = LiveCCDemoInfer(model=m_own, processor=StubProcessor(cfg))
    assert ref.streaming_eos_token_id == own.streaming_eos_token_id
    assert ref.system_prompt_offset == own.system_prompt_offset
    for name in ("fps", "initial_fps_frames", "streaming_fps_frames", "initial_time_interval",
                 "streaming_time_interval", "frame_time_interval"):
        assert getattr(mod.LiveCCDemoInfer, name) == getattr(LiveCCDemoInfer, name)
    outs_ref, st_ref = _drive(ref, path, 12, **thr)
    outs_own, st_own = _drive(own, path, 12, **thr)
    assert len(outs_ref) == len(outs_own) >= 9
    for a, b in zip(outs_ref, outs_own):
        assert a == b
    assert len(m_ref.calls) == len(m_own.calls) == len(outs_ref)
    for a, b in zip(m_ref.calls, m_own.calls):
        pa, pb = a.pop("procs"), b.pop("procs")
        assert a == b
        if thr:
            assert pa is not None and pb is not None and [p[1:] for p in pa] == [p[1:] for p in pb]
            assert pa[0][0] == pb[0][0] == "ThresholdLogitsProcessor"
        else:
            assert pa is None and pb is None
    assert m_ref.calls[0]["max_new_tokens"] == 16 and m_ref.calls[0]["pad_token_id"] == cfg.eos_token_id
    assert bool(st_ref.get("video_end")) == bool(st_own.get("video_end"))
    assert st_ref["message"] == st_own["message"]
    # nothing left to do once the stream ended: both generators are empty and make no further model calls
    st_ref["video_timestamp"] = st_own["video_timestamp"] = 99
    assert list(ref.live_cc(message="", state=st_ref)) == [] and list(own.live_cc(message="", state=st_own)) == []
    assert len(m_ref.calls) == len(m_own.calls)


def test_reference_threshold_processor_is_the_golden_source():
    """The committed golden (tests/golden/threshold_golden.json) is what the reference class produces today."""
    import json

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    try:
        import make_threshold_golden as mk
    finally:
        sys.path.pop(0)
    Proc = mk.reference_class()
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "threshold_golden.json")))
    for c in gold["cases"]:
        proc = Proc(c["token"], c["base"], c["step"])
        scores = torch.tensor(c["scores"])
        for k in range(c["n_calls"]):
            s = proc(torch.zeros((1, 1), dtype=torch.long), scores.clone()[None])[0]
            assert bool(torch.isinf(s[c["token"]])) == c["masked"][k] and int(s.argmax()) == c["argmax"][k]
