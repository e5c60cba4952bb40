"""B1/B2 of the drop-in boundary: a model object with the surface
`LiveCCDemoInfer` uses on `Qwen2VLForConditionalGeneration` (REF/demo/infer.py:43-50,158,165-174):

    model = LiveCCB200ForConditionalGeneration.from_pretrained(path, torch_dtype="auto", device_map="cuda")
    model.device, model.config.eos_token_id, model.config.video_token_id
    model.prepare_inputs_for_generation = functools.partial(fn, model)     # accepted, semantics are native
    out = model.generate(**inputs, past_key_values=state.get('past_key_values'), return_dict_in_generate=True,
                         do_sample=..., repetition_penalty=..., logits_processor=..., max_new_tokens=16,
                         pad_token_id=model.config.eos_token_id)
    out.sequences  [1, L+n] int64 on device ;  out.past_key_values -> pass back next call

Everything numeric runs in liblivecc_sm100a.so (hand-written sm_100a kernels) through the C ABI; this
file is host orchestration: argument checking, position bookkeeping, page allocation, CUDA-graph
replay of the decode step. There is no CPU or eager fallback: without the library or an sm_100 GPU
construction fails.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _cabi
from .checkpoint import EngineWeights, load_engine_weights, synthetic_tensors
from .config import LiveCCConfig
from .kv_cache import PagedKVCache, PagePool
from .positions import get_rope_index


class ThresholdLogitsProcessor:
    """Same constructor and meaning as REF/demo/infer.py:10-23; evaluated inside the sampling kernel
    (softmax(scores)[token_id] <= base_threshold + step * n_generated  ->  scores[token_id] = -inf)."""

    def __init__(self, token_id: int, base_threshold: float, step: float):
        self.token_id = int(token_id)
        self.base_threshold = float(base_threshold)
        self.step = float(step)
        self.count = 0


@dataclass
class GenerateOutput:
    """The two fields of GenerateDecoderOnlyOutput the reference reads (REF/demo/infer.py:173-175)."""
    sequences: torch.Tensor
    past_key_values: PagedKVCache
    logits: Optional[List[torch.Tensor]] = None  # raw fp32 logits per step (only with output_logits=True)


@dataclass
class GenerationDefaults:
    """generation_config.json of the Qwen2-VL family [public-config]: sampling is top_k = 1, i.e. greedy."""
    do_sample: bool = True
    top_k: int = 1
    top_p: float = 0.001
    temperature: float = 0.01
    repetition_penalty: float = 1.0
    eos_token_id: object = None  # int | [int, int] | None (= config.eos_token_id); HF stops on any id of the list

    @classmethod
    def from_json(cls, d: dict) -> "GenerationDefaults":
        g = cls()
        for k in ("do_sample", "top_k", "top_p", "temperature", "repetition_penalty", "eos_token_id"):
            if k in d and d[k] is not None:
                setattr(g, k, d[k])
        return g


class LiveCCB200ForConditionalGeneration:
    # ------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------
    def __init__(self, config: LiveCCConfig, weights: EngineWeights, device: torch.device):
        config.validate()
        if device.type != "cuda":
            raise _cabi.LiveCCNativeError("LiveCCB200ForConditionalGeneration needs a CUDA (sm_100a) device; "
                                          "there is no CPU path")
        self.config = config
        self.device = device
        self.generation_config = GenerationDefaults()
        self.legacy_4x_positions = False
        self.prepare_inputs_for_generation = None  # assignable (REF/demo/infer.py:50); semantics are built in
        self.weights = weights
        torch.cuda.set_device(device)
        self.ctx = _cabi.Context(device.index if device.index is not None else torch.cuda.current_device())
        t, v = config.text_config, config.vision_config
        # rotary tables, computed exactly like the reference does (mq2vl.py:175-183, 274-279)
        self.text_inv_freq = (1.0 / (t.rope_theta ** (torch.arange(0, t.head_dim, 2, dtype=torch.int64).to(torch.float)
                                                      / t.head_dim))).to(device)
        hd2 = v.head_dim // 2
        self.vit_inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd2, 2, dtype=torch.float) / hd2))).to(device)
        self._native = self._build_native()
        self.pool = PagePool(t.num_hidden_layers, t.num_key_value_heads, device)
        self._cap_patches, self._cap_tokens = 0, 0
        self._ensure_workspace(3072, 1024)
        self._graphs = {}
        self._captured_launches = 0   # launches recorded into graphs (counted by the library at capture, never executed then)
        self._replayed_launches = 0   # kernel nodes executed through graph replays
        self.use_cuda_graph = os.environ.get("LIVECC_B200_NO_GRAPH", "0") != "1"
        self.use_vit_graph = self.use_cuda_graph and os.environ.get("LIVECC_B200_VIT_GRAPH", "1") != "0"
        self._vit_graphs = {}
        # split-KV factor of the decode attention: at most one CTA per SM, at least ~8 KV tiles (512 tokens) per split
        self.max_nsplit = max(1, min(64, (self.ctx.num_sms + t.num_key_value_heads - 1) // t.num_key_value_heads))
        self.nsplit = self.max_nsplit
        self.last_stats = {}
        # per-phase device time of the last generate() (CUDA events on the launch stream) and running totals
        self.phase_ms_total = {"vit": 0.0, "prefill": 0.0, "decode": 0.0, "calls": 0, "decode_steps": 0}
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    @classmethod
    def from_synthetic(cls, config: LiveCCConfig, seed: int = 1234, device="cuda"):
        """Synthetic checkpoint (livecc_b200.checkpoint): no weights exist offline."""
        device = torch.device(device if ":" in str(device) else f"{device}:{torch.cuda.current_device()}")
        w = load_engine_weights(config, synthetic_tensors(config, seed, torch.bfloat16, device, gen_device=device), device)
        return cls(config, w, device)

    @classmethod
    def from_state_dict(cls, config: LiveCCConfig, tensors, device="cuda"):
        device = torch.device(device if ":" in str(device) else f"{device}:{torch.cuda.current_device()}")
        return cls(config, load_engine_weights(config, tensors, device), device)

    @classmethod
    def from_pretrained(cls, model_path: str, torch_dtype="auto", device_map="cuda", attn_implementation=None, **_):
        """Loads an HF Qwen2-VL / LiveCC safetensors checkpoint directory (REF/demo/infer.py:43-47): config.json,
        generation_config.json (EOS list, sampling defaults) and *.safetensors in the pre-5.x (`visual.*`, `model.*`)
        or 5.x (`model.visual.*`, `model.language_model.*`) naming. `attn_implementation` is accepted and ignored
        (attention is always the native kernels); weights are always held in bf16 (`torch_dtype` other than
        "auto"/bf16 is rejected)."""
        from .checkpoint import iter_hf_checkpoint, read_hf_configs

        if torch_dtype not in ("auto", None, torch.bfloat16, "bfloat16"):
            raise NotImplementedError(f"torch_dtype={torch_dtype!r}: the sm_100a kernels compute in bf16 only")
        cfg, gen = read_hf_configs(model_path)
        model = cls.from_state_dict(cfg, iter_hf_checkpoint(model_path), device_map if isinstance(device_map, str) else "cuda")
        model.generation_config = GenerationDefaults.from_json(gen)
        if model.generation_config.do_sample and model.generation_config.top_k != 1:
            import warnings

            warnings.warn(f"generation_config.json has top_k={model.generation_config.top_k}: do_sample=True will be "
                          "rejected (the native sampling kernel is greedy, i.e. top_k = 1)")
        return model

    def _build_native(self) -> _cabi.NativeModel:
        t, v, w = self.config.text_config, self.config.vision_config, self.weights
        cfg = _cabi.ModelConfig(
            vit_depth=v.depth, vit_dim=v.embed_dim, vit_heads=v.num_heads, vit_mlp=v.mlp_dim, patch_dim=v.patch_dim,
            merge=v.spatial_merge_size, vit_out=v.hidden_size, hidden=t.hidden_size, inter=t.intermediate_size,
            layers=t.num_hidden_layers, q_heads=t.num_attention_heads, kv_heads=t.num_key_value_heads,
            vocab=t.vocab_size, rms_eps=t.rms_norm_eps, rope_theta=t.rope_theta, mrope_t=t.mrope_section[0],
            mrope_h=t.mrope_section[1], video_token_id=self.config.video_token_id)
        p = lambda x: x.data_ptr()
        vit_arr = (_cabi.VitBlockW * v.depth)(*[
            _cabi.VitBlockW(p(b.norm1_w), p(b.norm1_b), p(b.norm2_w), p(b.norm2_b), p(b.qkv_w), p(b.qkv_b), p(b.proj_w),
                            p(b.proj_b), p(b.fc1_w), p(b.fc1_b), p(b.fc2_w), p(b.fc2_b)) for b in w.vit_blocks])
        lay_arr = (_cabi.LayerW * t.num_hidden_layers)(*[
            _cabi.LayerW(p(l.ln1_w), p(l.qkv_w), p(l.qkv_b), p(l.o_w), p(l.ln2_w), p(l.gate_up_w), p(l.down_w))
            for l in w.layers])
        mw = _cabi.ModelWeights(
            p(w.patch_w), vit_arr, p(w.merger_ln_w), p(w.merger_ln_b), p(w.merger_fc1_w), p(w.merger_fc1_b),
            p(w.merger_fc2_w), p(w.merger_fc2_b), p(w.embed), lay_arr, p(w.final_norm_w), p(w.lm_head),
            p(self.text_inv_freq), p(self.vit_inv_freq))
        return _cabi.NativeModel(self.ctx, cfg, mw, keepalive=(vit_arr, lay_arr, w))

    def _ensure_workspace(self, patches: int, tokens: int):
        if patches <= self._cap_patches and tokens <= self._cap_tokens:
            return
        self._cap_patches = max(self._cap_patches, patches)
        self._cap_tokens = max(self._cap_tokens, tokens)
        torch.cuda.synchronize(self.device)
        self._native.bind_workspace(self._cap_patches, self._cap_tokens, self.device)
        self._graphs = {}
        self._vit_graphs = {}  # captured against the old workspace pointers

    # fixed-shape ViT passes (the 2-frame streaming chunk, the 6-frame opening chunk) are replayed from a CUDA graph: ~260
    # launches per 1024 patches, many of them a few microseconds long
    VIT_GRAPH_MAX_PATCHES = 3072

    def _vit_replay(self, key, src: torch.Tensor, n_out_rows: int, launch):
        """`launch(static_src, static_out)` is captured once per (kind, shape); later calls copy `src` into the static input
        and replay. Returns the STATIC output buffer (valid until the next pass of the same shape)."""
        ent = self._vit_graphs.get(key)
        if ent is None:
            static_src = torch.empty_like(src)
            static_out = torch.empty((n_out_rows, self.config.vision_config.hidden_size), dtype=torch.bfloat16, device=self.device)
            static_src.copy_(src)
            launch(static_src, static_out)  # eager once: kernel attributes, lazy module loading
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _cabi.launch_count()
            with torch.cuda.graph(graph):
                launch(static_src, static_out)
            graph.lcc_nodes = _cabi.launch_count() - n0
            self._captured_launches += graph.lcc_nodes
            ent = self._vit_graphs[key] = (graph, static_src, static_out)
        graph, static_src, static_out = ent
        static_src.copy_(src, non_blocking=True)
        graph.replay()
        self._replayed_launches += graph.lcc_nodes
        return static_out

    # ------------------------------------------------------------------------------------------
    # HF-surface helpers
    # ------------------------------------------------------------------------------------------
    def eval(self):
        return self

    def to(self, *_, **__):
        return self

    def new_cache(self) -> PagedKVCache:
        return PagedKVCache(self.pool)

    def _sampling(self, repetition_penalty, logits_processor, max_new_tokens) -> _cabi.Sampling:
        thr_token, thr_base, thr_step = -1, 0.0, 0.0
        for proc in logits_processor or []:
            if isinstance(proc, ThresholdLogitsProcessor) or all(hasattr(proc, a) for a in ("token_id", "base_threshold", "step")):
                thr_token, thr_base, thr_step = int(proc.token_id), float(proc.base_threshold), float(proc.step)
            else:
                raise NotImplementedError(f"logits processor {type(proc).__name__} is not supported by the native "
                                          "sampling kernel (supported: ThresholdLogitsProcessor)")
        eos = self.generation_config.eos_token_id
        if eos is None:
            eos = self.config.eos_token_id
        eos = [int(e) for e in (eos if isinstance(eos, (list, tuple)) else [eos])]
        if len(eos) > 2:
            raise NotImplementedError(f"at most two EOS ids are supported by the native sampling kernel, got {eos}")
        return _cabi.Sampling(float(repetition_penalty), thr_token, thr_base, thr_step, eos[0],
                              int(max_new_tokens), 1.0 / float(repetition_penalty), eos[1] if len(eos) > 1 else -1)

    # ------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def get_video_features(self, pixel_values_videos: torch.Tensor, video_grid_thw: torch.Tensor, _static_ok: bool = False) -> torch.Tensor:
        """ViT + merger for the videos of one call (mq2vl.py:1094-1112): [sum t*h*w, 1176] f32 -> [n_tok, H] bf16."""
        grids = video_grid_thw.tolist() if isinstance(video_grid_thw, torch.Tensor) else list(video_grid_thw)
        v = self.config.vision_config
        m2 = v.spatial_merge_size ** 2
        n_rows = sum(t * h * w for t, h, w in grids)
        if pixel_values_videos.shape != (n_rows, v.patch_dim):
            raise ValueError(f"pixel_values_videos has shape {tuple(pixel_values_videos.shape)}, expected {(n_rows, v.patch_dim)}")
        px = pixel_values_videos.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        self._ensure_workspace(max(t * h * w for t, h, w in grids), 0)
        if self.use_vit_graph and len(grids) == 1 and n_rows <= self.VIT_GRAPH_MAX_PATCHES:
            t, h, w = grids[0]
            out = self._vit_replay(("rows", t, h, w), px, n_rows // m2, lambda a, b: self._native.vit_forward(a, t, h, w, b))
            return out if _static_ok else out.clone()
        out = torch.empty((n_rows // m2, v.hidden_size), dtype=torch.bfloat16, device=self.device)
        r0 = 0
        for t, h, w in grids:
            n = t * h * w
            self._native.vit_forward(px[r0:r0 + n], t, h, w, out[r0 // m2:(r0 + n) // m2])
            r0 += n
        return out

    @torch.inference_mode()
    def get_video_features_from_frames(self, frames: torch.Tensor, _static_ok: bool = False) -> torch.Tensor:
        """GPU frame ingest: uint8 [T,3,H,W] (H, W multiples of 28) -> [n_tok, H] bf16; the processor's
        rescale/normalize/patchify and the bf16 cast run fused in front of the patch-embed GEMM. Bit-identical to
        get_video_features(patchify_video(frames))."""
        from .processing import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD

        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[1] != 3:
            raise ValueError("video_frames must be a uint8 tensor [T,3,H,W]")
        T, _, H, W = frames.shape
        if H % 28 or W % 28:
            raise ValueError("frame size must be a multiple of 28 (apply smart_resize first)")
        v = self.config.vision_config
        fr = frames.to(device=self.device, non_blocking=True).contiguous()
        t, h, w = (T + 1) // 2, H // 14, W // 14
        self._ensure_workspace(t * h * w, 0)
        # the processor's fused constants (image_processing_backends.py:301-304): fp32(mean) * (1 / (1/255))
        mean255 = (torch.tensor(OPENAI_CLIP_MEAN) * (1.0 / (1 / 255))).tolist()
        std255 = (torch.tensor(OPENAI_CLIP_STD) * (1.0 / (1 / 255))).tolist()
        n_out = t * h * w // v.spatial_merge_size ** 2
        if self.use_vit_graph and t * h * w <= self.VIT_GRAPH_MAX_PATCHES:
            out = self._vit_replay(("frames", T, H, W), fr, n_out, lambda a, b: self._native.vit_forward_frames(a, mean255, std255, b))
            return out if _static_ok else out.clone()
        out = torch.empty((n_out, v.hidden_size), dtype=torch.bfloat16, device=self.device)
        self._native.vit_forward_frames(fr, mean255, std255, out)
        return out

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor = None, pixel_values_videos: Optional[torch.Tensor] = None,
                 video_grid_thw: Optional[torch.Tensor] = None, past_key_values: Optional[PagedKVCache] = None,
                 return_dict_in_generate: bool = True, do_sample: Optional[bool] = None,
                 repetition_penalty: float = 1.0, logits_processor=None, max_new_tokens: int = 16,
                 pad_token_id: Optional[int] = None, attention_mask=None, mm_token_type_ids=None,
                 pixel_values=None, image_grid_thw=None, output_logits: bool = False,
                 video_frames: Optional[torch.Tensor] = None, _forced_ids: Optional[List[int]] = None, **unsupported):
        """The subset of GenerationMixin.generate that LiveCC's streaming loop uses (see module docstring).
        input_ids: [1, L] int64 = full id history (REF/demo/infer.py:159-160); the cache decides how many
        trailing ids are new (gen/utils.py:3747-3758). Greedy decoding (do_sample=True is accepted only with
        the Qwen2-VL generation defaults, top_k = 1, which is greedy)."""
        if unsupported:
            raise TypeError(f"generate() got unsupported arguments {sorted(unsupported)}")
        self._check_common(pixel_values, image_grid_thw, do_sample, max_new_tokens)
        sp = self._sampling(repetition_penalty, logits_processor, max_new_tokens)
        self._ev[0].record()
        rec = self._begin_stream(input_ids, pixel_values_videos, video_grid_thw, past_key_values, video_frames, sp,
                                 max_new_tokens, slot=0, timed=True)
        cache, st, L = rec["cache"], rec["st"], rec["L"]
        logits_out = [self._raw_logits().clone()] if output_logits else None
        if _forced_ids is not None:
            self._force_token(cache, L, 0, _forced_ids, max_new_tokens)

        self._ev[2].record()
        # ---- decode steps ----
        nsplit = self._pick_nsplit(rec["past"] + rec["S"] + max_new_tokens)
        n_steps = max_new_tokens - 1
        logits_hist = None
        if _forced_ids is not None:
            # teacher forcing (parity tests): eager per-step launches, the host swaps the token between steps
            for i in range(n_steps):
                self._native.decode_steps(st, 1, nsplit, sp)
                if output_logits:
                    logits_out.append(self._raw_logits().clone())
                if i + 1 < len(_forced_ids):
                    self._force_token(cache, L, i + 1, _forced_ids, max_new_tokens)
        elif n_steps > 0:
            if output_logits:
                # production (CUDA-graph replay) path with the logits of every step copied out inside the replay loop:
                # the path whose logits the parity tests read is the path the benchmark times
                logits_hist = torch.empty((n_steps, self.config.text_config.vocab_size), dtype=torch.float32,
                                          device=self.device)
            self._run_decode([cache], [st], sp, n_steps, nsplit, logits_hist)

        self._ev[3].record()
        # ---- one host sync per generate(): read the stream scalars ----
        out = self._finish_stream(rec, cache.scalars.tolist(), logits_out, logits_hist, output_logits)
        ph = [self._ev[i].elapsed_time(self._ev[i + 1]) for i in range(3)]
        self.last_stats = {"prefill_tokens": rec["S"], "generated": rec["n_gen"], "kv_len": cache.seq_len, "vit_ms": ph[0],
                           "prefill_ms": ph[1], "decode_ms": ph[2]}
        tot = self.phase_ms_total
        tot["vit"] += ph[0]; tot["prefill"] += ph[1]; tot["decode"] += ph[2]
        tot["calls"] += 1; tot["decode_steps"] += max(rec["n_gen"] - 1, 0)
        return out if return_dict_in_generate else out.sequences

    @torch.inference_mode()
    def generate_batch(self, requests: List[dict], repetition_penalty: float = 1.0, logits_processor=None,
                       max_new_tokens: int = 16, do_sample: Optional[bool] = None, output_logits: bool = False,
                       pad_token_id: Optional[int] = None) -> List[GenerateOutput]:
        """Multi-stream batching: one generate() for up to 8 independent streams of this model
        (the reference demo admits 5 concurrent sessions on one model object, REF/demo/app.py:178, and serves them one
        after another). Every request is a dict with the per-stream arguments of generate() (`input_ids` [1, L],
        `pixel_values_videos` | `video_frames`, `video_grid_thw`, `past_key_values`). The ViT and the prefill run per
        stream; all decode steps run batched in the persistent decode kernel, which reads every weight byte once per
        step for all streams. Each stream's ids, logits and cache are bit-identical to calling generate() on it alone
        (tests/test_engine_gpu.py::test_batched_decode_equals_sequential)."""
        if not 1 <= len(requests) <= _cabi.MAX_BATCH:
            raise ValueError(f"generate_batch takes 1..{_cabi.MAX_BATCH} streams")
        self._check_common(None, None, do_sample, max_new_tokens)
        sp = self._sampling(repetition_penalty, logits_processor, max_new_tokens)
        allowed = {"input_ids", "pixel_values_videos", "video_grid_thw", "past_key_values", "video_frames",
                   "mm_token_type_ids", "attention_mask"}
        recs = []
        t_ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_ev[0].record()
        for b, rq in enumerate(requests):
            if set(rq) - allowed:
                raise TypeError(f"request {b} has unsupported keys {sorted(set(rq) - allowed)}")
            recs.append(self._begin_stream(rq.get("input_ids"), rq.get("pixel_values_videos"), rq.get("video_grid_thw"),
                                           rq.get("past_key_values"), rq.get("video_frames"), sp, max_new_tokens, slot=b))
        if len({id(r["cache"]) for r in recs}) != len(recs):
            raise ValueError("the same cache object appears in two requests of one batch")
        for r in recs:   # a later stream's page allocation may have moved the pool: take the pointers after all of them
            r["st"] = r["cache"].stream_state()
        V = self.config.text_config.vocab_size
        B = len(recs)
        first_logits = self._raw_logits_all()[:B].clone() if output_logits else None
        t_ev[1].record()
        n_steps = max_new_tokens - 1
        logits_hist = torch.empty((n_steps, B, V), dtype=torch.float32, device=self.device) if output_logits and n_steps else None
        if n_steps > 0:
            self._run_decode([r["cache"] for r in recs], [r["st"] for r in recs], sp, n_steps, 1, logits_hist)
        t_ev[2].record()
        scal = torch.stack([r["cache"].scalars for r in recs]).tolist()   # the one host sync of the call
        outs = []
        for b, r in enumerate(recs):
            lo = [first_logits[b]] if output_logits else None
            outs.append(self._finish_stream(r, scal[b], lo, logits_hist[:, b] if logits_hist is not None else None, output_logits))
        tot = self.phase_ms_total
        tot["prefill"] += t_ev[0].elapsed_time(t_ev[1]); tot["decode"] += t_ev[1].elapsed_time(t_ev[2])
        tot["calls"] += B; tot["decode_steps"] += max(max(r["n_gen"] for r in recs) - 1, 0)
        return outs

    @torch.inference_mode()
    def forward_mcq(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, letter_ids,
                    pixel_values_videos: Optional[torch.Tensor] = None, video_grid_thw: Optional[torch.Tensor] = None,
                    mm_token_type_ids=None):
        """The single-forward multiple-choice scoring of the reference's evaluators
        (REF/evaluation/distributed_mcq_predictor.py:72-105): a LEFT-padded batch (`padding_side='left'`,
        REF/evaluation/videomme/distributed_evaluate_videomme.py:39-42), one forward, the logits of the last position
        restricted to the answer-letter token ids, argmax. No decode loop: ViT + prefill only (BASELINE config #5).

        input_ids / attention_mask: [B, L] (pads on the left, mask 0 there); pixel_values_videos: the processor's rows of
        all samples concatenated; video_grid_thw: [n_videos, 3], one video per sample that contains <|video_pad|> ids,
        in batch order. Returns (pred [B] int64 = index into letter_ids, letter_logits [B, len(letter_ids)] fp32).
        The vision tower runs ONCE over all videos of the batch (frames attend only within themselves, so videos with
        the same patch grid concatenate into one pass); the decoder prefill runs per sample on a fresh cache."""
        if input_ids.dim() != 2 or attention_mask.shape != input_ids.shape:
            raise ValueError("input_ids and attention_mask must both be [B, L]")
        B = input_ids.shape[0]
        letters = torch.as_tensor(list(letter_ids), dtype=torch.long, device=self.device)
        ids_host, mask_host = input_ids.cpu(), attention_mask.cpu().bool()
        if not bool(mask_host[:, -1].all()):
            raise ValueError("forward_mcq expects left padding (the last column must be real tokens)")
        grids = video_grid_thw.tolist() if video_grid_thw is not None else []
        m2 = self.config.vision_config.spatial_merge_size ** 2
        feats = None
        if pixel_values_videos is not None:
            if len({(h, w) for _, h, w in grids}) == 1:   # one ViT pass for the whole batch
                t_all, (h, w) = sum(t for t, _, _ in grids), grids[0][1:]
                feats = self.get_video_features(pixel_values_videos, torch.tensor([[t_all, h, w]]))
            else:
                feats = self.get_video_features(pixel_values_videos, video_grid_thw)
        sp = self._sampling(1.0, None, 1)
        out = torch.empty((B, letters.numel()), dtype=torch.float32, device=self.device)
        caches, vid = [], 0
        row0 = 0
        for b in range(B):
            ids_b = ids_host[b][mask_host[b]].view(1, -1)
            n_vid_tok = int((ids_b == self.config.video_token_id).sum())
            fe = g = None
            if n_vid_tok:
                if vid >= len(grids):
                    raise ValueError("more samples with video placeholders than rows in video_grid_thw")
                t, h, w = grids[vid]
                n = t * h * w // m2
                fe, g = feats[row0:row0 + n], torch.tensor([grids[vid]])
                row0 += n
                vid += 1
            rec = self._begin_stream(ids_b, None, g, None, None, sp, 1, slot=0, video_embeds=fe)
            out[b].copy_(self._raw_logits()[letters])
            caches.append(rec)
        scal = torch.stack([r["cache"].scalars for r in caches]).tolist()   # one host sync for the batch
        try:
            for r, sc in zip(caches, scal):
                self._finish_stream(r, sc, None, None, False)
        finally:
            for r in caches:
                r["cache"].release()
        return out.argmax(dim=-1), out

    # ------------------------------------------------------------------------------------------
    def _check_common(self, pixel_values, image_grid_thw, do_sample, max_new_tokens):
        if pixel_values is not None or image_grid_thw is not None:
            raise NotImplementedError("image inputs are outside the LiveCC streaming path (video only)")
        if do_sample and self.generation_config.top_k != 1:
            raise NotImplementedError("do_sample=True is only supported with top_k=1 (the Qwen2-VL generation default)")
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")

    def _pick_nsplit(self, kv_tokens: int) -> int:
        """Split-KV factor of the per-op decode attention (unused by the persistent kernel, which plans its own items)."""
        kv_tiles = (kv_tokens + 63) // 64
        div = int(os.environ.get("LIVECC_B200_NSPLIT_DIV", "8"))
        want = max(1, (kv_tiles + div - 1) // div)   # >= ~8 KV tiles (512 tokens) per split
        nsplit = 1
        while nsplit < want and nsplit < self.max_nsplit:  # quantised to powers of two: at most 7 captured graphs
            nsplit *= 2
        self.nsplit = min(nsplit, self.max_nsplit)
        return self.nsplit

    def _begin_stream(self, input_ids, pixel_values_videos, video_grid_thw, past_key_values, video_frames, sp,
                      max_new_tokens, slot, timed=False, video_embeds=None):
        """Everything of one stream's turn up to and including the prefill and the first token selection; the first
        token's embedding and logits land in row `slot` of the decode-step buffers."""
        if input_ids is None or input_ids.dim() != 2 or input_ids.shape[0] != 1:
            raise ValueError("input_ids must be [1, L] (one stream per call)")
        cfg = self.config
        cache = past_key_values if past_key_values is not None else self.new_cache()
        if not isinstance(cache, PagedKVCache) or cache.pool is not self.pool:
            raise TypeError("past_key_values must be a PagedKVCache returned by this model's generate()")
        ids_dev = input_ids.to(self.device)
        L = ids_dev.shape[1]
        past = cache.seq_len
        S = L - past
        if S <= 0:
            raise ValueError(f"input_ids has {L} ids but the cache already holds {past}")

        if video_frames is not None and video_grid_thw is None:
            _T, _, _H, _W = video_frames.shape
            video_grid_thw = torch.tensor([[(_T + 1) // 2, _H // 14, _W // 14]])
        # ---- positions (host integers; first turn only needs the ids on the host) ----
        if past == 0 or cache.rope_delta is None:
            ids_host = ids_dev[0].tolist()
            grids = video_grid_thw.tolist() if video_grid_thw is not None else []
            pos3, delta = get_rope_index(ids_host, grids, cfg.video_token_id, cfg.image_token_id,
                                         cfg.vision_config.spatial_merge_size, self.legacy_4x_positions)
            cache.rope_delta = delta
            pos3_dev = pos3.to(torch.int32).to(self.device).contiguous()
        else:
            base = torch.arange(past + cache.rope_delta, past + cache.rope_delta + S, dtype=torch.int32, device=self.device)
            pos3_dev = base.view(1, -1).expand(3, -1).contiguous()

        # ---- vision tower ----
        n_video_expected = -1 if video_embeds is None else video_embeds.shape[0]
        if pixel_values_videos is not None and video_frames is not None:
            raise ValueError("pass either pixel_values_videos (HF processor rows) or video_frames (uint8 frames)")
        if pixel_values_videos is not None:
            if video_grid_thw is None:
                raise ValueError("video_grid_thw is required with pixel_values_videos")
            video_embeds = self.get_video_features(pixel_values_videos, video_grid_thw, _static_ok=True)  # consumed by this prefill
            n_video_expected = video_embeds.shape[0]
        elif video_frames is not None:
            video_embeds = self.get_video_features_from_frames(video_frames, _static_ok=True)
            n_video_expected = video_embeds.shape[0]
        if timed:
            self._ev[1].record()
        # ---- capacity, buffers, device scalars ----
        self._ensure_workspace(0, S)
        cache.ensure_tokens(past + S + max_new_tokens)
        cache.ensure_seq_capacity(L + max_new_tokens + 1)
        cache.seq_buf[:L].copy_(ids_dev[0])
        sc_host = torch.tensor([past + S, past + S + cache.rope_delta, 0, 0, L, 0, 0, 0], dtype=torch.int32)
        cache.scalars.copy_(sc_host, non_blocking=False)
        st = cache.stream_state()
        new_ids = cache.seq_buf[past:L]
        # ---- prefill + first token ----
        self._native.prefill(st, new_ids, pos3_dev, S, past, video_embeds, sp, slot=slot)
        return dict(cache=cache, st=st, L=L, S=S, past=past, n_video_expected=n_video_expected, n_gen=0)

    def _finish_stream(self, rec, sc, logits_out, logits_hist, output_logits) -> GenerateOutput:
        cache, L, past = rec["cache"], rec["L"], rec["past"]
        if sc[_cabi.SC_NATIVE_ERROR]:
            raise _cabi.LiveCCNativeError(f"persistent decode kernel gave up on a bounded wait (code "
                                          f"{sc[_cabi.SC_NATIVE_ERROR]}); the results of this call are invalid")
        n_gen = sc[_cabi.SC_N_GENERATED]
        rec["n_gen"] = n_gen
        if rec["n_video_expected"] >= 0:
            n_video_ids = sc[_cabi.SC_VIDEO_TOKENS]
            if n_video_ids != rec["n_video_expected"]:
                # the reference validates before the forward (mq2vl.py:1169-1175); here the count comes back with the
                # step's scalars, so the stream state is rolled back to what it was before this call: the cache still
                # reports `past` tokens (the pages written behind it are dead) and a first turn forgets its rope_delta
                cache.seq_len = past
                if past == 0:
                    cache.rope_delta = None
                raise ValueError(f"Video features and video tokens do not match, tokens: {n_video_ids}, "
                                 f"features: {rec['n_video_expected']}")
        cache.seq_len = sc[_cabi.SC_KV_LEN]
        sequences = cache.seq_buf[: L + n_gen].clone().view(1, -1)
        if output_logits and logits_out is not None:
            if logits_hist is not None:
                logits_out += list(logits_hist[: max(n_gen - 1, 0)])
            logits_out = logits_out[:n_gen]
        return GenerateOutput(sequences=sequences, past_key_values=cache, logits=logits_out if output_logits else None)

    # ------------------------------------------------------------------------------------------
    def _run_decode(self, caches, sts, sp: _cabi.Sampling, n_steps: int, nsplit: int, logits_hist=None):
        """n_steps decode steps for one stream (caches = [cache]) or a batch (<= 8 caches, persistent kernel)."""
        B = len(caches)

        def launch(n):
            if B == 1:
                self._native.decode_steps(sts[0], n, nsplit, sp)
            else:
                self._native.decode_batch(sts, n, sp)

        def copy_logits(i):
            if logits_hist is not None:  # a finished stream replays no-ops and leaves its last logits in place
                logits_hist[i].copy_(self._raw_logits() if B == 1 else self._raw_logits_all()[:B])

        if not self.use_cuda_graph:
            if logits_hist is None:
                launch(n_steps)
            else:
                for i in range(n_steps):
                    launch(1)
                    copy_logits(i)
            return
        key = (tuple(c.graph_key() for c in caches), self._native.workspace.data_ptr(), nsplit, sp.repetition_penalty,
               sp.thr_token, sp.thr_base, sp.thr_step, sp.eos_token_id, sp.eos_token_id2, sp.max_new_tokens)
        g = self._graphs.get(key)
        if g is None:
            # capture ONE decode step (every layer + lm_head + token selection); replay it n_steps times.
            # All step-varying state (kv_len, position, token, finished flag) lives in device memory.
            if len(self._graphs) > 32:
                self._graphs.clear()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            g = torch.cuda.CUDAGraph()
            n0 = _cabi.launch_count()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    launch(1)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g.lcc_nodes = _cabi.launch_count() - n0
            self._captured_launches += g.lcc_nodes
            self._graphs[key] = g
        for i in range(n_steps):
            g.replay()
            self._replayed_launches += g.lcc_nodes
            copy_logits(i)

    def kernel_launches(self) -> int:
        """Kernels of liblivecc_sm100a.so executed so far in this process (counted at the launch sites; launches
        recorded into CUDA graphs are counted per replay). NB: process-wide for the eager part."""
        return _cabi.launch_count() - self._captured_launches + self._replayed_launches

    def _raw_logits(self) -> torch.Tensor:
        off = self._native.logits_offset
        V = self.config.text_config.vocab_size
        return self._native.workspace[off:off + 4 * V].view(torch.float32)

    def _raw_logits_all(self) -> torch.Tensor:
        """[8, V] raw logits of the decode-step slots (row b = stream b of a batched step)."""
        off = self._native.logits_offset
        V = self.config.text_config.vocab_size
        return self._native.workspace[off:off + 4 * V * _cabi.MAX_BATCH].view(torch.float32).view(_cabi.MAX_BATCH, V)

    def _force_token(self, cache: PagedKVCache, L: int, step: int, forced: List[int], max_new_tokens: int):
        """Teacher forcing (parity tests): replace the token just selected by the oracle's token."""
        tok = int(forced[step])
        cache.seq_buf[L + step] = tok
        off = self._native.decode_hidden_offset
        H = self.config.text_config.hidden_size
        self._native.workspace[off:off + 2 * H].view(torch.bfloat16).copy_(self.weights.embed[tok])
        cache.scalars[_cabi.SC_LAST_TOKEN] = tok
        done = tok == self.config.eos_token_id or step + 1 >= max_new_tokens or step + 1 >= len(forced)
        cache.scalars[_cabi.SC_FINISHED] = 1 if done else 0
