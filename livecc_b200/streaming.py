"""`LiveCCDemoInfer` on the native engine: the streaming orchestrator whose behaviour is specified by
REF/demo/infer.py:25-310 (class constants, the `state` dictionary contract, the generator protocol
`for (start, stop), response, state in infer.live_cc(message=..., state=state, ...)`), so that
REF/demo/cli.py:13-24 drives it unchanged.

The control flow is decomposed into small planning steps (which timestamps are due, which frames they map
to, how the clip splits into the 6-frame opening chunk and 2-frame streaming chunks, how a turn's prompt is
glued to the history) so that each rule of the reference can be cited and tested on its own. Differences,
all outside the arithmetic:
  * the model is `LiveCCB200ForConditionalGeneration` (no liger / flash-attn / HF modeling);
  * the processor may be the offline `StubProcessor` (no tokenizer files exist in this environment);
  * `livecc_utils` is the in-tree implementation (decord / qwen_vl_utils are not installed);
  * per-chunk timings are recorded in `self.timings` for the benchmark.
"""
from __future__ import annotations

import functools
import time

import torch

from .engine import LiveCCB200ForConditionalGeneration, ThresholdLogitsProcessor
from .livecc_utils import (_read_video_decord_plus, _spatial_resize_video, get_smart_resized_clip,
                           get_smart_resized_video_reader, prepare_multiturn_multimodal_inputs_for_generation)
from .processing import StubProcessor

TURN_GLUE = "<|im_end|>\n"  # closes the previous assistant turn when a new user turn is appended to the history


class LiveCCDemoInfer:
    VIDEO_PLAY_END = object()
    VIDEO_PLAY_CONTINUE = object()
    # REF/demo/infer.py:28-33
    fps = 2
    initial_fps_frames = 6
    streaming_fps_frames = 2
    initial_time_interval = initial_fps_frames / fps
    streaming_time_interval = streaming_fps_frames / fps
    frame_time_interval = 1 / fps

    def __init__(self, model_path: str = None, device: str = None, model=None, processor=None, gpu_ingest: bool = None):
        """Builds (or adopts) the model and processor (REF/demo/infer.py:35-59). `model=` / `processor=` reuse
        existing objects (synthetic checkpoint); otherwise `model_path` must be a local HF checkpoint directory.
        `gpu_ingest` (extension): decoded frames that need resizing are uploaded as uint8 and resized on the GPU
        (bit-identical to the host resize); default = on whenever the processor hands uint8 frames to the engine."""
        if model is None:
            device = device or ("cuda" if torch.cuda.is_available() else "cpu")
            model = LiveCCB200ForConditionalGeneration.from_pretrained(model_path, torch_dtype="auto", device_map=device)
        if processor is None:
            processor = self._load_processor(model_path, model)
        self.model, self.processor = model, processor
        # the token after which the streaming-EOS threshold processor looks (infer.py:49)
        self.streaming_eos_token_id = processor.tokenizer(" ...").input_ids[-1]
        # accepted for interface compatibility; the engine implements these semantics natively (infer.py:50)
        model.prepare_inputs_for_generation = functools.partial(prepare_multiturn_multimodal_inputs_for_generation, model)
        probe = processor.apply_chat_template([{"role": "user", "content": [{"type": "text", "text": "livecc"}]}],
                                              tokenize=False)
        self.system_prompt_offset = probe.index("<|im_start|>user")  # later turns drop the system header
        self._cached_video_readers_with_hw = {}
        if gpu_ingest is None:
            gpu_ingest = bool(getattr(processor, "emit_frames", False)) and torch.device(model.device).type == "cuda"
        self.ingest_device = model.device if gpu_ingest else None
        self.timings = []  # one record per chunk: frames, new_tokens, ingest_s, preprocess_s, generate_s, kv_len

    @staticmethod
    def _load_processor(model_path, model):
        """`AutoProcessor.from_pretrained(model_path, use_fast=False)` (REF/demo/infer.py:48). Only a checkpoint directory
        WITHOUT any tokenizer / processor file (weights-only export; nothing can be downloaded offline) falls back to the
        offline StubProcessor, with a warning; a present-but-broken processor raises like the reference would."""
        import os
        import warnings

        names = ("tokenizer_config.json", "tokenizer.json", "vocab.json", "processor_config.json", "preprocessor_config.json")
        if model_path and os.path.isdir(model_path) and not any(os.path.exists(os.path.join(model_path, n)) for n in names):
            warnings.warn(f"{model_path!r} has no tokenizer files: using the offline StubProcessor (synthetic token ids)")
            return StubProcessor(model.config)
        from transformers import AutoProcessor

        return AutoProcessor.from_pretrained(model_path, use_fast=False)

    # ------------------------------------------------------------------------------------------
    # planning steps of live_cc
    # ------------------------------------------------------------------------------------------
    def _stream_source(self, state: dict, max_pixels: int):
        """Opens (once per path) the resized reader and publishes the pts table into `state`
        (infer.py:88-97). Returns (reader, H, W) or None when the stream cannot advance."""
        path = state.get("video_path")
        if not path:
            return None
        if path not in self._cached_video_readers_with_hw:
            entry = get_smart_resized_video_reader(path, max_pixels)
            self._cached_video_readers_with_hw[path] = entry
            entry[0].get_frame_timestamp(0)
            state["video_pts"] = torch.from_numpy(entry[0]._frame_pts[:, 1])
            state["last_video_pts_index"] = -1
        return self._cached_video_readers_with_hw[path] if state.get("video_pts") is not None else None

    def _due_timestamps(self, state: dict):
        """Timestamps (every 0.5 s) between the last processed one and the player position; the first call
        is stretched to the 3 s opening window (infer.py:84,98-111). None = nothing to do."""
        pts = state["video_pts"]
        last = state.get("last_timestamp", -self.frame_time_interval)
        if last + self.frame_time_interval > pts[-1]:
            state["video_end"] = True
            return None
        now = min(state.get("video_timestamp", 0), pts[-1])
        if last < 0:
            now = max(now, self.initial_time_interval)
        if now <= last + self.frame_time_interval:
            return None
        return torch.arange(last + self.frame_time_interval, now, self.frame_time_interval)

    def _chunks(self, clip: torch.Tensor, stamps: torch.Tensor, opening: bool):
        """6 frames for the opening chunk, then 2-frame chunks (infer.py:121-129)."""
        parts = []
        if opening:
            n = self.initial_fps_frames
            parts.append((clip[:n], stamps[:n]))
            clip, stamps = clip[n:], stamps[n:]
        if len(clip) > 0:
            parts.extend(zip(clip.split(self.streaming_fps_frames), stamps.split(self.streaming_fps_frames)))
        return parts

    def _turn_inputs(self, clip, span, query, past_ids):
        """Chat-template text of one user turn (+ the optional query) glued to the history, then the processor
        call and the id concatenation (infer.py:134-160)."""
        content = [{"type": "text", "text": f"Time={span[0]:.1f}-{span[1]:.1f}s"}, {"type": "video", "video": clip}]
        if query:
            content.append({"type": "text", "text": query})
        text = self.processor.apply_chat_template([{"role": "user", "content": content}], tokenize=False,
                                                  add_generation_prompt=True)
        if past_ids is not None:
            text = TURN_GLUE + text[self.system_prompt_offset:]
        inputs = self.processor(text=text, images=None, videos=[clip], return_tensors="pt", return_attention_mask=False)
        inputs.to(self.model.device)
        if past_ids is not None:
            inputs["input_ids"] = torch.cat([past_ids, inputs.input_ids], dim=1)
        return inputs

    # ------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def live_cc(self, message: str, state: dict, max_pixels: int = 384 * 28 * 28,
                default_query: str = "Please describe the video.", do_sample: bool = True,
                repetition_penalty: float = 1.05, streaming_eos_base_threshold: float = None,
                streaming_eos_threshold_step: float = None, hf_spaces: bool = False, max_new_tokens: int = 16,
                **kwargs):
        """Streaming commentary: consumes the frames that became due since the last call and yields one
        ((start, stop), text, state) per chunk (REF/demo/infer.py:62-180). State keys: video_path,
        video_timestamp, last_timestamp, last_video_pts_index, video_pts, message, past_ids, past_key_values,
        video_end."""
        t_ingest = time.perf_counter()
        source = self._stream_source(state, max_pixels)
        if source is None:
            return
        stamps = self._due_timestamps(state)
        if stamps is None:
            return
        reader, height, width = source
        opening = state.get("last_timestamp", -1) < 0
        clip, stamps, frame_idxs = get_smart_resized_clip(reader, height, width, stamps, state["video_pts"],
                                                          video_pts_index_from=state["last_video_pts_index"] + 1,
                                                          device=self.ingest_device)
        if len(frame_idxs) == 0:
            return
        state["last_video_pts_index"], state["last_timestamp"] = frame_idxs[-1], stamps[-1]
        ingest_s = time.perf_counter() - t_ingest

        for frames, ts in self._chunks(clip, stamps, opening):
            t0 = time.perf_counter()
            span = (ts[0].item(), ts[-1].item() + self.frame_time_interval)
            # the query is attached on the first turn and whenever it changes (infer.py:140-146)
            if not message and not state.get("message"):
                message = default_query
            query = None
            if message and state.get("message") != message:
                query = state["message"] = message
            inputs = self._turn_inputs(frames, span, query, state.get("past_ids"))
            processors = None
            if streaming_eos_base_threshold is not None:
                processors = [ThresholdLogitsProcessor(self.streaming_eos_token_id, streaming_eos_base_threshold,
                                                       streaming_eos_threshold_step)]
            t1 = time.perf_counter()
            out = self.model.generate(**inputs, past_key_values=state.get("past_key_values"),
                                      return_dict_in_generate=True, do_sample=do_sample,
                                      repetition_penalty=repetition_penalty, logits_processor=processors,
                                      max_new_tokens=max_new_tokens, pad_token_id=self.model.config.eos_token_id)
            # cache length == len(past_ids): the last sampled token never entered the cache (infer.py:173-174)
            state["past_key_values"], state["past_ids"] = out.past_key_values, out.sequences[:, :-1]
            new_tokens = out.sequences[0, inputs.input_ids.size(1):]
            text = self.processor.decode(new_tokens, skip_special_tokens=True)
            t2 = time.perf_counter()
            self.timings.append(dict(frames=int(frames.shape[0]), new_tokens=int(new_tokens.numel()), ingest_s=ingest_s,
                                     preprocess_s=t1 - t0, generate_s=t2 - t1,
                                     kv_len=out.past_key_values.get_seq_length()))
            ingest_s = 0.0
            visible = state if not hf_spaces else {k: v for k, v in state.items()
                                                   if k not in ("past_ids", "past_key_values")}
            yield span, text, visible

    # ------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def video_qa(self, message: str, history: list, state: dict, do_sample: bool = False,
                 repetition_penalty: float = 1.05, hf_spaces: bool = False, max_new_tokens: int = 512, **kwargs):
        """Multi-turn QA over one video (REF/demo/infer.py:183-242): the whole video enters on the first turn
        (the reference resolves it through qwen_vl_utils.process_vision_info -> the 'decord+' reader that
        livecc_utils registers, i.e. `_read_video_decord_plus` + `_spatial_resize_video`), later turns only
        add text on top of the KV cache."""
        pending_video = state.get("video_path")
        turns = []
        if hf_spaces:  # stateless hosting mode: replay the history, video attached to its first message
            for old in history:
                parts = [{"type": "text", "text": old["content"]}]
                if pending_video:
                    parts.insert(0, {"type": "video", "video": pending_video})
                    pending_video = None
                turns.append({"role": old["role"], "content": parts})
        past_ids = state.get("past_ids")
        parts = [{"type": "text", "text": message}]
        if past_ids is None and pending_video:
            parts.insert(0, {"type": "video", "video": pending_video})
        turns.append({"role": "user", "content": parts})
        videos = [_spatial_resize_video(_read_video_decord_plus({"video": item["video"], "remote_loader": None})[0])
                  for turn in turns for item in turn["content"] if item["type"] == "video"]  # float frames (:106)
        text = self.processor.apply_chat_template(turns, tokenize=False, add_generation_prompt=True)
        if past_ids is not None:
            text = TURN_GLUE + text[self.system_prompt_offset:]
        inputs = self.processor(text=text, images=None, videos=videos or None, return_tensors="pt",
                                return_attention_mask=False)
        inputs.to(self.model.device)
        if past_ids is not None:
            inputs["input_ids"] = torch.cat([past_ids, inputs.input_ids], dim=1)
        out = self.model.generate(**inputs, past_key_values=state.get("past_key_values"), return_dict_in_generate=True,
                                  do_sample=do_sample, repetition_penalty=repetition_penalty,
                                  max_new_tokens=max_new_tokens, pad_token_id=self.model.config.eos_token_id)
        keep = not hf_spaces
        state["past_key_values"] = out.past_key_values if keep else None
        state["past_ids"] = out.sequences[:, :-1] if keep else None
        answer = self.processor.decode(out.sequences[0, inputs.input_ids.size(1):], skip_special_tokens=True)
        return answer, state

    @torch.inference_mode()
    def live_cc_once_for_evaluation(self, query: str, video: str, video_start: float = 0, video_end: float = None,
                                    remote_loader: callable = None, max_new_tokens: int = 32,
                                    repetition_penalty: float = 1.05):
        """Offline variant of live_cc (REF/demo/infer.py:245-310): the clip is read once and walked with the
        same 6 + 2 + 2 ... chunking; returns [[t0, t1, text], ...] in absolute video time. On purpose the
        processor is called with `return_attention_mask=False` like the demo path: the reference passes a
        new-tokens-only mask next to full-history ids here, which transformers 5.x mis-slices."""
        clip, _ = _read_video_decord_plus({"video": video, "video_start": video_start, "video_end": video_end,
                                           "remote_loader": remote_loader})
        clip = _spatial_resize_video(clip)
        pieces = [clip[: self.initial_fps_frames]]
        if len(clip) > self.initial_fps_frames:
            pieces.extend(clip[self.initial_fps_frames:].split(self.streaming_fps_frames))
        cache, past_ids, t_stop, results = None, None, 0.0, []
        for i, frames in enumerate(pieces):
            t_start, t_stop = (0.0, self.initial_time_interval) if i == 0 else (t_stop, t_stop + self.streaming_time_interval)
            inputs = self._turn_inputs(frames, (t_start, t_stop), query if cache is None else None, past_ids)
            out = self.model.generate(**inputs, past_key_values=cache, return_dict_in_generate=True,
                                      max_new_tokens=max_new_tokens, repetition_penalty=repetition_penalty,
                                      pad_token_id=self.model.config.eos_token_id)
            cache, past_ids = out.past_key_values, out.sequences[:, :-1]
            results.append([video_start + t_start, video_start + t_stop,
                            self.processor.decode(out.sequences[0, inputs.input_ids.size(1):], skip_special_tokens=True)])
        return results
