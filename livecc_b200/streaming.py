"""`LiveCCDemoInfer`: the streaming orchestrator of REF/demo/infer.py:25-180 on top of the native engine.

Same class constants, same `state` dictionary contract, same generator protocol
(`for (start, stop), response, state in infer.live_cc(message=..., state=state, ...)`), so
REF/demo/cli.py:13-24 runs unchanged against it. Differences, all outside the arithmetic:
  * the model is `LiveCCB200ForConditionalGeneration` (no liger / flash-attn / HF modeling);
  * the processor may be the offline `StubProcessor` (no tokenizer files exist in this environment);
  * `livecc_utils` is the in-tree re-implementation (decord / qwen_vl_utils are not installed);
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


class LiveCCDemoInfer:
    VIDEO_PLAY_END = object()
    VIDEO_PLAY_CONTINUE = object()
    fps = 2
    initial_fps_frames = 6
    streaming_fps_frames = 2
    initial_time_interval = initial_fps_frames / fps
    streaming_time_interval = streaming_fps_frames / fps
    frame_time_interval = 1 / fps

    def __init__(self, model_path: str = None, device: str = None, model=None, processor=None):
        """REF/demo/infer.py:35-59. Pass `model=`/`processor=` to reuse already-built objects
        (synthetic checkpoint); otherwise `model_path` must be a local HF checkpoint directory."""
        if model is None:
            if device is None:
                device = "cuda" if torch.cuda.is_available() else "cpu"
            model = LiveCCB200ForConditionalGeneration.from_pretrained(model_path, torch_dtype="auto", device_map=device)
        self.model = model
        if processor is None:
            try:
                from transformers import AutoProcessor

                processor = AutoProcessor.from_pretrained(model_path, use_fast=False)
            except Exception:
                processor = StubProcessor(model.config)
        self.processor = processor
        self.streaming_eos_token_id = self.processor.tokenizer(" ...").input_ids[-1]
        self.model.prepare_inputs_for_generation = functools.partial(
            prepare_multiturn_multimodal_inputs_for_generation, self.model)
        message = {"role": "user", "content": [{"type": "text", "text": "livecc"}]}
        texts = self.processor.apply_chat_template([message], tokenize=False)
        self.system_prompt_offset = texts.index("<|im_start|>user")
        self._cached_video_readers_with_hw = {}
        self.timings = []  # one dict per chunk: frames, new_tokens, ingest_s, generate_s

    @torch.inference_mode()
    def live_cc(
        self,
        message: str,
        state: dict,
        max_pixels: int = 384 * 28 * 28,
        default_query: str = "Please describe the video.",
        do_sample: bool = True,
        repetition_penalty: float = 1.05,
        streaming_eos_base_threshold: float = None,
        streaming_eos_threshold_step: float = None,
        hf_spaces: bool = False,
        max_new_tokens: int = 16,
        **kwargs,
    ):
        """REF/demo/infer.py:62-180 (state keys: video_path, video_timestamp, last_timestamp,
        last_video_pts_index, video_pts, message, past_ids, past_key_values, video_end)."""
        # 1. preparation: video_reader, and last processing info
        t_ingest0 = time.perf_counter()
        video_timestamp, last_timestamp = state.get("video_timestamp", 0), state.get("last_timestamp", -1 / self.fps)
        video_path = state.get("video_path", None)
        if not video_path:
            return
        if video_path not in self._cached_video_readers_with_hw:
            self._cached_video_readers_with_hw[video_path] = get_smart_resized_video_reader(video_path, max_pixels)
            video_reader = self._cached_video_readers_with_hw[video_path][0]
            video_reader.get_frame_timestamp(0)
            state["video_pts"] = torch.from_numpy(video_reader._frame_pts[:, 1])
            state["last_video_pts_index"] = -1
        video_pts = state.get("video_pts", None)
        if video_pts is None:
            return
        video_timestamp = min(video_timestamp, video_pts[-1])
        if last_timestamp + self.frame_time_interval > video_pts[-1]:
            state["video_end"] = True
            return
        video_reader, resized_height, resized_width = self._cached_video_readers_with_hw[video_path]
        last_video_pts_index = state["last_video_pts_index"]

        # 2. which frames will be processed
        initialized = last_timestamp >= 0
        if not initialized:
            video_timestamp = max(video_timestamp, self.initial_time_interval)
        if video_timestamp <= last_timestamp + self.frame_time_interval:
            return
        timestamps = torch.arange(last_timestamp + self.frame_time_interval, video_timestamp, self.frame_time_interval)

        # 3. fetch frames in required timestamps
        clip, clip_timestamps, clip_idxs = get_smart_resized_clip(
            video_reader, resized_height, resized_width, timestamps, video_pts,
            video_pts_index_from=last_video_pts_index + 1)
        if len(clip_idxs) == 0:
            return
        state["last_video_pts_index"] = clip_idxs[-1]
        state["last_timestamp"] = clip_timestamps[-1]

        # 4. organize to interleave frames
        interleave_clips, interleave_timestamps = [], []
        if not initialized:
            interleave_clips.append(clip[: self.initial_fps_frames])
            interleave_timestamps.append(clip_timestamps[: self.initial_fps_frames])
            clip = clip[self.initial_fps_frames:]
            clip_timestamps = clip_timestamps[self.initial_fps_frames:]
        if len(clip) > 0:
            interleave_clips.extend(list(clip.split(self.streaming_fps_frames)))
            interleave_timestamps.extend(list(clip_timestamps.split(self.streaming_fps_frames)))
        ingest_s = time.perf_counter() - t_ingest0

        # 5. make conversation and send to model
        for clip, timestamps in zip(interleave_clips, interleave_timestamps):
            t0 = time.perf_counter()
            start_timestamp, stop_timestamp = timestamps[0].item(), timestamps[-1].item() + self.frame_time_interval
            conversation = [{
                "role": "user",
                "content": [
                    {"type": "text", "text": f"Time={start_timestamp:.1f}-{stop_timestamp:.1f}s"},
                    {"type": "video", "video": clip},
                ],
            }]
            if not message and not state.get("message", None):
                message = default_query
            if message and state.get("message", None) != message:
                conversation[0]["content"].append({"type": "text", "text": message})
                state["message"] = message
            texts = self.processor.apply_chat_template(conversation, tokenize=False, add_generation_prompt=True)
            past_ids = state.get("past_ids", None)
            if past_ids is not None:
                texts = "<|im_end|>\n" + texts[self.system_prompt_offset:]
            inputs = self.processor(text=texts, images=None, videos=[clip], return_tensors="pt",
                                    return_attention_mask=False)
            inputs.to(self.model.device)
            if past_ids is not None:
                inputs["input_ids"] = torch.cat([past_ids, inputs.input_ids], dim=1)
            if streaming_eos_base_threshold is not None:
                logits_processor = [ThresholdLogitsProcessor(self.streaming_eos_token_id, streaming_eos_base_threshold,
                                                             streaming_eos_threshold_step)]
            else:
                logits_processor = None
            t1 = time.perf_counter()
            outputs = self.model.generate(
                **inputs, past_key_values=state.get("past_key_values", None),
                return_dict_in_generate=True, do_sample=do_sample,
                repetition_penalty=repetition_penalty,
                logits_processor=logits_processor,
                max_new_tokens=max_new_tokens,
                pad_token_id=self.model.config.eos_token_id,
            )
            state["past_key_values"] = outputs.past_key_values
            state["past_ids"] = outputs.sequences[:, :-1]
            new_tokens = outputs.sequences[0, inputs.input_ids.size(1):]
            response = self.processor.decode(new_tokens, skip_special_tokens=True)
            t2 = time.perf_counter()
            self.timings.append(dict(frames=int(clip.shape[0]), new_tokens=int(new_tokens.numel()),
                                     ingest_s=ingest_s, preprocess_s=t1 - t0, generate_s=t2 - t1,
                                     kv_len=outputs.past_key_values.get_seq_length()))
            ingest_s = 0.0
            if hf_spaces:
                light_state = {k: v for k, v in state.items() if k not in ["past_ids", "past_key_values"]}
                yield (start_timestamp, stop_timestamp), response, light_state
            else:
                yield (start_timestamp, stop_timestamp), response, state

    @torch.inference_mode()
    def video_qa(
        self,
        message: str,
        history: list,
        state: dict,
        do_sample: bool = False,
        repetition_penalty: float = 1.05,
        hf_spaces: bool = False,
        max_new_tokens: int = 512,
        **kwargs,
    ):
        """REF/demo/infer.py:183-242: multi-turn QA over one video. The whole video enters on the first turn
        (the reference goes through qwen_vl_utils.process_vision_info -> the 'decord+' reader registered by
        livecc_utils, i.e. `_read_video_decord_plus` + `_spatial_resize_video`); later turns reuse the KV cache."""
        video_path = state.get("video_path", None)
        conversation = []
        if hf_spaces:
            for past_message in history:
                content = [{"type": "text", "text": past_message["content"]}]
                if video_path:  # only use once
                    content.insert(0, {"type": "video", "video": video_path})
                    video_path = None
                conversation.append({"role": past_message["role"], "content": content})
        past_ids = state.get("past_ids", None)
        content = [{"type": "text", "text": message}]
        if past_ids is None and video_path:  # only use once
            content.insert(0, {"type": "video", "video": video_path})
        conversation.append({"role": "user", "content": content})
        video_inputs = []
        for msg in conversation:  # process_vision_info equivalent for the video-only path
            for item in msg["content"]:
                if item["type"] == "video":
                    clip, _fps = _read_video_decord_plus({"video": item["video"], "remote_loader": None})
                    video_inputs.append(_spatial_resize_video(clip))  # float frames, like the reference (video_process_patch.py:106)
        texts = self.processor.apply_chat_template(conversation, tokenize=False, add_generation_prompt=True)
        if past_ids is not None:
            texts = "<|im_end|>\n" + texts[self.system_prompt_offset:]
        inputs = self.processor(text=texts, images=None, videos=video_inputs or None, return_tensors="pt",
                                return_attention_mask=False)
        inputs.to(self.model.device)
        if past_ids is not None:
            inputs["input_ids"] = torch.cat([past_ids, inputs.input_ids], dim=1)
        outputs = self.model.generate(
            **inputs, past_key_values=state.get("past_key_values", None),
            return_dict_in_generate=True, do_sample=do_sample,
            repetition_penalty=repetition_penalty,
            max_new_tokens=max_new_tokens,
            pad_token_id=self.model.config.eos_token_id,
        )
        state["past_key_values"] = outputs.past_key_values if not hf_spaces else None
        state["past_ids"] = outputs.sequences[:, :-1] if not hf_spaces else None
        response = self.processor.decode(outputs.sequences[0, inputs.input_ids.size(1):], skip_special_tokens=True)
        return response, state

    @torch.inference_mode()
    def live_cc_once_for_evaluation(
        self,
        query: str,
        video: str,
        video_start: float = 0,
        video_end: float = None,
        remote_loader: callable = None,
        max_new_tokens: int = 32,
        repetition_penalty: float = 1.05,
    ):
        """REF/demo/infer.py:245-310: offline variant of live_cc (clip read once, same 6+2+2... chunking).
        Difference from the reference, on purpose: `return_attention_mask=False` like the demo path — the
        reference passes a new-tokens-only mask next to full-history ids here, which transformers 5.x would
        mis-slice."""
        clip, _ = _read_video_decord_plus({"video": video, "video_start": video_start, "video_end": video_end,
                                           "remote_loader": remote_loader})
        clip = _spatial_resize_video(clip)
        interleave_clips = [clip[: self.initial_fps_frames]]
        clip = clip[self.initial_fps_frames:]
        if len(clip) > 0:
            interleave_clips.extend(list(clip.split(self.streaming_fps_frames)))
        past_key_values = None
        past_ids = None
        responses = []
        start_timestamp = stop_timestamp = 0
        for i, clip in enumerate(interleave_clips):
            if i == 0:
                start_timestamp, stop_timestamp = 0, self.initial_time_interval
            else:
                start_timestamp, stop_timestamp = stop_timestamp, stop_timestamp + self.streaming_time_interval
            message = {
                "role": "user",
                "content": [
                    {"type": "text", "text": f"Time={start_timestamp:.1f}-{stop_timestamp:.1f}s"},
                    {"type": "video", "video": clip},
                ],
            }
            if not past_key_values:
                message["content"].append({"type": "text", "text": query})
            texts = self.processor.apply_chat_template([message], tokenize=False, add_generation_prompt=True)
            if past_key_values:
                texts = "<|im_end|>\n" + texts[self.system_prompt_offset:]
            inputs = self.processor(text=texts, images=None, videos=[clip], return_tensors="pt",
                                    return_attention_mask=False)
            inputs.to(self.model.device)
            if past_key_values:
                inputs["input_ids"] = torch.cat([past_ids, inputs.input_ids], dim=1)
            outputs = self.model.generate(
                **inputs, past_key_values=past_key_values,
                return_dict_in_generate=True,
                max_new_tokens=max_new_tokens, repetition_penalty=repetition_penalty,
                pad_token_id=self.model.config.eos_token_id,
            )
            past_key_values = outputs.past_key_values
            past_ids = outputs.sequences[:, :-1]
            responses.append([
                video_start + start_timestamp,
                video_start + stop_timestamp,
                self.processor.decode(outputs.sequences[0, inputs.input_ids.size(1):], skip_special_tokens=True),
            ])
        return responses
