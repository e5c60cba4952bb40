"""REF/livecc-utils/src/livecc_utils/generation_patch.py:2-41, restated for the native engine.

The reference patches HF's `prepare_inputs_for_generation` so that, on a non-first prefill of a
multi-turn stream, `pixel_values_videos` is still forwarded iff the *new* ids contain the video token,
and position ids are left to the model. The engine implements exactly these semantics inside
`generate()` (it slices the new ids with the cache length, runs the ViT whenever pixel values are given
with a prefill, and derives positions from the per-stream rope_delta), so assigning this function to
`model.prepare_inputs_for_generation` (REF/demo/infer.py:50) is accepted and has no further effect.
The function itself stays callable and returns the same dictionary shape for callers that use it."""
from __future__ import annotations


def prepare_multiturn_multimodal_inputs_for_generation(
    self,
    input_ids,
    past_key_values=None,
    attention_mask=None,
    inputs_embeds=None,
    cache_position=None,
    position_ids=None,
    use_cache=True,
    pixel_values=None,
    pixel_values_videos=None,
    image_grid_thw=None,
    video_grid_thw=None,
    **kwargs,
):
    past = past_key_values.get_seq_length() if past_key_values is not None else 0
    new_ids = input_ids[:, past:]
    model_inputs = dict(
        input_ids=new_ids, past_key_values=past_key_values, attention_mask=attention_mask,
        inputs_embeds=inputs_embeds, use_cache=use_cache, pixel_values=pixel_values,
        pixel_values_videos=pixel_values_videos, image_grid_thw=image_grid_thw, video_grid_thw=video_grid_thw,
        cache_position=None, **kwargs,
    )
    # Qwen2-VL position ids are prepared with rope_deltas in forward (generation_patch.py:34-35)
    model_inputs["position_ids"] = None
    # streaming: keep the pixels on a later turn only if the new ids carry video tokens (:37-39)
    if past != 0 and bool((new_ids != self.config.video_token_id).all()):
        model_inputs["pixel_values"] = None
        model_inputs["pixel_values_videos"] = None
    return model_inputs
