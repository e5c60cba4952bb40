"""`prepare_multiturn_multimodal_inputs_for_generation` for the native engine.

Specification: REF/livecc-utils/src/livecc_utils/generation_patch.py:2-41 patches HF's
`prepare_inputs_for_generation` so that in a multi-turn stream (a) only the ids the cache has not seen are
forwarded, (b) position ids are left to the model (they derive from rope_deltas), and (c) on a turn after
the first one the pixel tensors are dropped unless the new ids still contain the video placeholder.
The engine's `generate()` implements (a)-(c) internally, so assigning this function to
`model.prepare_inputs_for_generation` (REF/demo/infer.py:50) has no further effect; it remains callable with
the reference's signature for code that invokes it directly.
"""
from __future__ import annotations

_PIXEL_KEYS = ("pixel_values", "pixel_values_videos")


def prepare_multiturn_multimodal_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None,
                                                       inputs_embeds=None, cache_position=None, position_ids=None,
                                                       use_cache=True, pixel_values=None, pixel_values_videos=None,
                                                       image_grid_thw=None, video_grid_thw=None, **kwargs):
    cached = 0 if past_key_values is None else past_key_values.get_seq_length()
    fresh_ids = input_ids[:, cached:]                       # (a) the cache decides what is new
    prepared = dict(kwargs)
    prepared.update(input_ids=fresh_ids, past_key_values=past_key_values, attention_mask=attention_mask,
                    inputs_embeds=inputs_embeds, use_cache=use_cache, image_grid_thw=image_grid_thw,
                    video_grid_thw=video_grid_thw, cache_position=None,
                    position_ids=None)                      # (b) positions come from rope_deltas
    carries_video = bool((fresh_ids == self.config.video_token_id).any())
    keep_pixels = cached == 0 or carries_video              # (c) streaming turns bring new frames
    prepared["pixel_values"] = pixel_values if keep_pixels else None
    prepared["pixel_values_videos"] = pixel_values_videos if keep_pixels else None
    return prepared
