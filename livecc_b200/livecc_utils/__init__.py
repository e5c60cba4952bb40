"""Drop-in for the `livecc_utils` package (REF/livecc-utils/src/livecc_utils/__init__.py:1-2): same
public names, so `from livecc_utils import ...` in REF/demo/infer.py:5 can be pointed here."""
from .generation_patch import prepare_multiturn_multimodal_inputs_for_generation
from .video_process_patch import (_read_video_decord_plus, _spatial_resize_video, get_smart_resized_clip,
                                  get_smart_resized_video_reader)

__all__ = ["prepare_multiturn_multimodal_inputs_for_generation", "_read_video_decord_plus", "_spatial_resize_video",
           "get_smart_resized_video_reader", "get_smart_resized_clip"]
