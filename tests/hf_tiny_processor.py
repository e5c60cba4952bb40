"""A REAL transformers `Qwen2VLProcessor` built offline (no hub files exist in this environment): byte-level BPE
tokenizer without merges (ids 0..255 = bytes) + the Qwen2-VL special tokens, the stock `Qwen2VLImageProcessor` /
`Qwen2VLVideoProcessor`, and a chat template with the Qwen2-VL layout, saved with `save_pretrained` so that
`AutoProcessor.from_pretrained(dir, use_fast=False)` -- the call REF/demo/infer.py:48 makes -- loads it. Used to run the
unmodified HF processor code under LiveCCDemoInfer."""
import copy

SPECIALS = ["<|im_end|>", "<|endoftext|>", "<|im_start|>", "<|vision_start|>", "<|vision_end|>", "<|vision_pad|>",
            "<|image_pad|>", "<|video_pad|>"]
TEMPLATE = (
    "{% for message in messages %}{% if loop.first and message['role'] != 'system' %}<|im_start|>system\n"
    "You are a helpful assistant.<|im_end|>\n{% endif %}<|im_start|>{{ message['role'] }}\n"
    "{% if message['content'] is string %}{{ message['content'] }}<|im_end|>\n{% else %}"
    "{% for content in message['content'] %}{% if content['type'] == 'image' %}<|vision_start|><|image_pad|><|vision_end|>"
    "{% elif content['type'] == 'video' %}<|vision_start|><|video_pad|><|vision_end|>"
    "{% elif 'text' in content %}{{ content['text'] }}{% endif %}{% endfor %}<|im_end|>\n{% endif %}{% endfor %}"
    "{% if add_generation_prompt %}<|im_start|>assistant\n{% endif %}")


def build_processor_dir(path: str) -> str:
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast, Qwen2VLImageProcessor, Qwen2VLProcessor, Qwen2VLVideoProcessor

    vocab = {ch: i for i, ch in enumerate(sorted(pre_tokenizers.ByteLevel.alphabet()))}
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|im_end|>", pad_token="<|endoftext|>",
                                   additional_special_tokens=SPECIALS)
    proc = Qwen2VLProcessor(image_processor=Qwen2VLImageProcessor(), tokenizer=fast,
                            video_processor=Qwen2VLVideoProcessor(min_pixels=3136, max_pixels=12845056),
                            chat_template=TEMPLATE)
    proc.save_pretrained(path)
    return path


def config_for(processor, base_cfg):
    """LiveCCConfig whose special ids are the tokenizer's (vocab 512 >= 264 ids)."""
    cfg = copy.deepcopy(base_cfg).with_vocab(512)
    t = processor.tokenizer
    ids = {s: t.convert_tokens_to_ids(s) for s in SPECIALS}
    cfg.eos_token_id = ids["<|im_end|>"]
    cfg.bos_token_id = cfg.pad_token_id = ids["<|endoftext|>"]
    cfg.im_start_token_id = ids["<|im_start|>"]
    cfg.vision_start_token_id, cfg.vision_end_token_id = ids["<|vision_start|>"], ids["<|vision_end|>"]
    cfg.image_token_id, cfg.video_token_id = ids["<|image_pad|>"], ids["<|video_pad|>"]
    cfg.newline_token_id = t("\n").input_ids[-1]
    return cfg
