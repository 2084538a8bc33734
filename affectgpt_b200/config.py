"""Paths and placeholder tokens, the counterpart of the reference's top-level config.py
(A/config.py:13-26, 121-127).  The zero-argument encoder constructors read their checkpoint directory
here, as the reference's do (A/my_affectgpt/models/encoder.py:183, 401)."""
import os

PATH_TO_VISUAL = {
    "CLIP_VIT_LARGE": os.environ.get("AGB_CLIP_VIT_LARGE", "/home/project/Dataset/Emotion/tools/transformer/visual/clip-vit-large-patch14"),
}
PATH_TO_AUDIO = {
    "HUBERT_LARGE": os.environ.get("AGB_HUBERT_LARGE", "/home/project/Dataset/Emotion/tools/transformer/chinese-hubert-large"),
}

DEFAULT_IMAGE_PATCH_TOKEN = "<ImageHere>"
DEFAULT_AUDIO_PATCH_TOKEN = "<AudioHere>"
DEFAULT_FRAME_PATCH_TOKEN = "<FrameHere>"
DEFAULT_FACE_PATCH_TOKEN = "<FaceHere>"
DEFAULT_MULTI_PATCH_TOKEN = "<MultiHere>"
IGNORE_INDEX = -100

LLM_DIM = 3584  # Qwen2.5-7B-Instruct hidden size (A/config.py:15)
