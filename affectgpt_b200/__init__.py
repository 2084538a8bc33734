"""affectgpt_b200 -- B200-native encode-and-fuse path of AffectGPT (CLIP ViT-L/14 + HuBERT-large + pre-fusion +
projectors) behind the reference's encoder-wrapper and fusion call signatures.  See DESIGN.md."""
from .registry import registry  # noqa: F401
from . import config  # noqa: F401
from .encoders import CLIP_VIT_LARGE, HUBERT_LARGE  # noqa: F401
from .fusion import AffectGPTFusion, GraphedClipBatch, build_path  # noqa: F401
from .engine import Engine, get_engine, launch_count  # noqa: F401
from . import sampling, splice, staging  # noqa: F401

__all__ = ["registry", "config", "CLIP_VIT_LARGE", "HUBERT_LARGE", "AffectGPTFusion", "GraphedClipBatch", "build_path", "Engine",
           "get_engine", "launch_count", "sampling", "splice", "staging"]
