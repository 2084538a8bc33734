"""Load the REAL reference sources from /root/reference into this interpreter (CPU container only).

The reference package cannot be imported as a package here: ``my_affectgpt/__init__.py`` needs omegaconf, and the
modules below import decord / pytorchvideo / peft / transformers-4.x symbols that are not installed.
This loader executes the reference's own files unmodified, with *stub modules* standing in for the missing
third-party imports - none of the stubs contains arithmetic of the path, except the pytorchvideo clip sampler, which
is absent from the container and is restated from its published source (flagged wherever it is used).

Used only by ``tests/golden/make_reference_golden.py`` to generate fixtures; nothing under tests/ -m gpu, bench.py or
smoke() imports it (``/root/reference`` does not exist on the GPU box).
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference/AffectGPT"


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name: str, relpath: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, f"{REF}/{relpath}")
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------ Q-Former (Qformer.py)
class _PreTrainedModelShim(nn.Module):
    """What Qformer.py uses of transformers-4.49 ``PreTrainedModel`` / ``ModuleUtilsMixin`` (eval-mode forward only)."""
    config_class = None
    base_model_prefix = "bert"

    def __init__(self, config, *args, **kwargs):
        super().__init__()
        self.config = config

    def init_weights(self):
        self.apply(self._init_weights)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
        assert head_mask is None
        return [None] * num_hidden_layers

    def invert_attention_mask(self, encoder_attention_mask):
        # transformers 4.49 modeling_utils.ModuleUtilsMixin.invert_attention_mask
        if encoder_attention_mask.dim() == 3:
            ext = encoder_attention_mask[:, None, :, :]
        else:
            ext = encoder_attention_mask[:, None, None, :]
        ext = ext.to(dtype=self.dtype)
        return (1.0 - ext) * torch.finfo(self.dtype).min


def load_qformer():
    import transformers.modeling_utils as mu
    from transformers import pytorch_utils as pu
    saved = {k: getattr(mu, k, None) for k in ("PreTrainedModel", "apply_chunking_to_forward",
                                               "find_pruneable_heads_and_indices", "prune_linear_layer")}
    mu.PreTrainedModel = _PreTrainedModelShim
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("pruning unused"))
    try:
        return _load("my_affectgpt.models.Qformer", "my_affectgpt/models/Qformer.py")
    finally:
        for k, v in saved.items():
            if v is None:
                delattr(mu, k)
            else:
                setattr(mu, k, v)


# ------------------------------------------------------------------------------------------ AffectGPT (affectgpt.py)
def load_affectgpt():
    """Returns the reference ``AffectGPT`` class (methods only; __init__ needs the 7B LLM and is never called)."""
    qf = load_qformer()
    reg = _load("my_affectgpt.common.registry", "my_affectgpt/common/registry.py")

    class BaseModel(nn.Module):                       # stand-in for my_affectgpt/models/base_model.py (needs omegaconf)
        pass

    class Blip2Base(BaseModel):                       # stand-in for my_affectgpt/models/blip2.py (needs omegaconf/timm)
        def maybe_autocast(self, dtype=torch.float16):
            return contextlib.nullcontext()           # blip2.py:36-49 returns nullcontext() on CPU

    for pkg in ("my_affectgpt", "my_affectgpt.models", "my_affectgpt.common"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    sys.modules["my_affectgpt.models"].BaseModel = BaseModel
    _module("my_affectgpt.models.blip2", Blip2Base=Blip2Base, disabled_train=lambda self, mode=True: self)
    _module("my_affectgpt.models.tokenizer", load_tokenizer_from_LLM=None)
    _module("my_affectgpt.models.encoder")
    _module("config")
    for pkg in ("my_affectgpt", "my_affectgpt.models", "my_affectgpt.common"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    from transformers import BertConfig
    orig = BertConfig.from_pretrained
    BertConfig.from_pretrained = classmethod(lambda cls, *a, **k: cls())   # "models/bert-base-uncased" == BertConfig defaults
    qf.BertConfig.from_pretrained = BertConfig.from_pretrained
    mod = _load("my_affectgpt.models.affectgpt", "my_affectgpt/models/affectgpt.py")
    mod._restore_bertconfig = lambda: setattr(BertConfig, "from_pretrained", orig)
    return mod.AffectGPT


def bare_affectgpt(AffectGPT, **attrs):
    """An AffectGPT instance without running its __init__ (which loads the LLM): only nn.Module state + attrs."""
    obj = AffectGPT.__new__(AffectGPT)
    nn.Module.__init__(obj)
    for k, v in attrs.items():
        setattr(obj, k, v)
    return obj


# ------------------------------------------------------------------------------------------ processors / samplers
class FakeVideoReader:
    """decord.VideoReader stand-in: a video of ``vlen`` frames whose frame i is a 2x2x3 image filled with i."""
    vlen = 0

    def __init__(self, uri=None, height=-1, width=-1):
        self.n = FakeVideoReader.vlen

    def __len__(self):
        return self.n

    def get_batch(self, indices):
        idx = torch.tensor(list(indices), dtype=torch.float32)
        return idx.view(-1, 1, 1, 1).expand(-1, 2, 2, 3).clone()

    def get_avg_fps(self):
        return 25.0


def load_video_processor():
    regmod = _load("my_affectgpt.common.registry", "my_affectgpt/common/registry.py")
    regmod.registry.register("MAX_INT", sys.maxsize)
    bridge = types.SimpleNamespace(set_bridge=lambda *_: None)
    _module("decord", VideoReader=FakeVideoReader, bridge=bridge)
    _module("omegaconf", OmegaConf=object)
    for pkg in ("my_affectgpt", "my_affectgpt.processors", "my_affectgpt.common"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    _module("my_affectgpt.processors.transforms_video")
    _module("my_affectgpt.processors.functional_video")
    class BaseProcessor:                               # stand-in for processors/base_processor.py (needs omegaconf)
        pass

    _module("my_affectgpt.processors.base_processor", BaseProcessor=BaseProcessor)
    sys.modules["my_affectgpt.processors"].BaseProcessor = BaseProcessor
    _module("my_affectgpt.processors.randaugment", VideoRandomAugment=object)
    sys.modules["my_affectgpt.processors"].transforms_video = sys.modules["my_affectgpt.processors.transforms_video"]
    sys.modules["my_affectgpt.processors"].functional_video = sys.modules["my_affectgpt.processors.functional_video"]
    return _load("my_affectgpt.processors.video_processor", "my_affectgpt/processors/video_processor.py")


class FakeAudio:
    """torchaudio.load stand-in: returns the waveform registered for the path."""
    table = {}

    @staticmethod
    def load(path):
        return FakeAudio.table[path]


def load_imagebind_data():
    import os
    import torchaudio
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    # pytorchvideo==0.1.5 is NOT in the container: its clip sampler is the restatement in oracle/sampling.py
    from oracle.sampling import ConstantClipsPerVideoSampler
    _module("pytorchvideo")
    _module("pytorchvideo.transforms")
    _module("pytorchvideo.data")
    _module("pytorchvideo.data.clip_sampling", ConstantClipsPerVideoSampler=ConstantClipsPerVideoSampler)
    _module("pytorchvideo.data.encoded_video", EncodedVideo=object)
    for pkg in ("my_affectgpt", "my_affectgpt.models", "my_affectgpt.models.ImageBind", "my_affectgpt.models.ImageBind.models"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    _module("my_affectgpt.models.ImageBind.models.multimodal_preprocessors", SimpleTokenizer=object)
    mod = _load("my_affectgpt.models.ImageBind.data", "my_affectgpt/models/ImageBind/data.py")
    mod.torchaudio = types.SimpleNamespace(load=FakeAudio.load, functional=torchaudio.functional,
                                           compliance=getattr(torchaudio, "compliance", None))
    return mod


# ------------------------------------------------------------------------------------------ encoder wrappers (encoder.py)
def load_encoders(path_visual: dict, path_audio: dict):
    """The reference's registry encoder classes (``encoder.py``), constructed from checkpoints on disk.

    Shims for transformers 4.49 -> 5.5 API drift, both outside the arithmetic of the path:
    * ``AutoFeatureExtractor.from_pretrained`` no longer resolves CLIP: mapped to ``CLIPImageProcessor``;
    * ``CLIPModel.get_image_features`` returns an output object in 5.x: unwrapped to ``pooler_output`` (the 4.49 tensor).
    """
    import transformers
    for pkg in ("my_affectgpt", "my_affectgpt.models", "my_affectgpt.common", "my_affectgpt.models.ImageBind",
                "my_affectgpt.models.ImageBind.models"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    du = _module("my_affectgpt.common.dist_utils", download_cached_file=None)
    sys.modules["my_affectgpt.common"].dist_utils = du
    _module("my_affectgpt.common.utils", is_url=lambda *_: False)
    _module("my_affectgpt.common.logger", MetricLogger=object)

    class BaseModel(nn.Module):
        pass

    class Blip2Base(BaseModel):
        def maybe_autocast(self, dtype=torch.float16):
            return contextlib.nullcontext()

    _module("my_affectgpt.models.base_model", BaseModel=BaseModel)
    sys.modules["my_affectgpt.models"].BaseModel = BaseModel
    _module("my_affectgpt.models.blip2", Blip2Base=Blip2Base, disabled_train=lambda self, mode=True: self)
    _module("my_affectgpt.models.eva_vit", create_eva_vit_g=None)
    _module("my_affectgpt.models.Qformer", BertConfig=None, BertLMHeadModel=None)
    ib = _module("my_affectgpt.models.ImageBind.models.imagebind_model", ImageBindModel=object, ModalityType=object)
    sys.modules["my_affectgpt.models.ImageBind.models"].imagebind_model = ib
    _module("config", PATH_TO_VISUAL=path_visual, PATH_TO_AUDIO=path_audio)
    _load("my_affectgpt.common.registry", "my_affectgpt/common/registry.py")

    class _AutoFeatureExtractor:
        @staticmethod
        def from_pretrained(model_dir, *a, **k):
            return transformers.CLIPImageProcessor.from_pretrained(model_dir)

    orig_afe = transformers.AutoFeatureExtractor
    orig_gif = transformers.CLIPModel.get_image_features

    def get_image_features(self, *a, **k):
        out = orig_gif(self, *a, **k)
        return out if isinstance(out, torch.Tensor) else out.pooler_output

    transformers.AutoFeatureExtractor = _AutoFeatureExtractor
    transformers.CLIPModel.get_image_features = get_image_features
    try:
        mod = _load("my_affectgpt.models.encoder", "my_affectgpt/models/encoder.py")
    finally:
        transformers.AutoFeatureExtractor = orig_afe
    return mod
