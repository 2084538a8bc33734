"""Drop-in encoder wrappers with the reference's names, registry keys and call signatures.

Reference: A/my_affectgpt/models/encoder.py -- ``CLIP_VIT_LARGE`` (:177-208) and ``HUBERT_LARGE``
(:395-432): zero-argument constructors (weights from ``config.PATH_TO_VISUAL / PATH_TO_AUDIO``), ``.hidden_size``,
``forward(image, raw_image) -> [b,t,768]`` and ``forward(audio, raw_audio) -> [b,t,1024]``; the normalised
``image`` / mel ``audio`` arguments are accepted and ignored exactly as the reference ignores them (SURVEY F1,
F3).  Outputs are fp32 on ``raw_*.device`` so that the callers' ``.cpu().numpy()`` keeps working.

What differs from the reference: the whole forward (uint8 truncation, CLIP normalisation, ViT-L/14, HuBERT
conv stack + transformer, last-4 mean + time mean) runs in libagb.so's sm_100a kernels on the GPU -- there is no
CPU round trip through PIL / numpy and no eager-PyTorch fallback.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import config
from .engine import Engine, get_engine
from .registry import registry


def _read_checkpoint(model_dir: str) -> Dict[str, torch.Tensor]:
    st = os.path.join(model_dir, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    pt = os.path.join(model_dir, "pytorch_model.bin")
    if os.path.exists(pt):
        return torch.load(pt, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {model_dir}")


def _read_json(path: str) -> dict:
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class _EncoderBase(nn.Module):
    family = ""

    def __init__(self):
        super().__init__()
        self._engine: Optional[Engine] = None
        self._state: Optional[Dict[str, torch.Tensor]] = None   # checkpoint tensors until uploaded
        self._engine_cfg: dict = {}
        self._device_hint: Optional[torch.device] = None

    # nn.Module plumbing the reference's callers use: .to(device), .eval(), named_parameters() (frozen / none).
    # A dtype request cannot be honoured - the kernels compute with bf16 operands / fp32 accumulation and return fp32 at
    # this boundary - so anything but float32 raises instead of being silently ignored.
    _PRECISION_MSG = ("{cls} computes with bf16 operands and fp32 accumulation inside libagb and returns float32; "
                      "it cannot be cast to {dtype} (cast the returned features instead)")

    def _reject_dtype(self, dtype) -> None:
        if dtype is not None and dtype != torch.float32:
            raise TypeError(self._PRECISION_MSG.format(cls=type(self).__name__, dtype=dtype))

    def to(self, *args, **kwargs):
        device = kwargs.get("device")
        dtype = kwargs.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif isinstance(a, (str, torch.device, int)):
                device = a
            elif isinstance(a, torch.Tensor):                   # .to(other): that tensor's device and dtype
                device, dtype = a.device, a.dtype
        self._reject_dtype(dtype)
        if device is not None:
            self._device_hint = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        return self

    def cuda(self, device=None):
        self._device_hint = torch.device("cuda", device if isinstance(device, int) else 0)
        return self

    def float(self):
        return self

    def half(self):
        self._reject_dtype(torch.float16)

    def bfloat16(self):
        self._reject_dtype(torch.bfloat16)

    def double(self):
        self._reject_dtype(torch.float64)

    def load_hf_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Provide the HuggingFace checkpoint tensors (keys of SURVEY appendix B); uploaded on first use."""
        self._state = {k: v.detach() for k, v in state_dict.items()}
        if self._engine is not None:
            self._upload()

    def attach_engine(self, engine: Engine) -> None:
        self._engine = engine
        if self._state is not None:
            self._upload()

    def _upload(self) -> None:
        self._engine.load_state_dict(self.family, self._state, strict=True)
        self._state = None

    def _ensure_engine(self, device: torch.device) -> Engine:
        if self._engine is None:
            if device.type != "cuda":
                raise RuntimeError(
                    f"{type(self).__name__}: inputs are on {device}; affectgpt_b200 runs on CUDA (sm_100a) only and has "
                    "no CPU fallback")
            self._engine = get_engine(device, **self._engine_cfg)
            if self._state is not None:
                self._upload()
        missing = self._engine.missing(self.family)
        if missing:
            raise RuntimeError(f"{type(self).__name__}: weights not loaded ({len(missing)} tensors missing, e.g. "
                               f"{missing[0]}); call load_hf_state_dict() or point config.PATH_TO_* at a checkpoint")
        return self._engine


@registry.register_visual_encoder("CLIP_VIT_LARGE")
class CLIP_VIT_LARGE(_EncoderBase):
    family = "clip"

    def __init__(self, model_dir: Optional[str] = None, num_layers: int = 24, engine_cfg: Optional[dict] = None):
        super().__init__()
        self.hidden_size = 768                      # projection_dim (encoder.py:192)
        self._engine_cfg = dict(engine_cfg or {})
        self._engine_cfg.setdefault("vit_layers", num_layers)
        model_dir = model_dir if model_dir is not None else config.PATH_TO_VISUAL.get("CLIP_VIT_LARGE")
        if model_dir and os.path.isdir(model_dir):
            cfg = _read_json(os.path.join(model_dir, "config.json")).get("vision_config", {})
            if cfg.get("num_hidden_layers"):
                self._engine_cfg["vit_layers"] = int(cfg["num_hidden_layers"])
            self._state = _read_checkpoint(model_dir)

    # image encoding: [b c t h w] => [b t h]   (encoder.py:196-208)
    @torch.no_grad()
    def forward(self, image, raw_image):
        del image  # ignored by the reference as well
        if raw_image.dim() != 5:
            raise ValueError(f"raw_image must be [b,3,t,224,224], got {tuple(raw_image.shape)}")
        b, c, t, h, w = raw_image.shape
        if (c, h, w) != (3, 224, 224):
            raise ValueError("CLIP_VIT_LARGE expects 224x224 RGB frames (load_video/load_face resize to 224)")
        eng = self._ensure_engine(raw_image.device)
        frames = raw_image.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)      # '(b t) c h w'
        emb = eng.clip_forward(frames)                                          # [(b t), 768] fp32
        return emb.reshape(b, t, self.hidden_size)


@registry.register_acoustic_encoder("HUBERT_LARGE")
class HUBERT_LARGE(_EncoderBase):
    family = "hubert"

    def __init__(self, model_dir: Optional[str] = None, num_layers: int = 24, feat_extract_norm: str = "layer",
                 do_normalize: bool = True, norm_scope: str = "clip", engine_cfg: Optional[dict] = None,
                 do_stable_layer_norm: bool = True, feat_proj_layer_norm: bool = True):
        """``norm_scope``: "clip" -- zero-mean/unit-variance statistics per clip (its t segments), which is what
        the reference computes at batch size 1, its only call pattern; "batch" -- jointly over the whole call,
        which reproduces the reference bit-for-bit in intent when it is fed b > 1."""
        super().__init__()
        if norm_scope not in ("clip", "batch"):
            raise ValueError("norm_scope must be 'clip' or 'batch'")
        self.hidden_size = 1024
        self.norm_scope = norm_scope
        self._engine_cfg = dict(engine_cfg or {})
        self._engine_cfg.setdefault("hubert_layers", num_layers)
        self._engine_cfg.setdefault("feat_extract_norm", feat_extract_norm)
        self._engine_cfg.setdefault("do_normalize", do_normalize)
        # HubertConfig switches of the 1024-wide architecture (encoder.py:403 loads whatever config.json says):
        # do_stable_layer_norm (pre-LN blocks, hubert-large family) vs post-LN blocks; feat_proj_layer_norm; conv_bias
        self._engine_cfg.setdefault("do_stable_layer_norm", do_stable_layer_norm)
        self._engine_cfg.setdefault("feat_proj_layer_norm", feat_proj_layer_norm)
        model_dir = model_dir if model_dir is not None else config.PATH_TO_AUDIO.get("HUBERT_LARGE")
        if model_dir and os.path.isdir(model_dir):
            cfg = _read_json(os.path.join(model_dir, "config.json"))
            pre = _read_json(os.path.join(model_dir, "preprocessor_config.json"))
            if cfg.get("num_hidden_layers"):
                self._engine_cfg["hubert_layers"] = int(cfg["num_hidden_layers"])
            if cfg.get("feat_extract_norm"):
                self._engine_cfg["feat_extract_norm"] = cfg["feat_extract_norm"]
            if cfg:
                if (cfg.get("hidden_size", 1024), cfg.get("num_attention_heads", 16), cfg.get("intermediate_size", 4096)) != (1024, 16, 4096):
                    raise NotImplementedError("HUBERT_LARGE kernels are built for hidden 1024 / 16 heads / FFN 4096 "
                                              f"(got {cfg.get('hidden_size')}/{cfg.get('num_attention_heads')}/{cfg.get('intermediate_size')})")
                self._engine_cfg["do_stable_layer_norm"] = bool(cfg.get("do_stable_layer_norm", False))
                self._engine_cfg["feat_proj_layer_norm"] = bool(cfg.get("feat_proj_layer_norm", True))
            if "do_normalize" in pre:
                self._engine_cfg["do_normalize"] = bool(pre["do_normalize"])
            self.load_hf_state_dict(_read_checkpoint(model_dir))

    def load_hf_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """conv_bias = false checkpoints carry no conv biases (HubertConfig.conv_bias): the kernels add a bias vector in
        every conv epilogue, so a missing one is filled with zeros - same result."""
        sd = dict(state_dict)
        for i in range(7):
            w, b = f"feature_extractor.conv_layers.{i}.conv.weight", f"feature_extractor.conv_layers.{i}.conv.bias"
            if w in sd and b not in sd:
                sd[b] = torch.zeros(sd[w].shape[0], dtype=torch.float32, device=sd[w].device)
        super().load_hf_state_dict(sd)

    # audio: [b, t, 1, 128, 204] mel (ignored); raw_audio: [b, t, 1, 32000]   (encoder.py:416-432)
    @torch.no_grad()
    def forward(self, audio, raw_audio):
        del audio  # ignored by the reference as well
        if raw_audio.dim() != 4 or raw_audio.shape[2] != 1 or raw_audio.shape[3] != 32000:
            raise ValueError(f"raw_audio must be [b,t,1,32000], got {tuple(raw_audio.shape)}")
        b, t = raw_audio.shape[:2]
        eng = self._ensure_engine(raw_audio.device)
        wav = raw_audio[:, :, 0, :].reshape(b * t, 32000)
        emb = eng.hubert_forward(wav, seg_per_norm_group=t if self.norm_scope == "clip" else 0)
        return emb.reshape(b, t, self.hidden_size)
