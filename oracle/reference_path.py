"""fp32 CPU restatement of the reference's encode-and-fuse path (TEST INFRASTRUCTURE ONLY).

Reference = /root/reference/AffectGPT (``A/``).  The heavy arithmetic of the
reference lives in HuggingFace ``transformers`` (un-vendored dependency), so the
oracle calls those very modules (random-init, eval, fp32, CPU) and restates only
the reference's own thin code around them:

* ``clip_vit_large_forward``  <- A/my_affectgpt/models/encoder.py:197-208 (+ :34-39)
* ``hubert_large_forward``    <- A/my_affectgpt/models/encoder.py:418-432
  (+ HF Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm quirk)
* ``encode_video_attention``  <- A/my_affectgpt/models/affectgpt.py:505-535, 585-599
* ``encode_audio_attention``  <- A/my_affectgpt/models/affectgpt.py:690-716, 751-764
* ``encode_multi_attention``  <- A/my_affectgpt/models/affectgpt.py:883-908
* mean-fusion variants        <- A/my_affectgpt/models/affectgpt.py:601-608, 766-773

Parity status: pinned against the reference's own source files run in the build
container (tests/golden/ref_loader.py -> tests/golden/reference_*.{json,npz});
the reference itself ships no golden vectors.  See oracle/__init__.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

LLM_DIM = 3584  # Qwen2.5-7B hidden size (A/config.py:15)


# --------------------------------------------------------------------------- #
# model construction (random init; there is no network for checkpoints)
# --------------------------------------------------------------------------- #
def build_clip_vision(num_layers: int = 24, seed: int = 1234):
    """CLIP ViT-L/14 vision tower + visual_projection (what get_image_features uses).

    HF: models/clip/modeling_clip.py (CLIPVisionTransformer, visual_projection).
    State-dict keys equal the ``vision_model.*`` / ``visual_projection.weight``
    keys of ``CLIPModel``.
    """
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    torch.manual_seed(seed)
    cfg = CLIPVisionConfig(
        hidden_size=1024, intermediate_size=4096, num_hidden_layers=num_layers,
        num_attention_heads=16, patch_size=14, image_size=224, hidden_act="quick_gelu",
        projection_dim=768, layer_norm_eps=1e-5,
    )
    model = CLIPVisionModelWithProjection(cfg).eval()
    randomize_affine(model, seed + 1)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def build_hubert(num_layers: int = 24, feat_extract_norm: str = "layer", seed: int = 1234,
                 do_stable_layer_norm: bool = True, conv_bias: bool = True, feat_proj_layer_norm: bool = True):
    """HuBERT at the large width, random init.  HF: models/hubert/modeling_hubert.py.  Defaults = the hubert-large family
    (stable / pre-LN blocks); do_stable_layer_norm=False selects HubertEncoder / HubertEncoderLayer (post-LN)."""
    from transformers import HubertConfig, HubertModel

    torch.manual_seed(seed)
    cfg = HubertConfig(
        hidden_size=1024, num_hidden_layers=num_layers, num_attention_heads=16,
        intermediate_size=4096, feat_extract_norm=feat_extract_norm, do_stable_layer_norm=do_stable_layer_norm,
        conv_bias=conv_bias, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
        conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2), conv_kernel=(10, 3, 3, 3, 3, 2, 2),
        feat_proj_layer_norm=feat_proj_layer_norm, layer_norm_eps=1e-5, hidden_act="gelu",
        feat_extract_activation="gelu",
    )
    model = HubertModel(cfg).eval()
    randomize_affine(model, seed + 1)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def randomize_affine(model: nn.Module, seed: int) -> None:
    """HF's default init zeroes every bias and sets every norm scale to 1, which would hide
    bias / affine bugs in the path under test.  Perturb them deterministically."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            is_norm = ("norm" in name) or ("layrnorm" in name)
            if name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * (0.1 if is_norm else 0.05))
            elif is_norm and name.endswith(".weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("parametrizations.weight.original0"):
                # weight-norm gain of the positional conv: make it non-trivial per tap
                p.copy_(p * (1.0 + 0.2 * torch.randn(p.shape, generator=g)))


def build_fusion_state(seed: int = 1234, llm_dim: int = LLM_DIM) -> Dict[str, torch.Tensor]:
    """Fusion heads + projectors with nn.Linear default init, in the construction order of
    A/my_affectgpt/models/affectgpt.py (:190 video_attention_mlp, :195 affectgpt_proj,
    :251 audio_attention_mlp, :255 audio_llama_proj, :312-313 multi_*_embs,
    :344-345 attention_mlp / fc_att, :349 multi_llama_proj)."""
    torch.manual_seed(seed)
    mods = {
        "video_attention_mlp": nn.Linear(768, 1),
        "affectgpt_proj": nn.Linear(768, llm_dim),
        "audio_attention_mlp": nn.Linear(1024, 1),
        "audio_llama_proj": nn.Linear(1024, llm_dim),
        "multi_video_embs": nn.Linear(768, 1024),
        "multi_audio_embs": nn.Linear(1024, 1024),
        "attention_mlp": nn.Linear(2048, 1024),
        "fc_att": nn.Linear(1024, 2),
        "multi_llama_proj": nn.Linear(1024, llm_dim),
    }
    state = {}
    for k, m in mods.items():
        state[k + ".weight"] = m.weight.detach().clone()
        state[k + ".bias"] = m.bias.detach().clone()
    return state


# --------------------------------------------------------------------------- #
# encoder wrappers
# --------------------------------------------------------------------------- #
def clip_preprocess(raw_image: torch.Tensor) -> torch.Tensor:
    """encoder.py:200-204 + :34-39.  ``raw_image`` [b,3,t,224,224] float holding 0..255.
    PIL round trip == uint8 truncation; HF CLIP processor at 224x224 == (x/255-mean)/std
    (checked against CLIPImageProcessor in tests/test_oracle.py)."""
    b, c, t, h, w = raw_image.shape
    assert (c, h, w) == (3, 224, 224), "reference path is fed 224x224 frames (load_video height/width)"
    frames = raw_image.permute(0, 2, 3, 4, 1).reshape(b * t, h, w, c)            # '(b t) h w c'
    u8 = torch.from_numpy(frames.cpu().numpy().astype(np.uint8))                  # Image.fromarray(...astype(uint8))
    x = u8.to(torch.float32) * (1.0 / 255.0)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32)
    std = torch.tensor(CLIP_STD, dtype=torch.float32)
    x = (x - mean) / std
    return x.permute(0, 3, 1, 2).contiguous()                                      # [(b t),3,224,224]


@torch.no_grad()
def clip_vit_large_forward(model, raw_image: torch.Tensor) -> torch.Tensor:
    """CLIP_VIT_LARGE.forward (encoder.py:197-208): ``image`` is ignored (F1); output [b,t,768]."""
    b, _, t, _, _ = raw_image.shape
    pixel = clip_preprocess(raw_image)
    out = model(pixel_values=pixel)
    emb = out.image_embeds                                  # == get_image_features(...).pooler_output (F9)
    return emb.reshape(b, t, -1)


def wav2vec2_normalize(x2d: np.ndarray) -> np.ndarray:
    """HF Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm applied to ONE item: the reference
    passes a 2-D torch tensor, which the extractor does not recognise as batched, so the
    statistics run over the whole [(b t), s] array (population variance, eps 1e-7) -- SURVEY F3."""
    x2d = np.asarray(x2d, dtype=np.float32)
    return (x2d - x2d.mean()) / np.sqrt(x2d.var() + 1e-7)


@torch.no_grad()
def hubert_large_forward(model, raw_audio: torch.Tensor, do_normalize: bool = True,
                         return_hidden_states: bool = False):
    """HUBERT_LARGE.forward (encoder.py:418-432): ``audio`` (mel) ignored; raw_audio [b,t,1,s]."""
    raw = raw_audio[:, :, 0, :]
    b, t, s = raw.shape
    raw = raw.reshape(b * t, s)
    x = raw.cpu().numpy()
    if do_normalize:
        x = wav2vec2_normalize(x)
    input_values = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    hs = model(input_values, output_hidden_states=True).hidden_states       # tuple of [(b t), T, D]
    feature = torch.stack(hs)[[-4, -3, -2, -1]].mean(dim=0)                 # encoder.py:428
    feature = feature.mean(dim=1)                                            # :429
    emb = feature.reshape(b, t, -1)
    if return_hidden_states:
        return emb, hs
    return emb


# --------------------------------------------------------------------------- #
# fusion heads + projectors (shipping "attention" variant: no softmax anywhere)
# --------------------------------------------------------------------------- #
def _lin(state, name, x):
    return torch.nn.functional.linear(x, state[name + ".weight"], state[name + ".bias"])


@torch.no_grad()
def encode_video_attention(state, hidden: torch.Tensor, n_query: int = 1,
                           fusion_type: str = "attention") -> Tuple[torch.Tensor, torch.Tensor]:
    """affectgpt.py:505-535 (realtime) / :543-610 (pre-extracted).  hidden [b,t,768] or [b,768]."""
    if hidden.dim() == 2:                                  # :553-558 pre-extracted single vector
        hidden = hidden.unsqueeze(1)
    b, t, _ = hidden.shape
    store = hidden
    if fusion_type == "attention":
        if t == 1:                                         # :587-589 bypasses the attention
            fused = hidden.squeeze(1)
        else:
            att = _lin(state, "video_attention_mlp", hidden)            # [b,t,1]
            fused = torch.matmul(hidden.transpose(1, 2), att).squeeze(2)
    elif fusion_type == "mean":                            # :601-608
        fused = hidden.mean(dim=1)
    else:
        raise NotImplementedError(fusion_type)
    llm = _lin(state, "affectgpt_proj", fused).unsqueeze(1).expand(-1, n_query, -1)
    return store, llm


@torch.no_grad()
def encode_video_attention_realtime(state, hidden: torch.Tensor, n_query: int = 1):
    """affectgpt.py:505-535: the realtime branch always applies the attention (even t == 1)."""
    att = _lin(state, "video_attention_mlp", hidden)
    fused = torch.matmul(hidden.transpose(1, 2), att).squeeze(2)
    llm = _lin(state, "affectgpt_proj", fused).unsqueeze(1).expand(-1, n_query, -1)
    return hidden, llm


@torch.no_grad()
def encode_audio_attention(state, hidden: torch.Tensor, n_query: int = 1,
                           fusion_type: str = "attention"):
    """affectgpt.py:690-716 / :751-773.  hidden [b,t,1024]."""
    if fusion_type == "attention":
        att = _lin(state, "audio_attention_mlp", hidden)
        fused = torch.matmul(hidden.transpose(1, 2), att).squeeze(2)
    elif fusion_type == "mean":
        fused = hidden.mean(dim=1)
    else:
        raise NotImplementedError(fusion_type)
    llm = _lin(state, "audio_llama_proj", fused).unsqueeze(1).expand(-1, n_query, -1)
    return hidden, llm


@torch.no_grad()
def encode_multi_attention(state, video_hidden: torch.Tensor, audio_hidden: torch.Tensor,
                           n_query: int = 1):
    """affectgpt.py:883-908."""
    v = _lin(state, "multi_video_embs", video_hidden.mean(dim=1))
    a = _lin(state, "multi_audio_embs", audio_hidden.mean(dim=1))
    att = _lin(state, "fc_att", _lin(state, "attention_mlp", torch.cat([v, a], dim=1)))   # [b,2]
    multi = torch.matmul(torch.stack([v, a], dim=2), att.unsqueeze(2)).squeeze(2)
    llm = _lin(state, "multi_llama_proj", multi).unsqueeze(1).expand(-1, n_query, -1)
    return multi, llm


@torch.no_grad()
def fused_tokens(state, frame_h: Optional[torch.Tensor], face_h: Optional[torch.Tensor],
                 audio_h: torch.Tensor, face_or_frame: str = "multiface_audio_face_frame_text"):
    """Token set of AffectGPT.forward (affectgpt.py:924-947) in the prompt order of
    base_dataset.py:882-889: <Multi> <Audio> <Face> <Frame>.  Returns dict name -> [b,1,D]."""
    out = {}
    if frame_h is not None:
        out["frame"] = encode_video_attention_realtime(state, frame_h)[1]
    if face_h is not None:
        out["face"] = encode_video_attention_realtime(state, face_h)[1]
    out["audio"] = encode_audio_attention(state, audio_h)[1]
    if face_or_frame.startswith("multiface"):
        out["multi"] = encode_multi_attention(state, face_h, audio_h)[1]
    elif face_or_frame.startswith("multiframe"):
        out["multi"] = encode_multi_attention(state, frame_h, audio_h)[1]
    return out


# --------------------------------------------------------------------------- #
# synthetic inputs (SURVEY 8d "Synthetic inputs")
# --------------------------------------------------------------------------- #
def synth_frames(b: int, t: int = 8, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (b, 3, t, 224, 224), generator=g).float()


def synth_audio(b: int, t: int = 8, seed: int = 3, samples: int = 32000) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, t, 1, samples, generator=g)
    amp = torch.empty(b, t, 1, 1).uniform_(0.02, 0.5, generator=torch.Generator().manual_seed(seed + 1))
    return x * amp


# --------------------------------------------------------------------------- #
# comparison helpers (SURVEY 8d "Correctness gate")
# --------------------------------------------------------------------------- #
def cosine_rows(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    x = x.double().reshape(-1, x.shape[-1])
    y = y.double().reshape(-1, y.shape[-1])
    return (x * y).sum(-1) / (x.norm(dim=-1) * y.norm(dim=-1) + 1e-30)


def max_rel_err(x: torch.Tensor, y: torch.Tensor) -> float:
    """max|x-y| / max|y| per tensor (element-wise relative error is ill-posed near zero)."""
    return float((x.double() - y.double()).abs().max() / (y.double().abs().max() + 1e-30))


def nn_identity(x: torch.Tensor, y: torch.Tensor) -> bool:
    """argmax_j cos(x_i, y_j) == i on batch-mean-centred rows (discriminative check)."""
    x = x.double().reshape(-1, x.shape[-1])
    y = y.double().reshape(-1, y.shape[-1])
    if x.shape[0] < 2:
        return True
    mu = y.mean(0, keepdim=True)
    xc, yc = x - mu, y - mu
    xc = xc / (xc.norm(dim=-1, keepdim=True) + 1e-30)
    yc = yc / (yc.norm(dim=-1, keepdim=True) + 1e-30)
    sim = xc @ yc.T
    return bool((sim.argmax(dim=1) == torch.arange(x.shape[0])).all())
