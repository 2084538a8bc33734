"""Random-init weights and synthetic inputs of the path's exact shapes (there is no network for checkpoints or
datasets).  Keys/shapes follow SURVEY appendix B (HF CLIPModel vision tower, HF HubertModel large, AffectGPT
ckpt['model'] heads).  Used by bench.py and the smoke test; numerics tests use the oracle's HF-initialised
weights instead."""
from __future__ import annotations

from typing import Dict

import torch


def _n(gen, shape, std, device):
    return (torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * std)


def clip_state(num_layers: int = 24, device="cpu", seed: int = 1234) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    d, ff = 1024, 4096
    sd = {
        "vision_model.embeddings.class_embedding": _n(g, (d,), 0.02, device),
        "vision_model.embeddings.patch_embedding.weight": _n(g, (d, 3, 14, 14), 0.02, device),
        "vision_model.embeddings.position_embedding.weight": _n(g, (257, d), 0.02, device),
        "vision_model.pre_layrnorm.weight": 1 + _n(g, (d,), 0.1, device),
        "vision_model.pre_layrnorm.bias": _n(g, (d,), 0.1, device),
        "vision_model.post_layernorm.weight": 1 + _n(g, (d,), 0.1, device),
        "vision_model.post_layernorm.bias": _n(g, (d,), 0.1, device),
        "visual_projection.weight": _n(g, (768, d), 0.03, device),
    }
    for i in range(num_layers):
        p = f"vision_model.encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = _n(g, (d, d), 0.02, device)
            sd[p + f"self_attn.{nm}.bias"] = _n(g, (d,), 0.05, device)
        for ln in ("layer_norm1", "layer_norm2"):
            sd[p + ln + ".weight"] = 1 + _n(g, (d,), 0.1, device)
            sd[p + ln + ".bias"] = _n(g, (d,), 0.1, device)
        sd[p + "mlp.fc1.weight"] = _n(g, (ff, d), 0.02, device)
        sd[p + "mlp.fc1.bias"] = _n(g, (ff,), 0.05, device)
        sd[p + "mlp.fc2.weight"] = _n(g, (d, ff), 0.01, device)
        sd[p + "mlp.fc2.bias"] = _n(g, (d,), 0.05, device)
    return sd


def hubert_state(num_layers: int = 24, feat_extract_norm: str = "layer", device="cpu", seed: int = 1235) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    d, ff = 1024, 4096
    sd = {}
    kernels = (10, 3, 3, 3, 3, 2, 2)
    for i, k in enumerate(kernels):
        p = f"feature_extractor.conv_layers.{i}."
        cin = 1 if i == 0 else 512
        sd[p + "conv.weight"] = _n(g, (512, cin, k), (2.0 / (cin * k)) ** 0.5, device)
        sd[p + "conv.bias"] = _n(g, (512,), 0.05, device)
        if feat_extract_norm == "layer" or i == 0:
            sd[p + "layer_norm.weight"] = 1 + _n(g, (512,), 0.1, device)
            sd[p + "layer_norm.bias"] = _n(g, (512,), 0.1, device)
    sd["feature_projection.layer_norm.weight"] = 1 + _n(g, (512,), 0.1, device)
    sd["feature_projection.layer_norm.bias"] = _n(g, (512,), 0.1, device)
    sd["feature_projection.projection.weight"] = _n(g, (d, 512), 0.03, device)
    sd["feature_projection.projection.bias"] = _n(g, (d,), 0.05, device)
    sd["encoder.pos_conv_embed.conv.bias"] = _n(g, (d,), 0.05, device)
    v = _n(g, (d, 64, 128), 0.01, device)
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    sd["encoder.layer_norm.weight"] = 1 + _n(g, (d,), 0.1, device)
    sd["encoder.layer_norm.bias"] = _n(g, (d,), 0.1, device)
    for i in range(num_layers):
        p = f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"attention.{nm}.weight"] = _n(g, (d, d), 0.02, device)
            sd[p + f"attention.{nm}.bias"] = _n(g, (d,), 0.05, device)
        for ln in ("layer_norm", "final_layer_norm"):
            sd[p + ln + ".weight"] = 1 + _n(g, (d,), 0.1, device)
            sd[p + ln + ".bias"] = _n(g, (d,), 0.1, device)
        sd[p + "feed_forward.intermediate_dense.weight"] = _n(g, (ff, d), 0.02, device)
        sd[p + "feed_forward.intermediate_dense.bias"] = _n(g, (ff,), 0.05, device)
        sd[p + "feed_forward.output_dense.weight"] = _n(g, (d, ff), 0.01, device)
        sd[p + "feed_forward.output_dense.bias"] = _n(g, (d,), 0.05, device)
    return sd


def fusion_state(llm_dim: int = 3584, device="cpu", seed: int = 1236) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = {"video_attention_mlp": (1, 768), "affectgpt_proj": (llm_dim, 768), "audio_attention_mlp": (1, 1024),
              "audio_llama_proj": (llm_dim, 1024), "multi_video_embs": (1024, 768), "multi_audio_embs": (1024, 1024),
              "attention_mlp": (1024, 2048), "fc_att": (2, 1024), "multi_llama_proj": (llm_dim, 1024)}
    sd = {}
    for k, (o, i) in shapes.items():
        sd[k + ".weight"] = _n(g, (o, i), (1.0 / i) ** 0.5, device)
        sd[k + ".bias"] = _n(g, (o,), 0.02, device)
    return sd


def frames(b: int, t: int = 8, seed: int = 1, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """[b,3,t,224,224] holding 0..255 (the reference's raw_image contract)."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randint(0, 256, (b, 3, t, 224, 224), generator=g, device=device, dtype=torch.uint8)
    return x if dtype == torch.uint8 else x.to(dtype)


def audio(b: int, t: int = 8, seed: int = 3, device="cpu") -> torch.Tensor:
    """[b,t,1,32000]: noise with a per-segment amplitude in [0.02, 0.5] so normalisation scope matters."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((b, t, 1, 32000), generator=g, device=device)
    amp = 0.02 + 0.48 * torch.rand((b, t, 1, 1), generator=g, device=device)
    return x * amp
