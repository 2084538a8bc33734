"""Host-side mirror of the reference interface: registry, wrapper signatures, error behaviour (CPU only)."""
import inspect

import pytest
import torch

import affectgpt_b200 as agb
from affectgpt_b200.registry import Registry, registry


def test_registry_keys_and_zero_arg_constructors():
    # A/my_affectgpt/models/encoder.py:177, 395 register these names; callers do get_*_class(name)()
    assert registry.get_visual_encoder_class("CLIP_VIT_LARGE") is agb.CLIP_VIT_LARGE
    assert registry.get_acoustic_encoder_class("HUBERT_LARGE") is agb.HUBERT_LARGE
    assert registry.get_model_class("affectgpt") is agb.AffectGPTFusion
    assert registry.get_visual_encoder_class("nope") is None
    v = registry.get_visual_encoder_class("CLIP_VIT_LARGE")()
    a = registry.get_acoustic_encoder_class("HUBERT_LARGE")()
    assert v.hidden_size == 768 and a.hidden_size == 1024          # encoder.py:192, 411
    assert isinstance(v, torch.nn.Module) and v.to("cpu").eval() is v
    assert list(v.named_parameters()) == []                         # frozen in the reference; none here


def test_duplicate_registration_raises_keyerror():
    with pytest.raises(KeyError):                                    # registry.py:147-152
        @Registry.register_visual_encoder("CLIP_VIT_LARGE")
        class Dup:  # noqa
            pass


def test_forward_signatures_match_reference():
    assert list(inspect.signature(agb.CLIP_VIT_LARGE.forward).parameters) == ["self", "image", "raw_image"]
    assert list(inspect.signature(agb.HUBERT_LARGE.forward).parameters) == ["self", "audio", "raw_audio"]
    sig = inspect.signature(agb.AffectGPTFusion.encode_video_merge)
    assert list(sig.parameters) == ["self", "video", "raw_video", "is_preextracted"]
    assert sig.parameters["is_preextracted"].default is False
    assert list(inspect.signature(agb.AffectGPTFusion.encode_audio_merge).parameters) == ["self", "audio", "raw_audio", "is_preextracted"]
    assert list(inspect.signature(agb.AffectGPTFusion.encode_multi_merge).parameters) == ["self", "video_hidden_state", "audio_hidden_state"]


def test_cpu_inputs_fail_loudly():
    v = agb.CLIP_VIT_LARGE(model_dir="")
    with pytest.raises(RuntimeError):
        v(None, torch.zeros(1, 3, 8, 224, 224))
    with pytest.raises(ValueError):
        v(None, torch.zeros(1, 3, 8, 200, 200))
    a = agb.HUBERT_LARGE(model_dir="")
    with pytest.raises(RuntimeError):
        a(None, torch.zeros(1, 8, 1, 32000))
    with pytest.raises(ValueError):
        a(None, torch.zeros(1, 8, 1, 16000))


def test_model_errors_match_reference():
    m = agb.AffectGPTFusion(None, None)
    with pytest.raises(RuntimeError):                                # affectgpt.py:614-615
        m.encode_video_merge(torch.zeros(1, 3, 8, 224, 224), torch.zeros(1, 3, 8, 224, 224))
    with pytest.raises(RuntimeError):                                # :695-696
        m.encode_audio_merge(None, torch.zeros(1, 8, 1, 32000))
    with pytest.raises(ValueError):                                  # :559-560
        m.encode_video_merge(torch.zeros(1, 2, 3, 4), None, is_preextracted=True)
    with pytest.raises(ValueError):
        agb.AffectGPTFusion(None, None, video_fusion_type="transformer")
    qf = agb.AffectGPTFusion(None, None, video_fusion_type="qformer", num_video_query_token=4)     # all three variants exist
    assert qf._engine_cfg["qf_video_queries"] == 4 and "video_attention_mlp" not in qf._needed_heads()
    missing, unexpected = m.load_state_dict({"llama_model.x": torch.zeros(1), "fc_att.weight": torch.zeros(2, 1024)})
    assert "fc_att.bias" in missing and unexpected == ["llama_model.x"]


def test_product_does_not_import_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "affectgpt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_resample_filter_bank_is_torchaudios():
    """staging.resample_bank restates torchaudio.functional.resample's windowed-sinc bank (the call load_audio makes at
    ImageBind/data.py:186-187) with the same float32 operations: identical taps, width and phase count."""
    import math
    import pytest
    import torch
    from affectgpt_b200 import staging
    F = pytest.importorskip("torchaudio.functional.functional")
    for orig, new in ((44100, 16000), (48000, 16000), (22050, 16000), (8000, 16000), (11025, 16000), (32000, 16000)):
        bank, o, n, width = staging.resample_bank(orig, new)
        ref, ref_width = F._get_sinc_resample_kernel(orig, new, math.gcd(orig, new), dtype=torch.float32)
        assert width == ref_width and (o, n) == (orig // math.gcd(orig, new), new // math.gcd(orig, new))
        assert bank.shape == (n, 2 * width + o) and torch.equal(bank, ref[:, 0, :])
    with pytest.raises(ValueError):
        staging.resample_bank(44100.5, 16000)


def test_encoder_to_accepts_devices_and_rejects_dtypes():
    """Callers of the reference do ``.to(device)`` / ``.eval()``; a dtype cast cannot be honoured by the drop-in (bf16 / fp32
    inside libagb, fp32 at the boundary) and must raise instead of being ignored."""
    import pytest
    import torch
    import affectgpt_b200 as agb
    enc = agb.CLIP_VIT_LARGE(model_dir="")
    assert enc.to("cuda:0") is enc and enc.to(torch.device("cpu")) is enc and enc.eval() is enc
    assert enc.to(torch.float32) is enc and enc.float() is enc and enc.to(device="cuda", dtype=torch.float32) is enc
    for bad in (lambda: enc.to(torch.float16), lambda: enc.half(), lambda: enc.bfloat16(), lambda: enc.double(),
                lambda: enc.to("cuda:0", torch.bfloat16), lambda: enc.to(dtype=torch.float16),
                lambda: enc.to(torch.zeros(1, dtype=torch.float16))):
        with pytest.raises(TypeError):
            bad()
    aud = agb.HUBERT_LARGE(model_dir="")
    with pytest.raises(TypeError):
        aud.half()
