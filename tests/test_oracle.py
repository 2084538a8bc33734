"""Pin the oracle: (1) its restated preprocessing equals the real HF processors, (2) its outputs equal the
committed golden fixtures (tests/golden/make_golden.py), (3) the F3/F4 semantics the survey found."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_path as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_clip_preprocess_equals_hf_image_processor():
    from PIL import Image
    from transformers import CLIPImageProcessor
    proc = CLIPImageProcessor(size={"shortest_edge": 224}, crop_size={"height": 224, "width": 224},
                              image_mean=list(R.CLIP_MEAN), image_std=list(R.CLIP_STD))
    raw = R.synth_frames(1, 2, seed=5) + 0.6                                      # non-integer values: truncation matters
    frames = raw.permute(0, 2, 3, 4, 1).reshape(2, 224, 224, 3)
    pil = [Image.fromarray(f.numpy().astype(np.uint8)) for f in frames]            # encoder.py:34-39
    want = proc(images=pil, return_tensors="pt")["pixel_values"]
    got = R.clip_preprocess(raw)
    assert got.shape == want.shape and float((got - want).abs().max()) < 1e-5


def test_wav2vec2_norm_quirk_equals_hf_feature_extractor():
    from transformers import Wav2Vec2FeatureExtractor
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                  return_attention_mask=False)
    raw = R.synth_audio(2, 3, seed=9)[:, :, 0, :].reshape(6, 32000)
    iv = fe(raw, sampling_rate=16000, return_tensors="pt").input_values           # encoder.py:425: 2-D tensor = ONE item
    assert tuple(iv.shape) == (1, 6, 32000)
    got = R.wav2vec2_normalize(raw.numpy())
    assert np.allclose(iv[0].numpy(), got, atol=1e-5)
    per_row = (raw - raw.mean(1, keepdim=True)) / (raw.var(1, unbiased=False, keepdim=True) + 1e-7).sqrt()
    assert not np.allclose(per_row.numpy(), got, atol=1e-2)                       # NOT per-segment statistics (F3)


def test_golden_small_reproduces():
    gold = np.load(os.path.join(GOLD, "golden_small.npz"))
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    got = R.clip_vit_large_forward(clip, R.synth_frames(1, 2, seed=1)).numpy()
    assert got.shape == (1, 2, 768)
    np.testing.assert_allclose(got, gold["clip2_frames_seed1_b1_t2"], rtol=1e-4, atol=1e-5)
    for norm in ("layer", "group"):
        hub = R.build_hubert(num_layers=3, feat_extract_norm=norm, seed=1234)
        got = R.hubert_large_forward(hub, R.synth_audio(1, 2, seed=3)).numpy()
        assert got.shape == (1, 2, 1024)
        np.testing.assert_allclose(got, gold[f"hubert3_{norm}_audio_seed3_b1_t2"], rtol=1e-4, atol=1e-5)
    st = R.build_fusion_state(seed=1234)
    g = torch.Generator().manual_seed(7)
    fh, ch, ah = torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 1024, generator=g)
    toks = R.fused_tokens(st, fh, ch, ah, "multiface_audio_face_frame_text")
    assert set(toks) == {"frame", "face", "audio", "multi"}
    for k, v in toks.items():
        assert tuple(v.shape) == (2, 1, 3584)
        np.testing.assert_allclose(v.numpy(), gold[f"fusion_seed7_{k}"], rtol=1e-4, atol=1e-5)


def test_hidden_state_taps_f4():
    """Feature = mean of hidden_states[-4:] then mean over time; the last one is the final-LN output (F4)."""
    hub = R.build_hubert(num_layers=3, seed=1234)
    emb, hs = R.hubert_large_forward(hub, R.synth_audio(1, 1, seed=3), return_hidden_states=True)
    assert len(hs) == 4 and tuple(hs[0].shape) == (1, 99, 1024)
    manual = (hs[0] + hs[1] + hs[2] + hs[3]) / 4
    assert torch.allclose(manual.mean(1), emb[0], atol=1e-5)
    # hs[3] is the encoder's final LayerNorm output, hs[0..2] are raw residual streams (not normalised)
    assert abs(float(hs[3].mean())) < 1.0


def test_fusion_has_no_softmax_and_preextracted_bypass():
    st = R.build_fusion_state(seed=1)
    h = torch.randn(2, 8, 768, generator=torch.Generator().manual_seed(0))
    _, llm = R.encode_video_attention_realtime(st, h)
    att = h @ st["video_attention_mlp.weight"].T + st["video_attention_mlp.bias"]
    fused = (h * att).sum(1)                                                        # linear in h*att: no softmax
    want = fused @ st["affectgpt_proj.weight"].T + st["affectgpt_proj.bias"]
    assert torch.allclose(llm[:, 0], want, atol=1e-4)
    one = h[:, :1]
    _, llm1 = R.encode_video_attention(st, one[:, 0])                               # [b,768] pre-extracted
    want1 = one[:, 0] @ st["affectgpt_proj.weight"].T + st["affectgpt_proj.bias"]   # attention bypassed (:587-589)
    assert torch.allclose(llm1[:, 0], want1, atol=1e-4)
