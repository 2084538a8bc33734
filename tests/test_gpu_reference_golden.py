"""CUDA path vs fixtures produced by the REFERENCE's own code (tests/golden/make_reference_golden.py): encoder wrapper
classes, attention / mean heads, and the three Q-Former fusion variants.  Everything goes through the
reference-shaped Python surface and therefore through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import qformer as OQ
from oracle import reference_path as R
from oracle import sampling as OS
from tests.gpu_util import gate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FUS = np.load(os.path.join(GOLD, "reference_fusion.npz"))
ENC = np.load(os.path.join(GOLD, "reference_encoders.npz"))


def _inputs():
    g = torch.Generator().manual_seed(7)
    frame_h = torch.randn(3, 8, 768, generator=g)
    face_h = torch.randn(3, 8, 768, generator=g)
    audio_h = torch.randn(3, 8, 1024, generator=g)
    frame_vec = torch.randn(3, 768, generator=g)
    return frame_h.to(DEV), face_h.to(DEV), audio_h.to(DEV), frame_vec.to(DEV)


def _close(got, key, tol):
    want = torch.from_numpy(FUS[key])
    assert tuple(got.shape) == tuple(want.shape), (key, tuple(got.shape), tuple(want.shape))
    err = float((got.float().cpu() - want).abs().max()) / max(float(want.abs().max()), 1e-12)
    print(f"[ref-golden] {key}: rel err {err:.2e}")
    assert err < tol, (key, err)


def test_encoder_classes_match_reference_encoder_py():
    """reference_encoders.npz = CLIP_VIT_LARGE / HUBERT_LARGE of the reference's encoder.py on the oracle's small seeded
    checkpoints; here the drop-in classes of the same names, fed the same state dicts."""
    import affectgpt_b200 as agb
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), R.build_fusion_state(seed=1234), device=DEV, vit_layers=2,
                           hubert_layers=3, norm_scope="batch")
    got = model.visual_encoder(None, (R.synth_frames(1, 2, seed=1) + 0.6).to(DEV))
    gate(got, torch.from_numpy(ENC["clip2_frames_seed1_plus06_b1_t2"]), "reference CLIP_VIT_LARGE.forward")
    for key, audio in (("hubert3_layer_audio_seed3_b1_t2", R.synth_audio(1, 2, seed=3)),
                       ("hubert3_layer_audio_seed5_b2_t2", R.synth_audio(2, 2, seed=5))):
        got = model.acoustic_encoder(None, audio.to(DEV))           # norm_scope="batch": statistics over the call (F3)
        gate(got, torch.from_numpy(ENC[key]), "reference HUBERT_LARGE.forward " + key)


def test_attention_and_mean_heads_match_reference_methods():
    import affectgpt_b200 as agb
    frame_h, face_h, audio_h, frame_vec = _inputs()
    st = R.build_fusion_state(seed=1234)
    m = agb.AffectGPTFusion(lambda a, b: frame_h, lambda a, b: audio_h)
    m.load_state_dict(st)
    _close(m.encode_video_merge(None, frame_h)[1], "att_video_merge_realtime_llm", 1e-4)
    _close(m.encode_video_merge(face_h, None, is_preextracted=True)[1], "att_video_pre3d_llm", 1e-4)
    h, l = m.encode_video_merge(frame_vec, None, is_preextracted=True)
    _close(l, "att_video_pre2d_llm", 1e-4)
    _close(h, "att_video_pre2d_hidden", 1e-6)
    _close(m.encode_audio_merge(None, audio_h)[1], "att_audio_realtime_llm", 1e-4)
    _close(m.encode_audio_merge(audio_h, None, is_preextracted=True)[1], "att_audio_pre_llm", 1e-4)
    mh, ml = m.encode_multi_merge(face_h, audio_h)
    _close(mh, "att_multi_hidden", 1e-4)
    _close(ml, "att_multi_llm", 1e-4)
    mm = agb.AffectGPTFusion(lambda a, b: frame_h, lambda a, b: audio_h, video_fusion_type="mean", audio_fusion_type="mean")
    mm.load_state_dict(st)
    _close(mm.encode_video_merge(None, frame_h)[1], "mean_video_realtime_llm", 1e-4)
    _close(mm.encode_video_merge(face_h, None, is_preextracted=True)[1], "mean_video_pre3d_llm", 1e-4)
    _close(mm.encode_audio_merge(None, audio_h)[1], "mean_audio_realtime_llm", 1e-4)
    _close(mm.encode_audio_merge(audio_h, None, is_preextracted=True)[1], "mean_audio_pre_llm", 1e-4)


def test_qformer_variants_match_reference_qformer_py():
    import affectgpt_b200 as agb
    frame_h, face_h, audio_h, frame_vec = _inputs()
    qs = OQ.fusion_qformer_state(seed=4321, n_query=4)
    m = agb.AffectGPTFusion(lambda a, b: frame_h, lambda a, b: audio_h, video_fusion_type="qformer",
                            audio_fusion_type="qformer", multi_fusion_type="qformer", num_video_query_token=4,
                            num_audio_query_token=4, num_multi_query_token=4)
    m.load_state_dict(qs)
    h, l = m.encode_video_merge(None, frame_h)
    assert torch.equal(h, frame_h)
    _close(l, "qf_video_realtime_llm", 2e-4)
    _close(m.encode_video_merge(face_h, None, is_preextracted=True)[1], "qf_video_pre3d_llm", 2e-4)
    _close(m.encode_video_merge(frame_vec, None, is_preextracted=True)[1], "qf_video_pre2d_llm", 2e-4)
    _close(m.encode_audio_merge(None, audio_h)[1], "qf_audio_realtime_llm", 2e-4)
    mh, ml = m.encode_multi_merge(face_h, audio_h)
    _close(mh, "qf_multi_hidden", 2e-4)
    _close(ml, "qf_multi_llm", 2e-4)
    # a mixed model: Q-Former for the multi token only, shipping attention heads elsewhere
    st = dict(R.build_fusion_state(seed=1234))
    st.update({k: v for k, v in qs.items() if k.startswith("multi_")})
    mix = agb.AffectGPTFusion(lambda a, b: frame_h, lambda a, b: audio_h, multi_fusion_type="qformer",
                              num_multi_query_token=4)
    mix.load_state_dict(st)
    _close(mix.encode_video_merge(None, frame_h)[1], "att_video_merge_realtime_llm", 1e-4)
    _close(mix.encode_multi_merge(face_h, audio_h)[1], "qf_multi_llm", 2e-4)
    with pytest.raises(RuntimeError):                       # 300 tokens exceed multi_position_embedding's 264 rows
        mix.encode_multi_merge(torch.zeros(1, 292, 768, device=DEV), torch.zeros(1, 8, 1024, device=DEV))


def test_face_staging_is_bit_exact_with_cv2_through_reference_load_face():
    """reference_face_resize.npz = the reference's load_face (np.load -> cv2.resize each crop -> sample) on random crops."""
    from affectgpt_b200 import staging
    d = np.load(os.path.join(GOLD, "reference_face_resize.npz"))
    got = staging.load_face(d["faces"], n_frms=3, device=DEV)               # [3(c), 3(t), 224, 224] float
    assert got.dtype == torch.float32 and tuple(got.shape) == (3, 3, 224, 224)
    assert torch.equal(got.cpu().to(torch.uint8), torch.from_numpy(d["resized"]))
    got8 = staging.load_face(d["faces"], n_frms=8, device=DEV)              # 3 crops -> indices [0,1,2,2,2,2,2,2]
    assert torch.equal(got8[:, 3], got8[:, 2]) and torch.equal(got8[:, :3], got)


@pytest.mark.parametrize("h,w", [(224, 224), (112, 112), (448, 300), (31, 17), (223, 225), (1, 1), (2, 3), (720, 1280)])
def test_face_staging_matches_cv2_resize(h, w):
    cv2 = pytest.importorskip("cv2")
    from affectgpt_b200 import staging
    rng = np.random.default_rng(h * 1000 + w)
    faces = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    got = staging.stage_faces(faces, [1, 0, 1], device=DEV).cpu().numpy()
    for t, i in enumerate([1, 0, 1]):
        want = cv2.resize(faces[i], (224, 224)).transpose(2, 0, 1)
        assert np.array_equal(got[t], want), (h, w, t, int(np.abs(got[t].astype(int) - want.astype(int)).max()))


def test_audio_staging_matches_reference_load_audio_slices():
    import json
    from affectgpt_b200 import sampling, staging
    with open(os.path.join(GOLD, "reference_sampling.json")) as f:
        cases = json.load(f)["audio"]
    for case in cases:
        n, ch = case["n"], case["channels"]
        base = torch.arange(n, dtype=torch.float32) + 1.0
        wav = torch.stack([base, base]) if ch == 2 else base[None]
        got = staging.stage_audio(wav, device=DEV).cpu()
        assert tuple(got.shape) == (8, 1, 32000)
        assert [int(x) - 1 for x in got[:, 0, 0]] == case["starts"], case
        assert [int(x) - 1 for x in got[:, 0, -1]] == case["last_sample"], case
        assert torch.equal(got, sampling.segment_audio(wav))                     # whole tensor, bit for bit
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(2, 70001, generator=g)
    assert torch.equal(staging.stage_audio(wav, device=DEV).cpu(), sampling.segment_audio(wav))   # true stereo mix-down


TOK = {"frame": 1001, "face": 1002, "audio": 1003, "multi": 1004, "image": 1005}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("nq", [{"frame": 1, "face": 1, "audio": 1, "multi": 1}, {"frame": 4, "face": 4, "audio": 2, "multi": 3}])
def test_gpu_splice_matches_reference_loop(nq, dtype):
    """oracle.sampling.splice is the literal per-sample restatement of affectgpt.py:956-1009."""
    from affectgpt_b200 import splice as P
    from oracle import sampling as O
    g = torch.Generator().manual_seed(0)
    b, L, D = 5, 300, 3584
    ids = torch.randint(5, 900, (b, L), generator=g)
    for i in range(b - 1):                                   # the last sample carries no placeholder at all
        pos = 3 + 7 * i
        for name in ["multi", "audio", "face", "frame"]:     # base_dataset.py:882-889 prompt order
            ids[i, pos:pos + nq[name]] = TOK[name]
            pos += nq[name] + 2 + i
    emb = torch.randn(b, L, D, generator=g).to(dtype)
    feats = {k: torch.randn(b, nq[k], D, generator=g) for k in nq}
    want = torch.stack([O.splice(emb[i:i + 1].float(), ids[i:i + 1], TOK, nq, {k: v[i:i + 1] for k, v in feats.items()})[0]
                        for i in range(b)])
    got = P.splice_tokens(emb.to(DEV), ids.to(DEV), TOK, nq, {k: v.to(DEV) for k, v in feats.items()})
    assert got.dtype == dtype
    assert torch.equal(got.float().cpu(), want.to(dtype).float())
    assert torch.equal(got[-1].cpu(), emb[-1])


def test_gpu_splice_raises_like_reference():
    from affectgpt_b200 import splice as P
    nq = {"frame": 2, "audio": 1}
    emb = torch.zeros(2, 16, 8, device=DEV)
    feats = {"frame": torch.ones(2, 2, 8, device=DEV), "audio": torch.ones(2, 1, 8, device=DEV)}
    ids = torch.full((2, 16), 7)
    ids[1, 3] = TOK["frame"]                                 # count != n_query in one sample
    with pytest.raises(ValueError, match="number of audio patch tokens"):
        P.splice_tokens(emb, ids, TOK, nq, feats)
    ids[1, 9] = TOK["frame"]                                 # right count, not consecutive
    with pytest.raises(ValueError, match="consecutive"):
        P.splice_tokens(emb, ids, TOK, nq, feats)
    ids2 = torch.full((2, 16), 7)
    ids2[0, 5] = TOK["audio"]
    with pytest.raises(AssertionError):                      # placeholder present, embeds missing
        P.splice_tokens(emb, ids2, TOK, nq, {"frame": feats["frame"]})
    ids3 = torch.full((2, 16), 7)
    out = P.splice_tokens(emb, ids3, TOK, nq, feats)         # no placeholders: identity
    assert torch.equal(out, emb)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["nq1", "nq_many", "other_order", "sample_without_placeholder", "frame_audio_only",
                                  "err_count", "err_consecutive", "err_missing_face"])
def test_gpu_splice_matches_reference_forward_fixture(name, dtype):
    """reference_splice.npz: inputs_embeds produced by the reference's own AffectGPT.forward loop (affectgpt.py:956-1009,
    tests/golden/make_reference_splice.py).  agb_splice_tokens reproduces it bit for bit and raises what it raises."""
    from affectgpt_b200 import splice as P
    from tests import splice_cases as SC
    c = SC.stub_case(SC.fixture(), name)
    emb = c["embeds"].to(dtype).to(DEV)
    toks = {k: v.to(DEV) for k, v in c["tokens"].items()}
    if c["error"] is not None:
        exc = {"ValueError": ValueError, "AssertionError": AssertionError}[c["error"]["type"]]
        with pytest.raises(exc) as e:
            P.splice_tokens(emb, c["ids"].to(DEV), SC.TOK, c["nq"], toks)
        assert str(e.value) == c["error"]["message"]
        return
    got = P.splice_tokens(emb, c["ids"].to(DEV), SC.TOK, c["nq"], toks)
    assert got.dtype == dtype
    assert torch.equal(got.float().cpu(), c["want"].to(dtype).float())


def test_gpu_fused_tokens_spliced_like_reference_forward():
    """'heads' case: reference encode_*_merge + its splice loop, end to end (pre-extracted features -> inputs_embeds).
    Product: AffectGPTFusion heads (agb_*_fuse) + agb_splice_tokens.  Text rows bit-exact, fused rows to 1e-4."""
    import affectgpt_b200 as agb
    from affectgpt_b200 import splice as P
    from tests import splice_cases as SC
    fix = SC.fixture()
    ids = torch.from_numpy(fix["heads.ids"])
    want = torch.from_numpy(fix["heads.inputs_embeds"])
    emb = SC.embed_ids(ids, 3584, int(fix["heads.embed_seed"]))
    frame_h, face_h, audio_h = (torch.from_numpy(fix[f"heads.{k}"]).to(DEV) for k in ("frame_h", "face_h", "audio_h"))
    m = agb.AffectGPTFusion(None, None)
    m.load_state_dict(R.build_fusion_state(seed=1234))
    _, frame_llm = m.encode_video_merge(frame_h, None, is_preextracted=True)
    face_hid, face_llm = m.encode_video_merge(face_h, None, is_preextracted=True)
    audio_hid, audio_llm = m.encode_audio_merge(audio_h, None, is_preextracted=True)
    _, multi_llm = m.encode_multi_merge(face_hid, audio_hid)
    nq = {"frame": 1, "face": 1, "audio": 1, "multi": 1}
    got = P.splice_tokens(emb.to(DEV), ids.to(DEV), SC.TOK, nq, {"frame": frame_llm, "face": face_llm, "audio": audio_llm,
                                                                  "multi": multi_llm}).cpu()
    ph = torch.zeros_like(ids, dtype=torch.bool)
    for t in SC.TOK.values():
        ph |= ids == t
    assert int(ph.sum()) == 8
    assert torch.equal(got[~ph], want[~ph])
    err = float((got[ph] - want[ph]).abs().max()) / float(want[ph].abs().max())
    print(f"[ref-golden] spliced fused tokens vs reference forward: rel err {err:.2e}")
    assert err < 1e-4, err


@pytest.mark.parametrize("orig,n,ch", [(44100, 97531, 1), (48000, 120000, 2), (22050, 30011, 1), (8000, 15999, 1), (11025, 40000, 2)])
def test_gpu_resample_matches_torchaudio(orig, n, ch):
    """agb_resample_audio vs torchaudio.functional.resample (what load_audio calls, data.py:186-187): same filter bank,
    fp32 FIR - equal up to the summation order; then the staged 2-second slices of the resampled waveform."""
    F = pytest.importorskip("torchaudio.functional")
    from affectgpt_b200 import staging
    g = torch.Generator().manual_seed(orig + n)
    wav = torch.randn(ch, n, generator=g) * 0.3
    want = F.resample(wav, orig_freq=orig, new_freq=16000)
    got = staging.resample(wav, orig, 16000, DEV).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))
    segs = staging.stage_audio(wav, device=DEV, orig_sample_rate=orig).cpu()
    ref_segs, _ = OS.audio_segments(want)
    assert segs.shape == ref_segs.shape
    assert float((segs - ref_segs).abs().max()) <= 2e-6


def test_load_audio_drop_in(tmp_path):
    """staging.load_audio(paths, device) == the reference's load_audio: one 16 kHz mono file, one 44.1 kHz stereo file,
    one shorter than 2 s; compared with torchaudio.functional.resample + the restated slicing."""
    F = pytest.importorskip("torchaudio.functional")
    from scipy.io import wavfile
    from affectgpt_b200 import staging
    rng = np.random.default_rng(3)
    specs = [(16000, 55357, 1), (44100, 150000, 2), (16000, 20000, 1)]
    paths, want = [], []
    for i, (sr, n, ch) in enumerate(specs):
        pcm = (rng.standard_normal((n, ch)) * 4000).astype(np.int16)
        path = str(tmp_path / f"a{i}.wav")
        wavfile.write(path, sr, pcm if ch == 2 else pcm[:, 0])
        paths.append(path)
        wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).T.contiguous()
        if sr != 16000:
            wav = F.resample(wav, orig_freq=sr, new_freq=16000)
        want.append(OS.audio_segments(wav)[0])
    got = staging.load_audio(paths, DEV)
    assert got.device.type == "cuda" and tuple(got.shape) == (3, 8, 1, 32000)
    assert float((got.cpu() - torch.stack(want)).abs().max()) <= 2e-6
    assert staging.load_audio(None, DEV) is None
