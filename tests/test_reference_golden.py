"""Parity pinned against the REFERENCE's own code: fixtures in tests/golden/reference_*.{json,npz} were produced by
executing /root/reference/AffectGPT's unmodified sources in the build container (tests/golden/make_reference_golden.py
+ ref_loader.py).  Here the oracle (oracle/) and the product's host code (affectgpt_b200/) are held to them:
integer contracts bit-exact, fp32 fusion math to 1e-5.  CPU only."""
import json
import os
import random

import numpy as np
import pytest
import torch

from affectgpt_b200 import sampling as S
from oracle import qformer as OQ
from oracle import reference_path as R
from oracle import sampling as OS

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_sampling.json")) as f:
    SAMP = json.load(f)
FUS = np.load(os.path.join(HERE, "golden", "reference_fusion.npz"))


@pytest.mark.parametrize("impl", ["host", "oracle"])
def test_uniform_indices_match_reference(impl):
    fn = S.uniform_frame_indices if impl == "host" else OS.uniform_indices
    for n, table in ((8, SAMP["uniform8"]), (16, SAMP["uniform16"])):
        for vlen, want in table.items():
            assert fn(int(vlen), n) == want, (vlen, n)
    for vlen, want in SAMP["face_uniform8"].items():            # load_face uses the same rule (:272-283)
        assert fn(int(vlen), 8) == want


@pytest.mark.parametrize("impl", ["host", "oracle"])
def test_emotion_peak_indices_match_reference(impl):
    peak = S.emotion_peak_frame_indices if impl == "host" else OS.emotion_peak_indices
    smart = S.smart_frame_indices if impl == "host" else OS.smart_frame_indices
    for vlen, want in SAMP["emotion_peak_noinfo"].items():
        assert peak(int(vlen), None, 8) == want
    for case in SAMP["smart"]:
        assert smart(case["au_info"], case["total_video_frames"]) == case["indices"], case
    for case in SAMP["emotion_peak_json"]:
        assert peak(case["vlen"], case["au_info"], 8) == case["indices"], case


@pytest.mark.parametrize("impl", ["host", "oracle"])
def test_headtail_indices_match_reference_rng_stream(impl):
    fn = S.headtail_frame_indices if impl == "host" else OS.headtail_indices
    for case in SAMP["headtail"]:
        random.seed(case["seed"])
        assert fn(case["vlen"], 8) == case["indices"], case


@pytest.mark.parametrize("impl", ["host", "oracle"])
def test_audio_segments_match_reference(impl):
    for case in SAMP["audio"]:
        n, ch = case["n"], case["channels"]
        base = torch.arange(n, dtype=torch.float32) + 1.0        # sample k holds k + 1, zero = padding
        wav = torch.stack([base, base]) if ch == 2 else base[None]
        clips = S.segment_audio(wav) if impl == "host" else OS.audio_segments(wav)[0]
        assert tuple(clips.shape) == (8, 1, 32000)
        assert [int(x) - 1 for x in clips[:, 0, 0]] == case["starts"], case
        assert [int(x) - 1 for x in clips[:, 0, -1]] == case["last_sample"], case
        if impl == "host":
            bounds = S.audio_segment_bounds(n)
            assert [a for a, _ in bounds] == [max(s, 0) for s in case["starts"]]


def _inputs():
    g = torch.Generator().manual_seed(7)
    frame_h = torch.randn(3, 8, 768, generator=g)
    face_h = torch.randn(3, 8, 768, generator=g)
    audio_h = torch.randn(3, 8, 1024, generator=g)
    frame_vec = torch.randn(3, 768, generator=g)
    return frame_h, face_h, audio_h, frame_vec


def _close(got, key, tol=1e-5):
    want = torch.from_numpy(FUS[key])
    assert tuple(got.shape) == tuple(want.shape), (key, got.shape, want.shape)
    err = float((got - want).abs().max()) / max(float(want.abs().max()), 1e-12)
    assert err < tol, (key, err)


def test_oracle_attention_and_mean_heads_match_reference():
    st = R.build_fusion_state(seed=1234)
    frame_h, face_h, audio_h, frame_vec = _inputs()
    _close(R.encode_video_attention_realtime(st, frame_h)[1], "att_video_realtime_llm")
    _close(R.encode_video_attention_realtime(st, frame_h)[1], "att_video_merge_realtime_llm")
    _close(R.encode_video_attention(st, face_h)[1], "att_video_pre3d_llm")
    _close(R.encode_video_attention(st, frame_vec)[1], "att_video_pre2d_llm")
    _close(R.encode_video_attention(st, frame_vec)[0], "att_video_pre2d_hidden")
    _close(R.encode_audio_attention(st, audio_h)[1], "att_audio_realtime_llm")
    _close(R.encode_audio_attention(st, audio_h)[1], "att_audio_pre_llm")
    mh, ml = R.encode_multi_attention(st, face_h, audio_h)
    _close(mh, "att_multi_hidden")
    _close(ml, "att_multi_llm")
    _close(R.encode_video_attention(st, frame_h, fusion_type="mean")[1], "mean_video_realtime_llm")
    _close(R.encode_video_attention(st, face_h, fusion_type="mean")[1], "mean_video_pre3d_llm")
    _close(R.encode_audio_attention(st, audio_h, fusion_type="mean")[1], "mean_audio_realtime_llm")
    _close(R.encode_audio_attention(st, audio_h, fusion_type="mean")[1], "mean_audio_pre_llm")


def test_oracle_qformer_variants_match_reference():
    qs = OQ.fusion_qformer_state(seed=4321, n_query=4)
    frame_h, face_h, audio_h, frame_vec = _inputs()
    with torch.no_grad():
        _close(OQ.encode_video_qformer(qs, frame_h)[1], "qf_video_realtime_llm")
        _close(OQ.encode_video_qformer(qs, face_h)[1], "qf_video_pre3d_llm")
        _close(OQ.encode_video_qformer(qs, frame_vec.unsqueeze(1))[1], "qf_video_pre2d_llm")
        _close(OQ.encode_audio_qformer(qs, audio_h)[1], "qf_audio_realtime_llm")
        mh, ml = OQ.encode_multi_qformer(qs, face_h, audio_h)
        _close(mh, "qf_multi_hidden")
        _close(ml, "qf_multi_llm")


def test_oracle_encoder_wrappers_match_reference_classes():
    """reference_encoders.npz = outputs of the reference's CLIP_VIT_LARGE / HUBERT_LARGE classes (encoder.py) on the
    oracle's small seeded checkpoints: pins uint8 truncation, call-wide waveform normalisation, last-4 / time mean."""
    enc = np.load(os.path.join(HERE, "golden", "reference_encoders.npz"))
    tower = R.build_clip_vision(num_layers=2, seed=1234)
    got = R.clip_vit_large_forward(tower, R.synth_frames(1, 2, seed=1) + 0.6)
    assert float((got - torch.from_numpy(enc["clip2_frames_seed1_plus06_b1_t2"])).abs().max()) < 1e-4
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    for key, audio in (("hubert3_layer_audio_seed3_b1_t2", R.synth_audio(1, 2, seed=3)),
                       ("hubert3_layer_audio_seed5_b2_t2", R.synth_audio(2, 2, seed=5))):
        got = R.hubert_large_forward(hub, audio)
        assert float((got - torch.from_numpy(enc[key])).abs().max()) < 1e-4, key


# ------------------------------------------------------------------------------------------------ token layout (a15)
from affectgpt_b200 import splice as PS  # noqa: E402
from tests import splice_cases as SC  # noqa: E402

SPL = SC.fixture()


@pytest.mark.parametrize("name", SC.STUB_CASES)
def test_splice_oracle_and_host_positions_match_reference_forward(name):
    """reference_splice.npz = inputs_embeds of the reference's own AffectGPT.forward (affectgpt.py:956-1009) run in the
    build container; the oracle restatement reproduces it bit for bit and raises the same exceptions, and the host-side
    position scan of the product agrees with where the reference put the tokens."""
    c = SC.stub_case(SPL, name)
    if c["error"] is not None:
        exc = {"ValueError": ValueError, "AssertionError": AssertionError}[c["error"]["type"]]
        with pytest.raises(exc) as e1:
            OS.splice(c["embeds"], c["ids"], SC.TOK, c["nq"], c["tokens"])
        assert str(e1.value) == c["error"]["message"]
        with pytest.raises(exc) as e2:
            PS.placeholder_positions(c["ids"], SC.TOK, c["nq"], list(c["tokens"]))
        assert str(e2.value) == c["error"]["message"]
        return
    got = OS.splice(c["embeds"], c["ids"], SC.TOK, c["nq"], c["tokens"])
    assert torch.equal(got, c["want"])
    pos = PS.placeholder_positions(c["ids"], SC.TOK, c["nq"], list(c["tokens"]))
    for k, starts in pos.items():
        for i, s in enumerate(starts.tolist()):
            if s >= 0:                                          # the reference's output carries the tokens exactly there
                assert torch.equal(c["want"][i, s:s + c["nq"][k]], c["tokens"][k][i]), (name, k, i)


def test_splice_after_reference_heads_end_to_end():
    """'heads' case: the reference's real encode_*_merge (attention heads + projectors to 3584) in front of its loop.
    Oracle heads + oracle splice vs that output: layout bit-exact (text rows equal), fused rows to 1e-5."""
    st = R.build_fusion_state(seed=1234)
    ids = torch.from_numpy(SPL["heads.ids"])
    want = torch.from_numpy(SPL["heads.inputs_embeds"])
    emb = SC.embed_ids(ids, 3584, int(SPL["heads.embed_seed"]))
    frame_h, face_h, audio_h = (torch.from_numpy(SPL[f"heads.{k}"]) for k in ("frame_h", "face_h", "audio_h"))
    toks = R.fused_tokens(st, frame_h, face_h, audio_h, "multiface_audio_face_frame_text")
    nq = {"frame": 1, "face": 1, "audio": 1, "multi": 1}
    got = OS.splice(emb, ids, SC.TOK, nq, toks)
    placeholder = torch.zeros_like(ids, dtype=torch.bool)
    for t in SC.TOK.values():
        placeholder |= ids == t
    assert int(placeholder.sum()) == 8
    assert torch.equal(got[~placeholder], want[~placeholder])
    err = float((got[placeholder] - want[placeholder]).abs().max()) / float(want[placeholder].abs().max())
    assert err < 1e-5, err
