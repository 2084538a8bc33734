"""BASELINE.json's full sizes (24-layer encoders, C2: 64 clips; C4: three streams, 16 frames + 16 faces + 16 x 2 s),
checked through size-independent properties - the oracle needs ~5 s per clip on the host, so it only anchors a
two-clip sample here; the rest is invariance: a batch equals its clips one at a time (the data-parallel premise, and
what makes W GPUs reproduce one GPU bit for bit), permuting clips permutes outputs, chunking does not change a bit,
every clip is closest to itself among the batch (nearest-neighbour identity on mean-centred features)."""
import pytest
import torch

from oracle import reference_path as R
from tests.gpu_util import gate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def full_path():
    import affectgpt_b200 as agb
    from affectgpt_b200 import synthetic as S
    return agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV)


def _tokens(model, frames, audio, faces=None):
    samples = {"raw_frames": frames, "raw_audios": audio,
               "face_or_frame": "multiface_audio_face_frame_text" if faces is not None else "multiframe_audio_frame_text"}
    if faces is not None:
        samples["raw_faces"] = faces
    out = model.encode_clip_batch(samples)
    keys = [k for k in ("multi", "audio", "face", "frame") if k in out]
    return torch.cat([out[k] for k in keys], dim=1), out


def test_c2_batch_64_invariants(full_path):
    from affectgpt_b200 import synthetic as S
    B = 64
    frames = S.frames(B, 8, seed=31).to(DEV)
    audio = S.audio(B, 8, seed=32).to(DEV)
    toks, out = _tokens(full_path, frames, audio)
    assert tuple(toks.shape) == (B, 3, 3584) and bool(torch.isfinite(toks).all())
    for i in (0, 17, 63):                                           # batch == clip by clip, bit for bit
        ti, _ = _tokens(full_path, frames[i:i + 1], audio[i:i + 1])
        assert torch.equal(ti, toks[i:i + 1]), i
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(5)).to(DEV)
    tp, _ = _tokens(full_path, frames[perm], audio[perm])           # permutation equivariance
    assert torch.equal(tp, toks[perm])
    for key in ("frame_hiddens", "audio_hiddens"):                  # nearest-neighbour identity on centred features
        h = out[key].flatten(1).float()
        h = h - h.mean(dim=0, keepdim=True)
        h = torch.nn.functional.normalize(h, dim=1)
        assert bool(((h @ h.T).argmax(dim=1) == torch.arange(B, device=DEV)).all()), key
        assert float((h @ h.T - torch.eye(B, device=DEV)).abs().max()) < 0.999, key      # no two clips collapse


def test_c2_two_clip_sample_against_oracle(full_path):
    """Anchor of the full-size run: 2 of the clips through the fp32 oracle on identical weights (BASELINE gate)."""
    from affectgpt_b200 import synthetic as S
    from transformers import CLIPVisionModelWithProjection, HubertModel
    clip = R.build_clip_vision(24, seed=1)
    clip.load_state_dict({k: v.cpu() for k, v in S.clip_state(24, DEV).items()}, strict=False)
    hub = R.build_hubert(24, "layer", seed=1)
    missing, unexpected = hub.load_state_dict({k: v.cpu() for k, v in S.hubert_state(24, "layer", DEV).items()}, strict=False)
    assert not unexpected and all("masked_spec_embed" in k for k in missing), (missing, unexpected)
    fus = {k: v.cpu() for k, v in S.fusion_state(3584, DEV).items()}
    frames, audio = S.frames(2, 8, seed=31), S.audio(2, 8, seed=32)
    toks, _ = _tokens(full_path, frames.to(DEV), audio.to(DEV))
    for i in range(2):
        fh = R.clip_vit_large_forward(clip, frames[i:i + 1])
        ah = R.hubert_large_forward(hub, audio[i:i + 1])
        want = R.fused_tokens(fus, fh, None, ah, "multiframe_audio_frame_text")
        ref = torch.cat([want["multi"], want["audio"], want["frame"]], dim=1)
        gate(toks[i:i + 1].reshape(3, 3584), ref.reshape(3, 3584), f"full-size C2 clip {i} tokens")
    assert isinstance(clip, CLIPVisionModelWithProjection) and isinstance(hub, HubertModel)


def test_c4_three_stream_shape_invariants(full_path):
    """C4: 16 frames + 16 face crops + 16 x 2 s of audio per clip, three-stream pre-fusion (face + audio -> multi)."""
    from affectgpt_b200 import synthetic as S
    B = 6
    frames = S.frames(B, 16, seed=41).to(DEV)
    faces = S.frames(B, 16, seed=42).to(DEV)
    audio = S.audio(B, 16, seed=43).to(DEV)
    toks, out = _tokens(full_path, frames, audio, faces)
    assert tuple(toks.shape) == (B, 4, 3584) and bool(torch.isfinite(toks).all())     # <Multi> <Audio> <Face> <Frame>
    assert tuple(out["frame_hiddens"].shape) == (B, 16, 768) and tuple(out["audio_hiddens"].shape) == (B, 16, 1024)
    t1, _ = _tokens(full_path, frames[2:3], audio[2:3], faces[2:3])
    assert torch.equal(t1, toks[2:3])
    swapped, _ = _tokens(full_path, faces, audio, frames)           # frame and face streams share weights (:929-932)
    assert torch.equal(swapped[:, 3], toks[:, 2]) and torch.equal(swapped[:, 2], toks[:, 3])


def test_chunked_engine_is_bit_identical_at_full_size():
    import affectgpt_b200 as agb
    from affectgpt_b200 import synthetic as S
    small = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV,
                           max_frames_per_chunk=48, max_segments_per_chunk=40)
    big = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV)
    frames, audio = S.frames(16, 8, seed=51).to(DEV), S.audio(16, 8, seed=52).to(DEV)
    a, _ = _tokens(small, frames, audio)                            # 128 frames in chunks of 48, 128 segments in 40s
    b, _ = _tokens(big, frames, audio)
    assert torch.equal(a, b)


def test_graphed_step_is_bit_identical_and_reusable(full_path):
    import affectgpt_b200 as agb
    from affectgpt_b200 import synthetic as S
    B = 4
    g = agb.GraphedClipBatch(full_path, B, 8, 8, device=DEV)
    assert g.kernels_per_replay > 250                               # 5 kernels per encoder block + front-ends + heads
    for seed in (61, 62):                                           # two different batches through the same graph
        frames, audio = S.frames(B, 8, seed=seed).to(DEV), S.audio(B, 8, seed=seed + 10).to(DEV)
        out = g(frames, audio)
        want, _ = _tokens(full_path, frames, audio)
        got = torch.cat([out["multi"], out["audio"], out["frame"]], dim=1)
        assert torch.equal(got, want), seed


# ------------------------------------------------------------------------------------------------ oracle anchors at depth
@pytest.fixture(scope="module")
def oracle_models():
    """The fp32 oracle (HF modules) loaded with the same synthetic 24-layer checkpoints as ``full_path``."""
    from affectgpt_b200 import synthetic as S
    clip = R.build_clip_vision(24, seed=1)
    clip.load_state_dict({k: v.cpu() for k, v in S.clip_state(24, DEV).items()}, strict=False)
    hub = R.build_hubert(24, "layer", seed=1)
    missing, unexpected = hub.load_state_dict({k: v.cpu() for k, v in S.hubert_state(24, "layer", DEV).items()}, strict=False)
    assert not unexpected and all("masked_spec_embed" in k for k in missing), (missing, unexpected)
    fus = {k: v.cpu() for k, v in S.fusion_state(3584, DEV).items()}
    return clip, hub, fus


def test_c4_clip_against_oracle(full_path, oracle_models):
    """One C4-shaped clip - 16 frames + 16 face crops + 16 x 2 s of audio, three-stream pre-fusion - through the fp32
    oracle on identical weights (BASELINE gate on every token and on the per-modality hiddens)."""
    from affectgpt_b200 import synthetic as S
    clip, hub, fus = oracle_models
    frames, faces, audio = S.frames(1, 16, seed=71), S.frames(1, 16, seed=72), S.audio(1, 16, seed=73)
    toks, out = _tokens(full_path, frames.to(DEV), audio.to(DEV), faces.to(DEV))
    fh, ch, ah = R.clip_vit_large_forward(clip, frames), R.clip_vit_large_forward(clip, faces), R.hubert_large_forward(hub, audio)
    gate(out["frame_hiddens"], fh, "C4 frame hiddens [1,16,768]")
    gate(out["face_hiddens"], ch, "C4 face hiddens [1,16,768]")
    gate(out["audio_hiddens"], ah, "C4 audio hiddens [1,16,1024]")
    want = R.fused_tokens(fus, fh, ch, ah, "multiface_audio_face_frame_text")
    ref = torch.cat([want["multi"], want["audio"], want["face"], want["frame"]], dim=1)
    gate(toks.reshape(4, 3584), ref.reshape(4, 3584), "C4 clip tokens <Multi><Audio><Face><Frame>")
    assert R.nn_identity(out["frame_hiddens"][0].cpu(), fh[0]) and R.nn_identity(out["audio_hiddens"][0].cpu(), ah[0])


def test_hubert_group_norm_at_24_layers_against_oracle():
    """feat_extract_norm='group' (GroupNorm(512, 512) after conv0 only, base-family front-end) under the full 24-layer
    encoder: one clip of 8 segments vs HF HubertModel with that config."""
    import affectgpt_b200 as agb
    hub = R.build_hubert(num_layers=24, feat_extract_norm="group", seed=4321)
    eng = agb.Engine(DEV, vit_layers=2, hubert_layers=24, feat_extract_norm="group")
    enc = agb.HUBERT_LARGE(model_dir="", num_layers=24, feat_extract_norm="group")
    enc.load_hf_state_dict(hub.state_dict())
    enc.attach_engine(eng)
    raw = R.synth_audio(1, 8, seed=81)
    got = enc(None, raw.to(DEV))
    want = R.hubert_large_forward(hub, raw)
    gate(got, want, "hubert24[group] features [1,8,1024]")
    assert R.nn_identity(got[0].cpu(), want[0])
    eng.close()


@pytest.mark.parametrize("scale", [60.0, 100.0])
def test_outlier_channels_in_the_residual_stream(scale):
    """Trained CLIP-L / HuBERT carry 'massive activation' channels in the residual stream (a handful of channels 50-100x
    the others) that random init never produces.  Inject them: a few output-bias entries of an early MLP are set to
    +-scale, so every later block sees a residual row whose variance is dominated by 3-5 channels.  Same gate as
    everywhere else, on 8-layer encoders (the oracle stays in seconds), plus the per-block taps after the injection."""
    import affectgpt_b200 as agb
    clip = R.build_clip_vision(num_layers=8, seed=77)
    hub = R.build_hubert(num_layers=8, feat_extract_norm="layer", seed=77)
    chans = [7, 300, 511, 901, 1023]
    with torch.no_grad():
        for i, c in enumerate(chans):
            sign = 1.0 if i % 2 == 0 else -1.0
            clip.vision_model.encoder.layers[1].mlp.fc2.bias[c] = sign * scale * (0.6 + 0.1 * i)
            hub.encoder.layers[1].feed_forward.output_dense.bias[c] = -sign * scale * (0.6 + 0.1 * i)
    fus = R.build_fusion_state(seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), fus, device=DEV, vit_layers=8, hubert_layers=8)
    eng = model._engine
    raw = R.synth_frames(1, 2, seed=91)
    vt = {n: eng.tap(n, (2 * 257, 1024)) for n in ("vit.block1", "vit.block4")}
    got = model.visual_encoder(None, raw.to(DEV))
    eng.clear_taps()
    px = R.clip_preprocess(raw)
    with torch.no_grad():
        hs = clip.vision_model(pixel_values=px, output_hidden_states=True).hidden_states
    assert float(hs[2].abs().max()) > 0.5 * scale                                  # the outliers are really there
    gate(vt["vit.block1"].reshape(2, 257, 1024), hs[2], f"outlier x{scale:g}: vit.block1")
    gate(vt["vit.block4"].reshape(2, 257, 1024), hs[5], f"outlier x{scale:g}: vit.block4")
    gate(got, R.clip_vit_large_forward(clip, raw), f"outlier x{scale:g}: clip features")
    audio = R.synth_audio(1, 2, seed=92)
    at = {n: eng.tap(n, (2 * 100, 1024)) for n in ("hubert.block1", "hubert.block4")}
    got_a = model.acoustic_encoder(None, audio.to(DEV))
    eng.clear_taps()
    want_a, ahs = R.hubert_large_forward(hub, audio, return_hidden_states=True)
    assert float(ahs[2].abs().max()) > 0.5 * scale
    gate(at["hubert.block1"].reshape(2, 100, 1024)[:, :99], ahs[2], f"outlier x{scale:g}: hubert.block1")
    gate(at["hubert.block4"].reshape(2, 100, 1024)[:, :99], ahs[5], f"outlier x{scale:g}: hubert.block4")
    gate(got_a, want_a, f"outlier x{scale:g}: hubert features")
    eng.close()


def test_fused_layernorm_tail_is_bit_identical(full_path, monkeypatch):
    """AGB_FUSED_LN=1 moves every LayerNorm that follows a residual GEMM into that GEMM's tail (the CTAs of a 256-row block
    meet at a counter and normalise the rows from L2; gemm_tcgen05.cu 'LNF').  Same arithmetic (ln_row_apply), so the tokens
    must not change by a bit - including odd row counts (partial last block) and repeated launches (counters back at zero)."""
    from affectgpt_b200 import engine as E, synthetic as S
    for B in (4, 9):                                                # 9 clips: 72 frames = 18504 rows, 72.3 row blocks
        frames, audio = S.frames(B, 8, seed=81 + B).to(DEV), S.audio(B, 8, seed=91 + B).to(DEV)
        want, _ = _tokens(full_path, frames, audio)
        n0 = E.launch_count()
        want2, _ = _tokens(full_path, frames, audio)
        plain_launches = E.launch_count() - n0
        monkeypatch.setenv("AGB_FUSED_LN", "1")
        for rep in range(3):
            n0 = E.launch_count()
            got, _ = _tokens(full_path, frames, audio)
            fused_launches = E.launch_count() - n0
            assert torch.equal(got, want), (B, rep)
        monkeypatch.delenv("AGB_FUSED_LN")
        assert torch.equal(want2, want)
        assert fused_launches < plain_launches - 80, (fused_launches, plain_launches)    # 2 x (23 + 24) passes are gone
