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
    assert g.kernels_per_replay > 300
    for seed in (61, 62):                                           # two different batches through the same graph
        frames, audio = S.frames(B, 8, seed=seed).to(DEV), S.audio(B, 8, seed=seed + 10).to(DEV)
        out = g(frames, audio)
        want, _ = _tokens(full_path, frames, audio)
        got = torch.cat([out["multi"], out["audio"], out["frame"]], dim=1)
        assert torch.equal(got, want), seed
