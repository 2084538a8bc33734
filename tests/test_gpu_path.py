"""End-to-end parity of the CUDA path (through the reference-shaped wrappers and the C ABI) against the oracle
on identical seeded weights and inputs.  Gate (BASELINE.json): cosine >= 0.999 and max|x-y|/max|y| <= 2e-2 vs the
fp32 oracle, plus the discriminative checks of SURVEY 8d (per-block taps, nearest-neighbour identity)."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_path as R
from tests.gpu_util import gate, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def small_path():
    import affectgpt_b200 as agb
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    fus = R.build_fusion_state(seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), fus, device=DEV, vit_layers=2, hubert_layers=3)
    return model, clip, hub, fus


def test_vit_small_with_per_block_taps(small_path):
    model, clip, _, _ = small_path
    raw = R.synth_frames(2, 3, seed=1)                                              # 6 frames
    eng = model._engine
    taps = {n: eng.tap(n, (6 * 257, 1024)) for n in ("vit.embed", "vit.block0", "vit.block1")}
    got = model.visual_encoder(None, raw.to(DEV))
    eng.clear_taps()
    px = R.clip_preprocess(raw)
    with torch.no_grad():
        out = clip.vision_model(pixel_values=px, output_hidden_states=True)
        want = clip(pixel_values=px).image_embeds.reshape(2, 3, 768)
    hs = out.hidden_states                                                          # [embed+pre_LN, block0, block1]
    for name, ref in zip(("vit.embed", "vit.block0", "vit.block1"), hs):
        gate(taps[name].reshape(6, 257, 1024), ref, name)
    gate(got, want, "clip2 features")
    assert got.dtype == torch.float32 and got.device.type == "cuda" and tuple(got.shape) == (2, 3, 768)
    assert R.nn_identity(got.cpu(), want)


def test_vit_pruned_last_block_matches_dense_block(small_path):
    """Default: the last ViT block runs K/V for all tokens but Q / attention / out-proj / MLP for the pooled CLS row only
    (CLIPVisionTransformer pools hidden[:, 0], modeling_clip.py:685-686).  It must agree with the dense block, and both
    with the oracle; a tap on the last block switches that call to the dense block (all 257 rows valid)."""
    import affectgpt_b200 as agb
    model, clip, _, _ = small_path
    raw = R.synth_frames(2, 4, seed=5)
    frames = raw.permute(0, 2, 1, 3, 4).reshape(8, 3, 224, 224).to(DEV)
    pruned = model._engine.clip_forward(frames)
    dense_eng = agb.Engine(DEV, vit_layers=2, hubert_layers=3, vit_dense_last_block=True)
    dense_eng.load_state_dict("clip", clip.state_dict())
    dense = dense_eng.clip_forward(frames)
    dense_eng.close()
    gate(pruned, dense, "pruned vs dense last block", cos_min=0.99999, rel_max=5e-3)   # fp32 vs bf16 probabilities
    with torch.no_grad():
        want = clip(pixel_values=R.clip_preprocess(raw)).image_embeds
    gate(pruned, want, "pruned vs oracle")
    gate(dense, want, "dense vs oracle")
    tap = model._engine.tap("vit.block1", (8 * 257, 1024))
    tapped = model._engine.clip_forward(frames)
    model._engine.clear_taps()
    assert torch.equal(tapped, dense)                                               # tap => dense block, bit for bit
    assert float(tap.reshape(8, 257, 1024)[:, 1:].abs().max()) > 0
    assert torch.equal(model._engine.clip_forward(frames), pruned)                  # deterministic, back to pruned


def test_vit_u8_input_and_chunking_are_bit_identical(small_path):
    import affectgpt_b200 as agb
    model, clip, _, _ = small_path
    raw = R.synth_frames(1, 5, seed=2)
    frames = raw.permute(0, 2, 1, 3, 4).reshape(5, 3, 224, 224).to(DEV)
    eng = model._engine
    a = eng.clip_forward(frames)
    b = eng.clip_forward(frames.to(torch.uint8))
    assert torch.equal(a, b)
    assert torch.equal(a, eng.clip_forward(frames))                                 # deterministic
    eng2 = agb.Engine(DEV, vit_layers=2, hubert_layers=3, max_frames_per_chunk=2)
    eng2.load_state_dict("clip", clip.state_dict())
    assert torch.equal(eng2.clip_forward(frames), a)                                # chunks of 2,2,1 frames
    eng2.close()


@pytest.mark.parametrize("norm", ["layer", "group"])
def test_hubert_small_with_taps_and_golden(norm):
    import affectgpt_b200 as agb
    hub = R.build_hubert(num_layers=3, feat_extract_norm=norm, seed=1234)
    eng = agb.Engine(DEV, vit_layers=2, hubert_layers=3, feat_extract_norm=norm)
    enc = agb.HUBERT_LARGE(model_dir="", num_layers=3, feat_extract_norm=norm)
    enc.load_hf_state_dict(hub.state_dict())
    enc.attach_engine(eng)
    raw = R.synth_audio(1, 2, seed=3)
    taps = {n: eng.tap(n, (2 * 100, 1024)) for n in ("hubert.featproj", "hubert.posconv", "hubert.block0", "hubert.block2")}
    conv_tap = eng.tap("hubert.conv", (2 * 100, 512), dtype=torch.bfloat16)
    got = enc(None, raw.to(DEV))
    eng.clear_taps()
    want, hs = R.hubert_large_forward(hub, raw, return_hidden_states=True)
    with torch.no_grad():
        x = torch.from_numpy(R.wav2vec2_normalize(raw[:, :, 0, :].reshape(2, 32000).numpy()))
        conv = hub.feature_extractor(x).transpose(1, 2)                             # [2,99,512]
        proj, _ = hub.feature_projection(conv) if isinstance(hub.feature_projection(conv), tuple) else (hub.feature_projection(conv), None)
    gate(conv_tap.float().reshape(2, 100, 512)[:, :99], conv, f"hubert.conv[{norm}]", rel_max=3e-2)
    gate(taps["hubert.featproj"].reshape(2, 100, 1024)[:, :99], proj, "hubert.featproj")
    gate(taps["hubert.posconv"].reshape(2, 100, 1024)[:, :99], hs[0], "hubert.posconv")
    gate(taps["hubert.block0"].reshape(2, 100, 1024)[:, :99], hs[1], "hubert.block0")
    gate(got, want, f"hubert3[{norm}] features")
    gold = np.load(os.path.join(GOLD, "golden_small.npz"))[f"hubert3_{norm}_audio_seed3_b1_t2"]
    gate(got, torch.from_numpy(gold), f"hubert3[{norm}] vs golden")
    eng.close()


@pytest.mark.parametrize("norm,conv_bias,fp_ln", [("layer", True, True), ("group", False, False), ("group", True, True)])
def test_hubert_post_layer_norm_encoder(norm, conv_bias, fp_ln):
    """do_stable_layer_norm = false: HF HubertEncoder / HubertEncoderLayer (post-LN blocks, encoder.layer_norm in front of
    the stack, no LayerNorm after it) - what a base-family config.json selects, together with feat_extract_norm='group',
    conv_bias=false and feat_proj_layer_norm=false.  Same gate, per-block taps (hidden_states[l+1] = output of layer l)."""
    import affectgpt_b200 as agb
    hub = R.build_hubert(num_layers=4, feat_extract_norm=norm, seed=2468, do_stable_layer_norm=False, conv_bias=conv_bias,
                         feat_proj_layer_norm=fp_ln)
    sd = hub.state_dict()
    assert ("feature_extractor.conv_layers.1.conv.bias" in sd) == conv_bias
    assert ("feature_projection.layer_norm.weight" in sd) == fp_ln
    eng = agb.Engine(DEV, vit_layers=2, hubert_layers=4, feat_extract_norm=norm, do_stable_layer_norm=False,
                     feat_proj_layer_norm=fp_ln)
    enc = agb.HUBERT_LARGE(model_dir="", num_layers=4, feat_extract_norm=norm, do_stable_layer_norm=False,
                           feat_proj_layer_norm=fp_ln)
    enc.load_hf_state_dict(sd)
    enc.attach_engine(eng)
    raw = R.synth_audio(1, 3, seed=13)
    taps = {n: eng.tap(n, (3 * 100, 1024)) for n in ("hubert.block0", "hubert.block3")}
    got = enc(None, raw.to(DEV))
    eng.clear_taps()
    want, hs = R.hubert_large_forward(hub, raw, return_hidden_states=True)
    assert len(hs) == 5
    gate(taps["hubert.block0"].reshape(3, 100, 1024)[:, :99], hs[1], f"post-LN[{norm}] hubert.block0")
    gate(taps["hubert.block3"].reshape(3, 100, 1024)[:, :99], hs[4], f"post-LN[{norm}] hubert.block3")
    gate(got, want, f"post-LN[{norm}, conv_bias={conv_bias}, feat_proj_ln={fp_ln}] features")
    assert R.nn_identity(got[0].cpu(), want[0])
    eng.close()


def test_hubert_norm_scope_batch_vs_clip(small_path):
    """b=2: per-clip statistics (reference at b=1 per clip) vs joint statistics (reference fed b=2) -- SURVEY F3."""
    import affectgpt_b200 as agb
    model, _, hub, _ = small_path
    raw = R.synth_audio(2, 2, seed=11)
    raw[1] *= 4.0                                                                   # clips with different loudness
    enc = model.acoustic_encoder
    per_clip = enc(None, raw.to(DEV))
    want_clip = torch.cat([R.hubert_large_forward(hub, raw[i:i + 1]) for i in range(2)], 0)
    gate(per_clip, want_clip, "norm_scope=clip vs oracle at b=1")
    enc_b = agb.HUBERT_LARGE(model_dir="", num_layers=3, norm_scope="batch")
    enc_b.attach_engine(model._engine)
    joint = enc_b(None, raw.to(DEV))
    gate(joint, R.hubert_large_forward(hub, raw), "norm_scope=batch vs oracle at b=2")
    assert rel_err(per_clip, joint) > 1e-3                                          # the two scopes really differ


def test_fusion_heads_and_projectors(small_path):
    model, _, _, fus = small_path
    g = torch.Generator().manual_seed(7)
    fh, ch, ah = torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 1024, generator=g)
    want = R.fused_tokens(fus, fh, ch, ah, "multiface_audio_face_frame_text")
    gold = np.load(os.path.join(GOLD, "golden_small.npz"))
    h, frame = model.encode_video_merge(fh.to(DEV), None, is_preextracted=True)
    _, face = model.encode_video_merge(ch.to(DEV), None, is_preextracted=True)
    _, audio = model.encode_audio_merge(ah.to(DEV), None, is_preextracted=True)
    multi_h, multi = model.encode_multi_merge(ch.to(DEV), ah.to(DEV))
    for name, got in (("frame", frame), ("face", face), ("audio", audio), ("multi", multi)):
        assert tuple(got.shape) == (2, 1, 3584)
        assert rel_err(got, want[name]) < 1e-4, name                                # fp32 heads: tight
        assert rel_err(got, torch.from_numpy(gold[f"fusion_seed7_{name}"])) < 1e-4
    want_multi_h, _ = R.encode_multi_attention(fus, ch, ah)
    assert rel_err(multi_h, want_multi_h) < 1e-4 and tuple(multi_h.shape) == (2, 1024)
    # pre-extracted single vector [b,768]: attention bypassed (affectgpt.py:587-589)
    _, one = model.encode_video_merge(fh[:, 0].to(DEV), None, is_preextracted=True)
    assert rel_err(one, R.encode_video_attention(fus, fh[:, 0])[1]) < 1e-4
    # mean fusion variant
    import affectgpt_b200 as agb
    m2 = agb.AffectGPTFusion(None, None, video_fusion_type="mean", audio_fusion_type="mean", engine=model._engine)
    _, vm = m2.encode_video_merge(fh.to(DEV), None, is_preextracted=True)
    assert rel_err(vm, R.encode_video_attention(fus, fh, fusion_type="mean")[1]) < 1e-4
    _, am = m2.encode_audio_merge(ah.to(DEV), None, is_preextracted=True)
    assert rel_err(am, R.encode_audio_attention(fus, ah, fusion_type="mean")[1]) < 1e-4


def test_two_stream_encoder_overlap_is_bit_identical(small_path):
    """encode_clip_batch issues the audio chain on a side stream (joined before the heads): same kernels, same inputs,
    so every output equals the single-stream run bit for bit - eagerly and when replayed from a captured graph."""
    import affectgpt_b200 as agb
    model = small_path[0]
    s = {"raw_frames": R.synth_frames(2, 3, seed=7).to(DEV), "raw_faces": R.synth_frames(2, 3, seed=8).to(DEV),
         "raw_audios": R.synth_audio(2, 2, seed=9).to(DEV)}
    was = model.overlap_encoders
    model.overlap_encoders = True
    try:
        a = {k: v.clone() for k, v in model.encode_clip_batch(s).items()}
        g = agb.GraphedClipBatch(model, 2, 3, 2, with_faces=True, device=DEV)
        c = {k: v.clone() for k, v in g(s["raw_frames"], s["raw_audios"], s["raw_faces"]).items()}
        model.overlap_encoders = False
        b = model.encode_clip_batch(s)
        torch.cuda.synchronize()
        assert set(a) == set(b) == set(c)
        for k in a:
            assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k
    finally:
        model.overlap_encoders = was


def test_encode_clip_batch_token_set(small_path):
    model, clip, hub, fus = small_path
    frames, faces, audio = R.synth_frames(1, 2, seed=1), R.synth_frames(1, 2, seed=2), R.synth_audio(1, 2, seed=3)
    out = model.encode_clip_batch({"raw_frames": frames.to(DEV), "raw_faces": faces.to(DEV), "raw_audios": audio.to(DEV),
                                   "face_or_frame": "multiface_audio_face_frame_text"})
    fh, ch = R.clip_vit_large_forward(clip, frames), R.clip_vit_large_forward(clip, faces)
    ah = R.hubert_large_forward(hub, audio)
    want = R.fused_tokens(fus, fh, ch, ah, "multiface_audio_face_frame_text")
    for k in ("frame", "face", "audio", "multi"):
        gate(out[k], want[k], f"token[{k}]")


def test_full_config_c1_against_golden():
    """BASELINE config C1: one clip, 8 frames + 8 face crops + 8 x 2 s audio, 24-layer encoders."""
    import affectgpt_b200 as agb
    gold = np.load(os.path.join(GOLD, "golden_c1.npz"))
    clip = R.build_clip_vision(num_layers=24, seed=1234)
    hub = R.build_hubert(num_layers=24, feat_extract_norm="layer", seed=1234)
    fus = R.build_fusion_state(seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), fus, device=DEV)
    del clip, hub
    out = model.encode_clip_batch({"raw_frames": R.synth_frames(1, 8, seed=1).to(DEV),
                                   "raw_faces": R.synth_frames(1, 8, seed=2).to(DEV),
                                   "raw_audios": R.synth_audio(1, 8, seed=3).to(DEV),
                                   "face_or_frame": "multiface_audio_face_frame_text"})
    gate(out["frame_hiddens"], torch.from_numpy(gold["c1_frame_hidden"]), "c1 frame hiddens")
    gate(out["face_hiddens"], torch.from_numpy(gold["c1_face_hidden"]), "c1 face hiddens")
    gate(out["audio_hiddens"], torch.from_numpy(gold["c1_audio_hidden"]), "c1 audio hiddens")
    for k in ("frame", "face", "audio", "multi"):
        gate(out[k], torch.from_numpy(gold[f"c1_token_{k}"]), f"c1 token[{k}]")
    # discriminative: every frame's feature is closest to its own oracle feature after mean-centring
    assert R.nn_identity(out["frame_hiddens"][0].cpu(), torch.from_numpy(gold["c1_frame_hidden"][0]))
    assert R.nn_identity(out["audio_hiddens"][0].cpu(), torch.from_numpy(gold["c1_audio_hidden"][0]))
    model._engine.close()


def test_demo_clip_structured_input_against_golden():
    """a real clip instead of noise - the reference's demo mp4 / wav (decoded frames and PCM travel in the
    fixture, tests/golden/make_demo_golden.py) through staging (GPU audio slicing), both 24-layer encoders and the heads,
    against the oracle's features for the same inputs."""
    import affectgpt_b200 as agb
    from affectgpt_b200 import staging
    gold = np.load(os.path.join(GOLD, "golden_demo_clip.npz"))
    clip = R.build_clip_vision(num_layers=24, seed=1234)
    hub = R.build_hubert(num_layers=24, feat_extract_norm="layer", seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), R.build_fusion_state(seed=1234), device=DEV)
    del clip, hub
    raw = torch.from_numpy(gold["frames_u8"]).permute(3, 0, 1, 2).float()[None].to(DEV)          # [1,3,8,224,224] 0..255
    wav = torch.from_numpy(gold["pcm_i16"].astype(np.float32) / 32768.0)[None]                   # [1, 55357]
    segs = staging.stage_audio(wav, device=DEV)                                                   # [8,1,32000] on the GPU
    out = model.encode_clip_batch({"raw_frames": raw, "raw_audios": segs[None], "face_or_frame": "multiframe_audio_frame_text"})
    gate(out["frame_hiddens"], torch.from_numpy(gold["frame_hidden"]), "demo frame hiddens")
    gate(out["audio_hiddens"], torch.from_numpy(gold["audio_hidden"]), "demo audio hiddens")
    for k in ("frame", "audio", "multi"):
        gate(out[k], torch.from_numpy(gold[f"token_{k}"]), f"demo token[{k}]")
    # the uint8 entry point of the engine (what the sweep CLI uses) sees the same frames
    u8 = torch.from_numpy(gold["frames_u8"]).permute(0, 3, 1, 2).contiguous().to(DEV)
    assert torch.equal(model._engine.clip_forward(u8), out["frame_hiddens"][0])
    assert R.nn_identity(out["frame_hiddens"][0].cpu(), torch.from_numpy(gold["frame_hidden"][0]))
    assert R.nn_identity(out["audio_hiddens"][0].cpu(), torch.from_numpy(gold["audio_hidden"][0]))
    model._engine.close()


def test_batch_of_clips_equals_clip_by_clip(small_path):
    """Clips are independent (the data-parallel premise): a batch equals its clips run one at a time."""
    model, _, _, _ = small_path
    frames = R.synth_frames(3, 2, seed=21).to(DEV)
    audio = R.synth_audio(3, 2, seed=22).to(DEV)
    fb = model.visual_encoder(None, frames)
    ab = model.acoustic_encoder(None, audio)
    for i in range(3):
        assert torch.equal(model.visual_encoder(None, frames[i:i + 1]), fb[i:i + 1])
        assert torch.equal(model.acoustic_encoder(None, audio[i:i + 1]), ab[i:i + 1])


def test_missing_weights_raise():
    import affectgpt_b200 as agb
    eng = agb.Engine(DEV, vit_layers=2, hubert_layers=3)
    with pytest.raises(RuntimeError, match="weights not loaded"):
        eng.clip_forward(torch.zeros(1, 3, 224, 224, device=DEV))
    with pytest.raises(ValueError):
        eng.clip_forward(torch.zeros(1, 3, 100, 224, device=DEV))
    eng.close()
