"""End-to-end pre-extraction sweep on the GPU (small random-init encoders) against the oracle run sample by sample
the way the reference extractor does (b = 1 per call): same files, same shapes, features inside the bf16 gate."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_path as R
from oracle import sampling as OS
from tests.gpu_util import gate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _write_dataset(root, names):
    from scipy.io import wavfile
    rng = np.random.default_rng(11)
    v, f, a = root / "video", root / "face", root / "audio"
    for d in (v, f, a):
        d.mkdir()
    data = {}
    for i, n in enumerate(names):
        frames = rng.integers(0, 256, size=(20 + 3 * i, 224, 224, 3), dtype=np.uint8)      # pre-decoded RGB stack
        np.save(v / f"{n}.npy", frames)
        faces = rng.integers(0, 256, size=(11 + i, 96, 80, 3), dtype=np.uint8)
        (f / n).mkdir()
        np.save(f / n / f"{n}.npy", faces)
        nsamp = [55357, 30000, 96000][i % 3]
        pcm = (rng.standard_normal((nsamp, 2 if i == 1 else 1)) * 3000 * (1 + i)).astype(np.int16)
        wavfile.write(a / f"{n}.wav", 44100 if i == 4 else 16000, pcm if i == 1 else pcm[:, 0])    # sample 4: resampled on the GPU
        data[n] = (frames, faces, pcm)
    (root / "names.txt").write_text("\n".join(names + ["missing_everything"]) + "\n")
    return data


def test_sweep_matches_oracle_per_sample(tmp_path):
    import cv2
    import affectgpt_b200 as agb
    from affectgpt_b200 import extract as X
    names = ["clipA", "clipB", "clipC"]
    data = _write_dataset(tmp_path, names)
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    fus = R.build_fusion_state(seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), fus, device=DEV, vit_layers=2, hubert_layers=3)
    ex = X.MultimodalFeatureExtractor(model.visual_encoder, model.acoustic_encoder, model, device=DEV)
    args = X.build_parser().parse_args([
        "--dataset", "mer2024", "--video_root", str(tmp_path / "video"), "--face_root", str(tmp_path / "face"),
        "--audio_root", str(tmp_path / "audio"), "--sample_list", str(tmp_path / "names.txt"),
        "--save_root", str(tmp_path / "out"), "--quiet"])
    X.validate_args(args)
    stats = X.run_sweep(args, ex, X.read_sample_names(args), log=lambda *_: None)
    assert stats["frame_success"] == 3 and stats["face_success"] == 3 and stats["audio_success"] == 3
    assert stats["audio_failed"] == 1 and stats["multi_success"] == 3
    root = tmp_path / "out" / "mer2024"
    z = np.load(root / "audio_HUBERT_LARGE_8clips" / "missing_everything.npy")
    assert z.shape == (8, 1024) and not z.any()
    for n in names:
        frames, faces, pcm = data[n]
        # oracle, the way the reference extractor runs it: one sample per call
        idx = OS.uniform_indices(len(frames), 8)
        raw = torch.from_numpy(frames[idx]).permute(3, 0, 1, 2).float()[None]                 # [1,3,8,224,224]
        want_frame = R.clip_vit_large_forward(clip, raw)[0]
        fidx = OS.uniform_indices(len(faces), 8)
        rf = np.stack([cv2.resize(fc, (224, 224)) for fc in faces])[fidx]                     # load_face order: resize, sample
        want_face = R.clip_vit_large_forward(clip, torch.from_numpy(rf).permute(3, 0, 1, 2).float()[None])[0]
        wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).T.contiguous()               # [ch, n]
        segs, _ = OS.audio_segments(wav)
        want_audio = R.hubert_large_forward(hub, segs[None])[0]
        want_multi = R.encode_multi_attention(fus, want_face[None], want_audio[None])[0][0]
        for sub, want in (("frame_CLIP_VIT_LARGE_uniform_8frms", want_frame), ("face_CLIP_VIT_LARGE_8frms", want_face),
                          ("audio_HUBERT_LARGE_8clips", want_audio)):
            got = np.load(root / sub / f"{n}.npy")
            assert got.dtype == np.float32 and got.shape == tuple(want.shape)
            gate(torch.from_numpy(got), want, f"{sub}/{n}")
        gm = np.load(root / "multi_CLIP_VIT_LARGE_HUBERT_LARGE_complete" / f"{n}.npy")
        assert gm.shape == (1024,)
        gate(torch.from_numpy(gm)[None], want_multi[None], f"multi/{n}")


def test_batched_sweep_files_equal_per_sample_sweep_bit_for_bit(tmp_path):
    """--batch_size N shares one ViT call and one HuBERT call between N samples; the encoders are batch-invariant bit for
    bit, so every .npy must equal the file the reference-style per-sample loop writes."""
    import affectgpt_b200 as agb
    from affectgpt_b200 import extract as X
    names = ["clipA", "clipB", "clipC", "clipD", "clipE"]
    _write_dataset(tmp_path, names)
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), R.build_fusion_state(seed=1234), device=DEV, vit_layers=2,
                           hubert_layers=3)
    ex = X.MultimodalFeatureExtractor(model.visual_encoder, model.acoustic_encoder, model, device=DEV)
    trees, stats = [], []
    for mode, bs in (("single", 1), ("batched", 3)):
        args = X.build_parser().parse_args([
            "--dataset", "mer2024", "--video_root", str(tmp_path / "video"), "--face_root", str(tmp_path / "face"),
            "--audio_root", str(tmp_path / "audio"), "--sample_list", str(tmp_path / "names.txt"),
            "--save_root", str(tmp_path / mode), "--quiet", "--batch_size", str(bs)])
        X.validate_args(args)
        nm = X.read_sample_names(args)
        st = X.run_sweep(args, ex, nm, log=lambda *_: None) if bs == 1 else \
            X.run_sweep_batched(args, ex, nm, batch_size=bs, workers=2, log=lambda *_: None)
        stats.append(st)
        tree = {}
        for d, _, files in os.walk(tmp_path / mode):
            for f in files:
                tree[os.path.relpath(os.path.join(d, f), tmp_path / mode)] = np.load(os.path.join(d, f))
        trees.append(tree)
    assert stats[0] == stats[1] and stats[1]["multi_success"] == 5 and stats[1]["audio_failed"] == 1
    assert sorted(trees[0]) == sorted(trees[1]) and len(trees[0]) == 4 * 5 + 1
    for k in trees[0]:
        assert np.array_equal(trees[0][k], trees[1][k]), k


def test_emotion_peak_sweep_follows_the_extractor_not_load_video(tmp_path):
    """--frame_sampling emotion_peak: the reference extractor branches to extract_frame_features_smart
    (extract_multimodal_features_precompute.py:174-175): AU json missing -> np.linspace(0, vlen - 1, 8) (:400-403), AU json
    with a peak -> calculate_smart_frame_indices; never load_video's "middle frame repeated 8 times".  The file on disk must
    be the features of exactly those frames."""
    import json
    import affectgpt_b200 as agb
    from affectgpt_b200 import extract as X, sampling
    names = ["clipA", "clipB"]
    data = _write_dataset(tmp_path, names)
    mer = tmp_path / "mer_factory"
    (mer / "clipB").mkdir(parents=True)
    vlen_b = len(data["clipB"][0])
    au = {"total_frames": vlen_b, "peak_frames": [{"peak_index": 9, "frames_before_peak": 3, "frames_after_peak": 3}]}
    (mer / "clipB" / "clipB_au_analysis.json").write_text(json.dumps({"au_info": au}))
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
    model = agb.build_path(clip.state_dict(), hub.state_dict(), R.build_fusion_state(seed=1234), device=DEV, vit_layers=2,
                           hubert_layers=3)
    ex = X.MultimodalFeatureExtractor(model.visual_encoder, model.acoustic_encoder, model, device=DEV,
                                      mer_factory_output=str(mer))
    args = X.build_parser().parse_args([
        "--dataset", "mer2024", "--modality", "frame", "--video_root", str(tmp_path / "video"), "--sample_list",
        str(tmp_path / "names.txt"), "--save_root", str(tmp_path / "out"), "--quiet", "--frame_sampling", "emotion_peak",
        "--mer-factory-output", str(mer)])
    X.validate_args(args)
    X.run_sweep(args, ex, names, log=lambda *_: None)
    out = tmp_path / "out" / "mer2024" / "frame_CLIP_VIT_LARGE_emotion_peak_8frms"
    want_idx = {"clipA": sorted(np.linspace(0, len(data["clipA"][0]) - 1, 8).astype(int).tolist()),      # no AU json
                "clipB": OS.smart_frame_indices(au, vlen_b)}
    assert want_idx["clipB"] == sampling.smart_frame_indices(au, vlen_b) and len(set(want_idx["clipA"])) == 8
    for n in names:
        frames = data[n][0]
        raw = torch.from_numpy(frames[want_idx[n]]).permute(3, 0, 1, 2).float()[None]
        want = R.clip_vit_large_forward(clip, raw)[0]
        got = np.load(out / f"{n}.npy")
        assert got.shape == (8, 768)
        gate(torch.from_numpy(got), want, f"emotion_peak frame features {n}")
        mid = torch.from_numpy(frames[[len(frames) // 2] * 8]).permute(3, 0, 1, 2).float()[None]
        wrong = R.clip_vit_large_forward(clip, mid)[0]                    # what load_video's rule would have produced
        assert float((torch.from_numpy(got) - wrong).abs().max()) > 10 * float((torch.from_numpy(got) - want).abs().max())
