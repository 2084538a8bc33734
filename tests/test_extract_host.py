"""Sweep CLI host logic (no GPU): directory layout, resume, zero-fill and argument rules of the reference's
extract_multimodal_features_precompute.py (:823-849, 892-1024, 1098-1123), with a stub extractor."""
import os

import numpy as np
import pytest

from affectgpt_b200 import extract as X


class StubExtractor:
    def __init__(self):
        self.calls = []

    def extract_frame_features(self, video_path, n_frms=8, sampling="uniform", video_name=None):
        self.calls.append(("frame", video_name, n_frms, sampling))
        return np.full((n_frms, 768), 1.0, np.float32)

    def extract_face_features(self, face_npy_path, n_frms=8):
        self.calls.append(("face", os.path.basename(face_npy_path), n_frms))
        return np.full((n_frms, 768), 2.0, np.float32)

    def extract_audio_features(self, audio_path, clips_per_video=8):
        self.calls.append(("audio", os.path.basename(audio_path)))
        return None if "broken" in audio_path else np.full((clips_per_video, 1024), 3.0, np.float32)

    def extract_multi_features(self, face_features, audio_features):
        return np.full((1024,), 4.0, np.float32)


def _dataset(tmp_path, names):
    v, f, a = tmp_path / "video", tmp_path / "face", tmp_path / "audio"
    for d in (v, f, a):
        d.mkdir()
    for n in names:
        (v / f"{n}.mp4").write_bytes(b"x")
        (f / n).mkdir()
        np.save(f / n / f"{n}.npy", np.zeros((3, 4, 4, 3), np.uint8))
        if n != "noaudio":
            (a / f"{n}.wav").write_bytes(b"x")
    lst = tmp_path / "names.txt"
    lst.write_text("\n".join(names) + "\n")
    return ["--dataset", "mer2023", "--video_root", str(v), "--face_root", str(f), "--audio_root", str(a),
            "--sample_list", str(lst), "--save_root", str(tmp_path / "out"), "--quiet"]


def test_layout_zero_fill_and_resume(tmp_path):
    argv = _dataset(tmp_path, ["s1", "broken", "noaudio"]) + ["--frame_sampling", "emotion_peak", "--frame_n_frms", "1",
                                                                "--mer-factory-output", str(tmp_path)]
    args = X.build_parser().parse_args(argv)
    X.validate_args(args)
    names = X.read_sample_names(args)
    assert names == ["s1", "broken", "noaudio"]
    ex = StubExtractor()
    stats = X.run_sweep(args, ex, names, log=lambda *_: None)
    root = tmp_path / "out" / "mer2023"
    assert sorted(os.listdir(root)) == ["audio_HUBERT_LARGE_8clips", "face_CLIP_VIT_LARGE_8frms",
                                        "frame_CLIP_VIT_LARGE_emotion_peak_1frms", "multi_CLIP_VIT_LARGE_HUBERT_LARGE_complete"]
    assert np.load(root / "frame_CLIP_VIT_LARGE_emotion_peak_1frms" / "s1.npy").shape == (1, 768)
    assert np.load(root / "face_CLIP_VIT_LARGE_8frms" / "s1.npy").shape == (8, 768)
    for n in ("broken", "noaudio"):                          # failed processing / missing wav: zero-filled [clips, 1024]
        z = np.load(root / "audio_HUBERT_LARGE_8clips" / f"{n}.npy")
        assert z.shape == (8, 1024) and z.dtype == np.float32 and not z.any()
    assert np.load(root / "audio_HUBERT_LARGE_8clips" / "s1.npy").mean() == 3.0
    assert np.load(root / "multi_CLIP_VIT_LARGE_HUBERT_LARGE_complete" / "s1.npy").shape == (1024,)
    assert not (root / "multi_CLIP_VIT_LARGE_HUBERT_LARGE_complete" / "broken.npy").exists()
    assert stats["audio_success"] == 1 and stats["audio_failed"] == 2 and stats["multi_success"] == 1
    assert ("frame", "s1", 1, "emotion_peak") in ex.calls
    ex2 = StubExtractor()                                    # second run: everything that exists is skipped
    stats2 = X.run_sweep(args, ex2, names, log=lambda *_: None)
    assert stats2["frame_success"] == stats2["face_success"] == stats2["audio_success"] == 0
    assert [c for c in ex2.calls if c[0] == "frame"] == []


def test_skip_multi_and_modality_subsets(tmp_path):
    argv = _dataset(tmp_path, ["a"]) + ["--modality", "audio", "--clips_per_video", "4"]
    args = X.build_parser().parse_args(argv)
    X.validate_args(args)
    X.run_sweep(args, StubExtractor(), ["a"], log=lambda *_: None)
    root = tmp_path / "out" / "mer2023"
    assert os.listdir(root) == ["audio_HUBERT_LARGE_4clips"]
    args2 = X.build_parser().parse_args(_dataset_args_only(argv) + ["--skip-multi-preextract"])
    assert "multi" not in X.save_dirs(args2)


def _dataset_args_only(argv):
    out = list(argv)
    i = out.index("--modality")
    del out[i:i + 2]
    return out


def test_argument_rules_match_reference():
    p = X.build_parser()
    with pytest.raises(ValueError, match="video_root"):
        X.validate_args(p.parse_args(["--dataset", "meld", "--modality", "frame", "--sample_list", "x"]))
    with pytest.raises(ValueError, match="face_root"):
        X.validate_args(p.parse_args(["--dataset", "meld", "--modality", "multi", "--audio_root", "a", "--sample_list", "x"]))
    with pytest.raises(ValueError, match="audio_root"):
        X.validate_args(p.parse_args(["--dataset", "meld", "--modality", "audio", "--sample_list", "x"]))
    with pytest.raises(ValueError, match="Either"):
        X.read_sample_names(p.parse_args(["--dataset", "meld", "--modality", "audio", "--audio_root", "a"]))
    with pytest.raises(ValueError, match="both"):
        X.read_sample_names(p.parse_args(["--dataset", "meld", "--sample_list", "x", "--csv_path", "y"]))
    a = p.parse_args(["--dataset", "meld", "--modality", "audio", "--audio_root", "a", "--sample_list", "x", "--n_frms", "4"])
    X.validate_args(a)
    assert a.frame_n_frms == 4                               # deprecated --n_frms still honoured (:1091-1095)
    with pytest.raises(SystemExit):
        p.parse_args(["--dataset", "not-a-dataset"])


def test_csv_sample_source_and_frame_index_rules(tmp_path):
    csv = tmp_path / "s.csv"
    csv.write_text("names,label\nclip_a,1\nclip_b,0\n")
    args = X.build_parser().parse_args(["--dataset", "meld", "--csv_path", str(csv)])
    assert X.read_sample_names(args) == ["clip_a", "clip_b"]
    assert X.frame_indices(88, 8, "uniform") == [0, 11, 22, 33, 44, 55, 66, 77]
    assert X.frame_indices(88, 8, "emotion_peak") == [44] * 8          # no video name: load_video's rule (:229-237)
    # with a video name the extractor takes extract_frame_features_smart (extract_...precompute.py:174-175): a missing
    # *_au_analysis.json means np.linspace(0, vlen - 1, 8), never the repeated middle frame (:400-403)
    import numpy as np
    want = sorted(np.linspace(0, 87, 8).astype(int).tolist())
    assert X.frame_indices(88, 8, "emotion_peak", "clip_a", str(tmp_path)) == want
    assert X.frame_indices(88, 8, "emotion_peak", "clip_a", None) == want
    d = tmp_path / "clip_b"
    d.mkdir()
    (d / "clip_b_au_analysis.json").write_text('{"au_info": {"total_frames": 88, "peak_frames": []}}')   # no peaks: same fallback
    assert X.frame_indices(88, 8, "emotion_peak", "clip_b", str(tmp_path)) == want
    (d / "clip_b_au_analysis.json").write_text(
        '{"au_info": {"total_frames": 88, "peak_frames": [{"peak_index": 40, "frames_before_peak": 3, "frames_after_peak": 3}]}}')
    got = X.frame_indices(88, 8, "emotion_peak", "clip_b", str(tmp_path))
    assert len(got) == 8 and {38, 39, 40, 41, 42} <= set(got) and got == sorted(got)
    assert X.frame_indices(5, 8, "uniform") == [0, 1, 2, 3, 4, 4, 4, 4]


class BatchStub(StubExtractor):
    """The batched surface of MultimodalFeatureExtractor over the same fake features as StubExtractor."""

    def decode_frames(self, video_path, n_frms=8, sampling="uniform", video_name=None):
        self.calls.append(("frame", video_name, n_frms, sampling))
        return np.zeros((n_frms, 2, 2, 3), np.uint8)

    def decode_faces(self, face_npy_path, n_frms=8):
        self.calls.append(("face", os.path.basename(face_npy_path), n_frms))
        return np.zeros((n_frms, 2, 2, 3), np.uint8)

    def decode_audio(self, audio_path):
        self.calls.append(("audio", os.path.basename(audio_path)))
        return None if "broken" in audio_path else np.zeros((1, 4), np.float32)

    def prestage(self, decoded, clips_per_video=8):
        return {"decoded": decoded, "clips": clips_per_video}

    def submit_batch(self, staged, multi_of):
        d = staged["decoded"]
        self.calls.append(("visual_batch", sum(x.get("frame") is not None for x in d), sum(x.get("face") is not None for x in d)))
        return {"staged": staged, "multi_of": multi_of}

    def collect_batch(self, h):
        d, c = h["staged"]["decoded"], h["staged"]["clips"]
        ff = [None if x.get("frame") is None else np.full((len(x["frame"]), 768), 1.0, np.float32) for x in d]
        cf = [None if x.get("face") is None else np.full((len(x["face"]), 768), 2.0, np.float32) for x in d]
        af = [None if x.get("audio") is None else np.full((c, 1024), 3.0, np.float32) for x in d]
        mf = [np.full((1024,), 4.0, np.float32) if (i in h["multi_of"] and cf[i] is not None and af[i] is not None) else None
              for i in range(len(d))]
        return ff, cf, af, mf


def _tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            out[os.path.relpath(os.path.join(d, f), root)] = np.load(os.path.join(d, f))
    return out


@pytest.mark.parametrize("modality", ["all", "multi", "face", "audio"])
def test_batched_sweep_writes_what_the_per_sample_sweep_writes(tmp_path, modality):
    """run_sweep_batched (device batches + decode-ahead thread pool) == run_sweep (the reference's loop): same files,
    same contents, same statistics - including zero-filled audio, missing sources, and a resumed second pass."""
    names = ["s1", "broken", "noaudio", "s4", "s5"]
    argv = _dataset(tmp_path, names) + ["--modality", modality]
    (tmp_path / "face" / "s4" / "s4.npy").unlink()                       # a sample without face crops
    results = []
    for mode in ("single", "batched"):
        a = X.build_parser().parse_args(argv[:argv.index("--save_root") + 1] + [str(tmp_path / mode)] +
                                        argv[argv.index("--save_root") + 2:])
        X.validate_args(a)
        ex = BatchStub()
        if mode == "single":
            st = X.run_sweep(a, ex, names + ["ghost"], log=lambda *_: None)
        else:
            st = X.run_sweep_batched(a, ex, names + ["ghost"], batch_size=2, workers=3, log=lambda *_: None)
            if modality == "all":                                        # frames and faces of a batch share one encoder call
                assert ("visual_batch", 2, 2) in ex.calls
        results.append((st, _tree(tmp_path / mode)))
    (st_a, tree_a), (st_b, tree_b) = results
    assert st_a == st_b
    assert sorted(tree_a) == sorted(tree_b) and len(tree_a) > 0
    for k in tree_a:
        assert tree_a[k].dtype == tree_b[k].dtype and np.array_equal(tree_a[k], tree_b[k]), k
    a = X.build_parser().parse_args(argv[:argv.index("--save_root") + 1] + [str(tmp_path / "batched")] +
                                    argv[argv.index("--save_root") + 2:])
    st2 = X.run_sweep_batched(a, BatchStub(), names + ["ghost"], batch_size=4, log=lambda *_: None)   # resume
    assert st2["frame_success"] == st2["face_success"] == st2["audio_success"] == st2["multi_success"] == 0
    assert st2["skipped_existing"] == len(tree_b)


@pytest.mark.parametrize("bs", [1, 2])
def test_batched_sweep_falls_back_per_sample_when_the_device_batch_raises(tmp_path, bs):
    class Flaky(BatchStub):
        def submit_batch(self, staged, multi_of):
            raise RuntimeError("device batch failed")
    argv = _dataset(tmp_path, ["a", "b"])
    a = X.build_parser().parse_args(argv)
    X.validate_args(a)
    msgs = []
    st = X.run_sweep_batched(a, Flaky(), ["a", "b"], batch_size=bs, log=msgs.append)
    assert st["frame_success"] == 2 and st["face_success"] == 2 and st["multi_success"] == 2
    assert any("sample by sample" in m for m in msgs)


def _claim_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    claim = X.shared_batch_counter(rank, world)
    got = []
    for _ in range(50):
        b = claim()
        if b >= 12:
            break
        got.append(b)
        if rank == 1:
            import time
            time.sleep(0.005)                                  # the "slow GPU": must end up with fewer batches
    q.put((rank, got))
    claim.finish()                                             # rank 0 hosts the store: it waits here for the other rank


def test_dynamic_shards_claim_every_batch_exactly_once(tmp_path):
    """torchrun sweep with --batch_size > 1: batches of the global sample list are claimed from a counter all ranks share
    (TCPStore on MASTER_PORT + 1): every batch exactly once, the slower rank fewer of them."""
    import multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_claim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0] + res[1]) == list(range(12))
    assert len(res[0]) >= len(res[1])                        # rank 1 sleeps between claims


def test_batched_sweep_with_a_claim_function_covers_the_same_samples(tmp_path):
    """run_sweep_batched(claim=...) over the GLOBAL list: two "ranks" sharing one counter write, between them, exactly the
    files one rank writes alone (host logic with the stub extractor)."""
    names = [f"s{i:02d}" for i in range(11)]
    out = {}
    for mode in ("alone", "shared"):
        args = X.build_parser().parse_args(["--dataset", "meld", "--modality", "audio", "--audio_root", str(tmp_path / "a"),
                                            "--save_root", str(tmp_path / mode), "--quiet"])
        os.makedirs(tmp_path / "a", exist_ok=True)
        for n in names:
            (tmp_path / "a" / f"{n}.wav").write_bytes(b"x")
        if mode == "alone":
            X.run_sweep_batched(args, BatchStub(), names, batch_size=3, workers=2, log=lambda *_: None)
        else:
            counter = iter(range(100))
            claim = lambda: next(counter)                                             # noqa: E731 - one shared counter
            for _rank in range(2):                                                    # second "rank" finds nothing left
                X.run_sweep_batched(args, BatchStub(), names, batch_size=3, workers=2, log=lambda *_: None, claim=claim)
        out[mode] = sorted(os.path.relpath(os.path.join(d, f), tmp_path / mode) for d, _, fs in os.walk(tmp_path / mode) for f in fs)
    assert out["alone"] == out["shared"] and len(out["alone"]) == len(names)
