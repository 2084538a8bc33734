"""Pre-extracted feature sweep: drop-in for the reference's ``extract_multimodal_features_precompute.py``.

Same command line (``--dataset --modality --video_root --face_root --audio_root --sample_list/--csv_path --save_root
--visual_encoder --acoustic_encoder --frame_n_frms --frame_sampling --clips_per_video --skip-multi-preextract
--mer-factory-output --device --quiet``), same on-disk layout and semantics
(A/extract_multimodal_features_precompute.py:823-849, 892-1024, 1054-1098):

    <save_root>/<dataset>/frame_<visual>_<sampling>_<n>frms/<name>.npy      fp32 [n, 768]
    <save_root>/<dataset>/face_<visual>_8frms/<name>.npy                    fp32 [8, 768]     (always 8 frames, uniform)
    <save_root>/<dataset>/audio_<acoustic>_<clips>clips/<name>.npy          fp32 [clips, 1024]
    <save_root>/<dataset>/multi_<visual>_<acoustic>_complete/<name>.npy     fp32 [1024]

* a file that already exists is skipped (resume), a failed sample is skipped, except audio, which is zero-filled
  ``[clips, 1024]`` when the wav is missing or its processing fails (:965-968, 970-975);
* every exception inside an ``extract_*`` call is swallowed and turned into ``None`` (:227-229, 587-589, 610-612);
* the loaders the directories are read by: A/my_affectgpt/datasets/datasets/base_dataset.py:392-405, 481-492, 520-531.

What differs is where the work runs: frames, face crops and audio segments are staged on the GPU (``staging``), the
encoders and fusion heads are the sm_100a kernels behind the C ABI, and ``--batch_size`` samples (default 32) share one
ViT call and one HuBERT call while a thread pool decodes the next batch and writes the previous one
(``run_sweep_batched``; ``--batch_size 1`` is the reference's per-sample loop, ``run_sweep``).  Under
``torchrun`` each rank takes samples ``rank::world_size`` and writes its own files - no collective.
Decoding (mp4 / wav) is outside the path: videos are read with decord when it is installed, else OpenCV, else a
pre-decoded ``<video_root>/<name>.npy`` uint8 stack ``[N,h,w,3]`` (RGB); wavs with scipy.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import sampling, staging
from .parallel import shard_indices

DATASETS = ['mer2023', 'mer2024', 'mercaptionplus', 'cmumosei', 'cmumosi', 'iemocapfour', 'meld', 'sims', 'simsv2']


# --------------------------------------------------------------------------------------------- decoding (host, out of path)
def _au_info(video_name: Optional[str], mer_factory_output: Optional[str]) -> Optional[dict]:
    """video_processor.py:29-56 (_load_au_info)."""
    import json
    if not video_name or not mer_factory_output:
        return None
    path = os.path.join(mer_factory_output, video_name, f"{video_name}_au_analysis.json")
    if not os.path.exists(path):
        return None
    try:
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f).get("au_info")
    except Exception:
        return None


def frame_indices(vlen: int, n_frms: int, sampling_mode: str, video_name: Optional[str] = None,
                  mer_factory_output: Optional[str] = None) -> List[int]:
    """The index rule of the extractor's extract_frame_features (A/extract_multimodal_features_precompute.py:160-176).

    uniform / headtail go through load_video (video_processor.py:204-244).  ``emotion_peak`` with a video name does NOT:
    the extractor branches to extract_frame_features_smart (:174-175, :231-258), i.e. calculate_smart_frame_indices
    (:389-526) on the clip's au_info - always 8 frames, and ``np.linspace(0, vlen - 1, 8)`` when the AU json is missing,
    has no au_info or no peak_frames (:400-403).  Only without a video name does it fall through to load_video, whose
    emotion_peak rule without au_info repeats the middle frame (video_processor.py:229-237)."""
    if sampling_mode == "uniform":
        return sampling.uniform_frame_indices(vlen, n_frms)
    if sampling_mode == "headtail":
        return sampling.headtail_frame_indices(vlen, n_frms)
    if sampling_mode == "emotion_peak":
        if video_name:
            return sampling.smart_frame_indices(_au_info(video_name, mer_factory_output), vlen)
        return sampling.emotion_peak_frame_indices(vlen, None, n_frms)
    raise NotImplementedError(sampling_mode)


def read_video_frames(video_path: str, pick: Callable[[int], List[int]]) -> np.ndarray:
    """-> uint8 [T,224,224,3] RGB frames ``pick(vlen)`` of the video, resized to 224x224 by the decoder."""
    stack = os.path.splitext(video_path)[0] + ".npy"
    if not os.path.exists(video_path) and os.path.exists(stack):
        frames = np.load(stack, mmap_mode="r")
        idx = pick(len(frames))
        sel = np.ascontiguousarray(frames[idx])
        if sel.shape[1:3] != (224, 224):
            import cv2
            sel = np.stack([cv2.resize(f, (224, 224)) for f in sel])
        return sel
    try:
        import decord                                     # the reference's decoder (video_processor.py:204-205)
        vr = decord.VideoReader(uri=video_path, height=224, width=224)
        idx = pick(len(vr))
        return vr.get_batch(idx).asnumpy()
    except ImportError:
        pass
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise IOError(f"cannot open {video_path}")
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(fr)
    cap.release()
    if not frames:
        raise IOError(f"no frames decoded from {video_path}")
    idx = pick(len(frames))
    return np.stack([cv2.resize(cv2.cvtColor(frames[i], cv2.COLOR_BGR2RGB), (224, 224)) for i in idx])


read_wav = staging.read_wav          # wav decode lives next to the loader it feeds (staging.load_audio)


# --------------------------------------------------------------------------------------------- extractor
class MultimodalFeatureExtractor:
    """Same method surface as the reference's extractor class (extract_*_features -> np.ndarray or None)."""

    def __init__(self, visual_encoder=None, acoustic_encoder=None, fusion=None, device: str = "cuda:0",
                 mer_factory_output: Optional[str] = None):
        self.device = torch.device(device)
        self.encoders = {"visual": visual_encoder, "acoustic": acoustic_encoder}
        self.multi_fusion_model = fusion
        self.mer_factory_output = mer_factory_output
        self._side = None                                      # staging stream of run_sweep_batched
        self._pinned, self._pinned_busy, self._pinned_turn = [None, None], [None, None], 0

    # ---- single-sample calls, reference signatures
    def extract_frame_features(self, video_path, n_frms=8, sampling="uniform", video_name=None):
        try:
            frames = read_video_frames(video_path, lambda vlen: frame_indices(vlen, n_frms, sampling, video_name,
                                                                              self.mer_factory_output))
            u8 = torch.from_numpy(frames).to(self.device).permute(0, 3, 1, 2).contiguous()       # [T,3,224,224] u8
            return self._visual(u8[None])[0]
        except Exception as e:                                                                  # :227-229
            print(f"Error extracting frame features from {video_path}: {e}")
            return None

    def extract_face_features(self, face_npy_path, n_frms=8):
        try:
            faces = np.load(face_npy_path)
            idx = sampling.uniform_frame_indices(len(faces), n_frms)
            u8 = staging.stage_faces(faces, idx, self.device)                                    # [T,3,224,224] u8
            return self._visual(u8[None])[0]
        except Exception as e:                                                                  # :587-589
            print(f"Error extracting face features from {face_npy_path}: {e}")
            return None

    def extract_audio_features(self, audio_path, clips_per_video=8):
        try:
            wav, sr = read_wav(audio_path)
            raw = staging.stage_audio(wav, clips_per_video=clips_per_video, device=self.device, orig_sample_rate=sr)
            with torch.no_grad():
                feats = self.encoders["acoustic"](None, raw[None])                               # [1,T,1024]
            return feats[0].float().cpu().numpy()
        except Exception:                                                                        # :610-612 (silent)
            return None

    def extract_multi_features(self, face_features, audio_features):
        try:
            if self.multi_fusion_model is None:
                raise RuntimeError("Multi fusion model not loaded.")
            f = torch.from_numpy(np.asarray(face_features)).float()[None].to(self.device)
            a = torch.from_numpy(np.asarray(audio_features)).float()[None].to(self.device)
            multi_hidden, _ = self.multi_fusion_model.encode_multi_merge(f, a)
            return multi_hidden[0].float().cpu().numpy()
        except Exception as e:
            print(f"Error extracting multi features: {e}")
            return None

    # ---- batched calls (run_sweep_batched).  Host decoding (pool threads) -> prestage (H2D + device staging of a whole
    # batch on a side stream) -> submit (encoders + multi head, asynchronous on the calling stream) -> collect (D2H).
    # The sweep prestages batch k+1 between submit(k) and collect(k), so copies and staging kernels overlap the encoders.
    def decode_frames(self, video_path, n_frms=8, sampling="uniform", video_name=None) -> Optional[np.ndarray]:
        """-> uint8 [T,224,224,3] (host)."""
        try:
            return read_video_frames(video_path, lambda vlen: frame_indices(vlen, n_frms, sampling, video_name,
                                                                            self.mer_factory_output))
        except Exception as e:
            print(f"Error extracting frame features from {video_path}: {e}")
            return None

    def decode_faces(self, face_npy_path, n_frms=8) -> Optional[np.ndarray]:
        """-> uint8 [n_frms,h,w,3] (host): only the sampled crops travel to the device."""
        try:
            faces = np.load(face_npy_path, mmap_mode="r")
            idx = sampling.uniform_frame_indices(len(faces), n_frms)
            return np.ascontiguousarray(faces[idx])
        except Exception as e:
            print(f"Error extracting face features from {face_npy_path}: {e}")
            return None

    def decode_audio(self, audio_path):
        """-> (float32 [channels, n], sample rate) on the host."""
        try:
            return read_wav(audio_path)
        except Exception:
            return None

    def prestage(self, decoded: List[dict], clips_per_video: int = 8) -> dict:
        """decoded[i] = {'frame': u8 [T,224,224,3], 'face': u8 [8,h,w,3], 'audio': wav [ch,n]} (any subset) -> device
        tensors of the whole batch, produced on a side stream: frames in ONE pinned-buffer copy, face crops resized
        there (cv2-exact), audio mixed down / padded / sliced there.  A sample whose staging fails becomes None."""
        dev = self.device
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        n = len(decoded)
        st = {"n": n, "frame_owner": [], "face_owner": [], "audio_owner": [], "frames": None, "faces": None, "audio": None}
        with torch.cuda.device(dev), torch.cuda.stream(self._side):
            fr = [(i, d["frame"]) for i, d in enumerate(decoded) if d.get("frame") is not None]
            if fr:
                total = sum(f.shape[0] for _, f in fr)
                slot = self._pinned_slot(total)
                off = 0
                for i, f in fr:
                    slot[off:off + f.shape[0]].numpy()[...] = f
                    st["frame_owner"].append((i, f.shape[0]))
                    off += f.shape[0]
                st["frames"] = slot[:total].to(dev, non_blocking=True).permute(0, 3, 1, 2).contiguous()
                self._pinned_busy[self._pinned_turn].record(self._side)
            faces = []
            for i, d in enumerate(decoded):
                if d.get("face") is None:
                    continue
                try:
                    faces.append(staging.stage_faces(d["face"], list(range(len(d["face"]))), dev))
                    st["face_owner"].append((i, faces[-1].shape[0]))
                except Exception as e:
                    print(f"Error staging face crops of sample {i}: {e}")
            if faces:
                st["faces"] = torch.cat(faces, dim=0)
            segs = []
            for i, d in enumerate(decoded):
                if d.get("audio") is None:
                    continue
                try:
                    wav, sr = d["audio"]
                    segs.append(staging.stage_audio(wav, clips_per_video=clips_per_video, device=dev, orig_sample_rate=sr))
                    st["audio_owner"].append(i)
                except Exception:
                    pass
            if segs:
                st["audio"] = torch.stack(segs, dim=0)
            st["ready"] = torch.cuda.Event()
            st["ready"].record(self._side)
        return st

    def _pinned_slot(self, frames: int) -> torch.Tensor:
        """Two pinned uint8 [frames,224,224,3] upload slots used alternately; a slot is rewritten only after the copy
        that last read it has finished."""
        self._pinned_turn ^= 1
        t = self._pinned_turn
        if self._pinned[t] is None or self._pinned[t].shape[0] < frames:
            self._pinned[t] = torch.empty((frames, 224, 224, 3), dtype=torch.uint8).pin_memory()
            self._pinned_busy[t] = torch.cuda.Event()
        else:
            self._pinned_busy[t].synchronize()
        return self._pinned[t]

    def submit_batch(self, staged: dict, multi_of: List[int]) -> dict:
        """Encoders over everything staged (ONE ViT call for all frames and face crops, ONE HuBERT call) and the multi
        head for the samples listed in ``multi_of`` that have both face and audio features; asynchronous on the current
        stream.  Returns device tensors; ``collect_batch`` brings them to the host."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(staged["ready"])
        h = {"staged": staged, "visual": None, "audio": None, "multi": None, "multi_idx": []}
        parts = [t for t in (staged["frames"], staged["faces"]) if t is not None]
        for t in parts + ([staged["audio"]] if staged["audio"] is not None else []):
            t.record_stream(cur)                                   # allocated on the side stream, read here
        with torch.no_grad():
            if parts:
                allf = torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]
                enc = self.encoders["visual"]
                h["visual"] = (enc._ensure_engine(dev).clip_forward(allf) if hasattr(enc, "_ensure_engine")
                               else enc(None, allf.permute(1, 0, 2, 3)[None].float())[0])
            if staged["audio"] is not None:
                h["audio"] = self.encoders["acoustic"](None, staged["audio"])                     # [n,clips,1024]
            n_frame = sum(t for _, t in staged["frame_owner"])
            face_row, off = {}, n_frame
            for i, t in staged["face_owner"]:
                face_row[i] = (off, t)
                off += t
            audio_row = {i: j for j, i in enumerate(staged["audio_owner"])}
            idx = [i for i in multi_of if i in face_row and i in audio_row]
            if idx:
                if self.multi_fusion_model is None:
                    raise RuntimeError("Multi fusion model not loaded.")
                f = torch.stack([h["visual"][face_row[i][0]:face_row[i][0] + face_row[i][1]] for i in idx])
                a = torch.stack([h["audio"][audio_row[i]] for i in idx])
                h["multi"], _ = self.multi_fusion_model.encode_multi_merge(f, a)
                h["multi_idx"] = idx
        return h

    def collect_batch(self, h: dict):
        """-> (frame feats, face feats, audio feats, multi feats): lists over the batch of fp32 numpy arrays or None."""
        st = h["staged"]
        n = st["n"]
        ff, cf, af, mf = [None] * n, [None] * n, [None] * n, [None] * n
        if h["visual"] is not None:
            v = h["visual"].float().cpu().numpy()
            off = 0
            for owners, dst in ((st["frame_owner"], ff), (st["face_owner"], cf)):
                for i, t in owners:
                    dst[i] = v[off:off + t].copy()
                    off += t
        if h["audio"] is not None:
            a = h["audio"].float().cpu().numpy()
            for j, i in enumerate(st["audio_owner"]):
                af[i] = a[j].copy()
        if h["multi"] is not None:
            m = h["multi"].float().cpu().numpy()
            for j, i in enumerate(h["multi_idx"]):
                mf[i] = m[j].copy()
        return ff, cf, af, mf

    # ---- helpers
    def _visual(self, u8_btchw: torch.Tensor) -> np.ndarray:
        """[b,T,3,224,224] uint8 -> [b,T,768] fp32 numpy.  The encoder wrapper takes the reference's float layout
        [b,3,t,224,224]; the engine behind it also accepts uint8 frames directly, which skips the float round trip."""
        enc = self.encoders["visual"]
        b, t = u8_btchw.shape[:2]
        with torch.no_grad():
            if hasattr(enc, "_ensure_engine"):
                eng = enc._ensure_engine(u8_btchw.device)
                feats = eng.clip_forward(u8_btchw.reshape(b * t, 3, 224, 224)).reshape(b, t, -1)
            else:
                feats = enc(None, u8_btchw.permute(0, 2, 1, 3, 4).float())
        return feats.float().cpu().numpy()


# --------------------------------------------------------------------------------------------- the sweep
def save_dirs(args) -> Dict[str, str]:
    """Directory names of extract_multimodal_features_precompute.py:823-849."""
    root = os.path.join(args.save_root, args.dataset)
    d: Dict[str, str] = {}
    if args.modality in ("frame", "all"):
        d["frame"] = os.path.join(root, f"frame_{args.visual_encoder}_{args.frame_sampling}_{args.frame_n_frms}frms")
    if args.modality in ("face", "all"):
        d["face"] = os.path.join(root, f"face_{args.visual_encoder}_8frms")
    if args.modality in ("audio", "all", "multi"):
        d["audio"] = os.path.join(root, f"audio_{args.acoustic_encoder}_{args.clips_per_video}clips")
    if args.modality in ("multi", "all") and not args.skip_multi_preextract:
        d["multi"] = os.path.join(root, f"multi_{args.visual_encoder}_{args.acoustic_encoder}_complete")
    for p in d.values():
        os.makedirs(p, exist_ok=True)
    return d


def run_sweep(args, extractor, sample_names: List[str], log=print) -> Dict[str, int]:
    """The per-sample loop of the reference (:892-1024) over this rank's samples."""
    dirs = save_dirs(args)
    stats = {k: 0 for k in ("frame_success", "face_success", "audio_success", "audio_failed", "multi_success", "multi_failed",
                            "skipped_existing")}
    quiet = args.quiet
    for name in sample_names:
        if "frame" in dirs:
            out = os.path.join(dirs["frame"], f"{name}.npy")
            if os.path.exists(out):
                stats["skipped_existing"] += 1
            else:
                video = os.path.join(args.video_root, f"{name}.mp4")
                if os.path.exists(video) or os.path.exists(os.path.splitext(video)[0] + ".npy"):
                    f = extractor.extract_frame_features(video_path=video, n_frms=args.frame_n_frms,
                                                         sampling=args.frame_sampling, video_name=name)
                    if f is not None:
                        np.save(out, f)
                        stats["frame_success"] += 1
                        if not quiet:
                            log(f"Frame: {name} -> {f.shape}")
        if "face" in dirs:
            out = os.path.join(dirs["face"], f"{name}.npy")
            if os.path.exists(out):
                stats["skipped_existing"] += 1
            else:
                face_npy = os.path.join(args.face_root, name, f"{name}.npy")
                if os.path.exists(face_npy):
                    f = extractor.extract_face_features(face_npy, n_frms=8)
                    if f is not None:
                        np.save(out, f)
                        stats["face_success"] += 1
                        if not quiet:
                            log(f"Face: {name} -> {f.shape}")
                elif not quiet:
                    log(f"Face: {name} -> file not found: {face_npy}")
        if args.modality in ("audio", "all"):
            out = os.path.join(dirs["audio"], f"{name}.npy")
            if os.path.exists(out):
                stats["skipped_existing"] += 1
            else:
                wav = os.path.join(args.audio_root, f"{name}.wav")
                f = extractor.extract_audio_features(wav, clips_per_video=args.clips_per_video) if os.path.exists(wav) else None
                if f is not None:
                    np.save(out, f)
                    stats["audio_success"] += 1
                    if not quiet:
                        log(f"Audio: {name} -> {f.shape}")
                else:                                          # missing file or failed processing: zero-filled (:965-975)
                    np.save(out, np.zeros((args.clips_per_video, 1024), dtype=np.float32))
                    stats["audio_failed"] += 1
                    if not quiet:
                        log(f"Audio: {name} -> zero-padded")
        if "multi" in dirs:
            out = os.path.join(dirs["multi"], f"{name}.npy")
            if os.path.exists(out):
                stats["skipped_existing"] += 1
            else:
                face_npy = os.path.join(args.face_root, name, f"{name}.npy")
                wav = os.path.join(args.audio_root, f"{name}.wav")
                ff = extractor.extract_face_features(face_npy, n_frms=8) if os.path.exists(face_npy) else None
                af = extractor.extract_audio_features(wav, clips_per_video=args.clips_per_video) if os.path.exists(wav) else None
                m = extractor.extract_multi_features(ff, af) if (ff is not None and af is not None) else None
                if m is not None:
                    np.save(out, m)
                    stats["multi_success"] += 1
                    if not quiet:
                        log(f"Multi: {name} -> {m.shape}")
                else:
                    stats["multi_failed"] += 1
    return stats


def run_sweep_batched(args, extractor, sample_names: List[str], batch_size: int = 32, workers: int = 8,
                      log=print, claim: Optional[Callable[[], Optional[int]]] = None) -> Dict[str, int]:
    """Same files, same statistics and same per-sample failure rules as ``run_sweep``, but the device sees batches:
    a thread pool decodes and stages the next batch (and writes the previous one) while the GPU encodes, all frames and face crops of ``batch_size`` samples
    go through ONE ViT call and all their audio segments through ONE HuBERT call, and in ``--modality all`` the multi
    features reuse the face / audio features of the same pass (the reference extracts them a second time, :1001-1003;
    the encoders are deterministic and batch-invariant bit for bit, so the files are identical).  If the device work
    of a batch raises, that batch is redone sample by sample through ``run_sweep`` so one bad sample cannot take its
    neighbours with it."""
    from concurrent.futures import ThreadPoolExecutor
    dirs = save_dirs(args)
    keys = ("frame_success", "face_success", "audio_success", "audio_failed", "multi_success", "multi_failed",
            "skipped_existing")
    stats = {k: 0 for k in keys}
    quiet = args.quiet
    want_audio_file = args.modality in ("audio", "all")

    def plan(name):
        """Which outputs are missing for this sample and which sources have to be decoded for them."""
        j = {"name": name, "skipped": 0, "out": {}, "src": {}, "notes": []}
        video = os.path.join(args.video_root, f"{name}.mp4") if args.video_root else None
        face_npy = os.path.join(args.face_root, name, f"{name}.npy") if args.face_root else None
        wav = os.path.join(args.audio_root, f"{name}.wav") if args.audio_root else None
        for kind in ("frame", "face", "audio", "multi"):
            if kind not in dirs or (kind == "audio" and not want_audio_file):
                continue
            out = os.path.join(dirs[kind], f"{name}.npy")
            if os.path.exists(out):
                j["skipped"] += 1
            else:
                j["out"][kind] = out
        if "frame" in j["out"]:
            if os.path.exists(video) or os.path.exists(os.path.splitext(video)[0] + ".npy"):
                j["src"]["frame"] = video
            else:
                del j["out"]["frame"]
        has_face, has_wav = bool(face_npy) and os.path.exists(face_npy), bool(wav) and os.path.exists(wav)
        if "face" in j["out"] and not has_face:
            del j["out"]["face"]
            j["notes"].append(f"Face: {name} -> file not found: {face_npy}")
        if has_face and ("face" in j["out"] or "multi" in j["out"]):
            j["src"]["face"] = face_npy
        if has_wav and ("audio" in j["out"] or "multi" in j["out"]):
            j["src"]["audio"] = wav
        return j

    def decode(j):
        d = {}
        if "frame" in j["src"]:
            d["frame"] = extractor.decode_frames(j["src"]["frame"], args.frame_n_frms, args.frame_sampling, j["name"])
        if "face" in j["src"]:
            d["face"] = extractor.decode_faces(j["src"]["face"], 8)
        if "audio" in j["src"]:
            d["audio"] = extractor.decode_audio(j["src"]["audio"])
        return d

    def finish(jobs, handle):
        """-> per job {kind: array} of everything that has to be written, plus the batch's statistics."""
        st = {k: 0 for k in keys}
        ff, cf, af, mf = extractor.collect_batch(handle)
        writes = []
        for i, j in enumerate(jobs):
            st["skipped_existing"] += j["skipped"]
            w, name = {}, j["name"]
            for note in j["notes"]:
                if not quiet:
                    log(note)
            if "frame" in j["out"] and ff[i] is not None:
                w["frame"] = ff[i]; st["frame_success"] += 1
            if "face" in j["out"] and cf[i] is not None:
                w["face"] = cf[i]; st["face_success"] += 1
            if "audio" in j["out"]:
                if af[i] is not None:
                    w["audio"] = af[i]; st["audio_success"] += 1
                else:                                              # missing file or failed processing (:965-975)
                    w["audio"] = np.zeros((args.clips_per_video, 1024), dtype=np.float32); st["audio_failed"] += 1
            if "multi" in j["out"]:
                if mf[i] is not None:
                    w["multi"] = mf[i]; st["multi_success"] += 1
                else:
                    st["multi_failed"] += 1
            if not quiet:
                for kind, arr in w.items():
                    log(f"{kind.capitalize()}: {name} -> " + ("zero-padded" if kind == "audio" and af[i] is None else f"{arr.shape}"))
            writes.append(w)
        return writes, st

    # Batch source: this rank's own list cut into batches, or - ``claim`` given - batches of the GLOBAL list claimed one at
    # a time from a counter all ranks share (dynamic sharding: a GPU that runs slower under the power cap simply claims
    # fewer batches; samples are independent and the encoders batch-invariant, so the files do not depend on who wrote them).
    if claim is None:
        batch_iter = iter([sample_names[i:i + batch_size] for i in range(0, len(sample_names), batch_size)])
    else:
        n_batches = (len(sample_names) + batch_size - 1) // batch_size

        def _claimed():
            while True:
                b = claim()
                if b is None or b >= n_batches:
                    return
                yield sample_names[b * batch_size:(b + 1) * batch_size]
        batch_iter = _claimed()
    pending_writes = []
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        def submit_decode(names):
            jobs = [plan(n) for n in names]
            return jobs, [pool.submit(decode, j) for j in jobs]

        def staged_of(item):
            """Wait for the pool, then copy + stage the batch on the side stream; an exception is kept for the consumer."""
            jobs, futs = item
            try:
                return jobs, extractor.prestage([f.result() for f in futs], args.clips_per_video)
            except Exception as e:
                return jobs, e

        def refill(queue):
            nb = next(batch_iter, None)
            if nb is not None:
                queue.append(submit_decode(nb))

        queue = []
        refill(queue)
        refill(queue)                                               # decode two batches ahead of the device
        cur = staged_of(queue.pop(0)) if queue else None
        while cur is not None:
            jobs, staged = cur
            refill(queue)
            try:
                if isinstance(staged, Exception):
                    raise staged
                handle = extractor.submit_batch(staged, [i for i, j in enumerate(jobs) if "multi" in j["out"]])
                cur = staged_of(queue.pop(0)) if queue else None     # next batch's H2D + staging overlap these encoders
                writes, st = finish(jobs, handle)
            except Exception as e:
                log(f"batch of {len(jobs)} failed on the device ({e}); redoing it sample by sample")
                st = run_sweep(args, extractor, [j["name"] for j in jobs], log=log)
                writes = []
                if cur is not None and cur[0] is jobs:
                    cur = staged_of(queue.pop(0)) if queue else None
            for j, w in zip(jobs, writes):
                for kind, arr in w.items():
                    pending_writes.append(pool.submit(np.save, j["out"][kind], arr))
            for k in keys:
                stats[k] += st[k]
        for f in pending_writes:
            f.result()
    return stats


def shared_batch_counter(rank: int, world: int, host: Optional[str] = None, port: Optional[int] = None):
    """-> claim() for ``run_sweep_batched``: a counter shared by all ranks of a torchrun job (a TCPStore hosted by rank 0 on
    MASTER_PORT + 1; no process group, no collective).  Every call returns the next unclaimed batch index."""
    import datetime
    from torch.distributed import TCPStore
    host = host or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 1)
    store = TCPStore(host, port, world, is_master=(rank == 0), timeout=datetime.timedelta(seconds=300))

    def claim() -> int:
        return int(store.add("agb_next_batch", 1)) - 1

    def finish(poll_s: float = 0.05, timeout_s: float = 3600.0) -> None:
        """Call once per rank after its last claim: rank 0 hosts the counter and must not exit before the others are done."""
        import time
        store.add("agb_done", 1)
        if rank == 0:
            t0 = time.time()
            while int(store.add("agb_done", 0)) < world and time.time() - t0 < timeout_s:
                time.sleep(poll_s)
    claim.store = store                                          # keep the server alive as long as the closure lives
    claim.finish = finish
    return claim


def read_sample_names(args) -> List[str]:
    """--sample_list (one name per line) or --csv_path / --csv_column (:1120-1123 and the csv branch)."""
    if bool(args.sample_list) == bool(args.csv_path):
        raise ValueError("Either --sample_list or --csv_path must be provided" if not args.sample_list
                         else "Cannot specify both --sample_list and --csv_path, choose one")
    if args.sample_list:
        with open(args.sample_list, "r") as f:
            return [ln.strip() for ln in f if ln.strip()]
    import pandas as pd
    df = pd.read_csv(args.csv_path)
    return [str(x) for x in df[args.csv_column].tolist()]


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="AffectGPT multimodal feature extraction (B200-native path)")
    p.add_argument("--dataset", type=str, required=True, choices=DATASETS)
    p.add_argument("--modality", type=str, default="all", choices=["frame", "face", "audio", "multi", "all"])
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--quiet", action="store_true")
    p.add_argument("--skip-multi-preextract", action="store_true", dest="skip_multi_preextract")
    p.add_argument("--video_root", type=str)
    p.add_argument("--face_root", type=str)
    p.add_argument("--audio_root", type=str)
    p.add_argument("--sample_list", type=str)
    p.add_argument("--csv_path", type=str)
    p.add_argument("--csv_column", type=str, default="names")
    p.add_argument("--save_root", type=str, default="./preextracted_features")
    p.add_argument("--mer-factory-output", type=str, dest="mer_factory_output")
    p.add_argument("--visual_encoder", type=str, default="CLIP_VIT_LARGE", choices=["CLIP_VIT_LARGE"])
    p.add_argument("--acoustic_encoder", type=str, default="HUBERT_LARGE", choices=["HUBERT_LARGE"])
    p.add_argument("--frame_n_frms", type=int, default=8)
    p.add_argument("--frame_sampling", type=str, default="uniform", choices=["uniform", "headtail", "emotion_peak"])
    p.add_argument("--clips_per_video", type=int, default=8)
    p.add_argument("--n_frms", type=int, default=8, help="Deprecated: use --frame_n_frms instead")
    p.add_argument("--fusion_ckpt", type=str, help="AffectGPT checkpoint (ckpt['model']) for the multi features")
    p.add_argument("--batch_size", type=int, default=32,
                   help="samples per device batch (not in the reference, which runs one sample per call); 1 = the "
                        "reference's per-sample loop")
    p.add_argument("--workers", type=int, default=8, help="host threads decoding / writing around the device batches")
    p.add_argument("--static_shards", action="store_true",
                   help="under torchrun: samples rank::world_size per GPU instead of batches claimed from a shared counter "
                        "(the default with --batch_size > 1: a GPU that runs slower under the power cap claims fewer batches)")
    return p


def validate_args(args) -> None:
    """The argument checks of the reference's main() (:1098-1123)."""
    if args.frame_n_frms == 8 and args.n_frms != 8:
        args.frame_n_frms = args.n_frms
    if args.modality in ("frame", "all") and not args.video_root:
        raise ValueError("video_root is required when extracting frame features")
    if args.modality in ("face", "all", "multi") and not args.face_root:
        raise ValueError("face_root is required when extracting face or multi features")
    if args.modality in ("audio", "all", "multi") and not args.audio_root:
        raise ValueError("audio_root is required when extracting audio or multi features")
    if args.frame_sampling == "emotion_peak" and args.modality in ("frame", "all") and not args.mer_factory_output:
        raise ValueError("mer_factory_output is required when using emotion_peak sampling")


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    validate_args(args)
    names = read_sample_names(args)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    claim = None
    if world > 1:                                            # one replica per GPU, no collective on the data path
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        args.device = f"cuda:{local}"
        if args.batch_size > 1 and not args.static_shards:
            claim = shared_batch_counter(rank, world)        # batches of the global list, first come first served
        else:
            names = [names[i] for i in shard_indices(len(names), rank, world)]      # samples rank::world
    from .fusion import AffectGPTFusion
    from .registry import registry
    visual = registry.get_visual_encoder_class(args.visual_encoder)()
    acoustic = registry.get_acoustic_encoder_class(args.acoustic_encoder)()
    fusion = None
    if args.modality in ("multi", "all") and not args.skip_multi_preextract:
        if not args.fusion_ckpt:
            raise ValueError("--fusion_ckpt is required to pre-extract multi features (or pass --skip-multi-preextract)")
        ckpt = torch.load(args.fusion_ckpt, map_location="cpu")
        fusion = AffectGPTFusion(visual, acoustic)
        fusion.load_state_dict(ckpt.get("model", ckpt))
    extractor = MultimodalFeatureExtractor(visual, acoustic, fusion, device=args.device,
                                           mer_factory_output=args.mer_factory_output)
    if args.batch_size > 1:
        stats = run_sweep_batched(args, extractor, names, batch_size=args.batch_size, workers=args.workers, claim=claim)
    else:
        stats = run_sweep(args, extractor, names)
    print(f"[rank {rank}/{world}] " + ", ".join(f"{k}={v}" for k, v in stats.items()))
    if claim is not None:
        claim.finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
