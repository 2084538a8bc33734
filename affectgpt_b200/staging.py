"""GPU-side input staging with the reference's loader signatures.

Reference: ``load_face`` (A/my_affectgpt/processors/video_processor.py:262-296) and ``load_audio``
(A/my_affectgpt/models/ImageBind/data.py:170-214).  Decoding (np.load / torchaudio / decord) and resampling stay with
the caller; what moves to the device is everything after it: the cv2-exact bilinear resize of the *sampled* face crops
(the reference resizes every crop on the CPU and then samples 8), the mono mix-down, the 2-second zero padding and the
eight-slice gather.  Index rules run on the host and are bit-exact (``sampling``).  No CPU fallback.
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np
import torch

from . import _lib, sampling


def _stream(dev: torch.device) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


def _cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"staging runs on CUDA devices only (got {dev}); there is no CPU fallback")
    return dev


def stage_faces(faces: Union[np.ndarray, torch.Tensor], indices: Sequence[int], device="cuda:0") -> torch.Tensor:
    """faces [N,h,w,3] uint8 -> [T,3,224,224] uint8 on ``device``: ``cv2.resize(face, (224, 224))`` of crops ``indices``."""
    dev = _cuda(device)
    f = torch.as_tensor(faces)
    if f.dim() != 4 or f.shape[3] != 3 or f.dtype != torch.uint8:
        raise ValueError(f"expected a uint8 [N,h,w,3] crop stack, got {tuple(f.shape)} {f.dtype}")
    n, h, w, _ = f.shape
    idx = [int(i) for i in indices]
    if not idx or min(idx) < 0 or max(idx) >= n:
        raise IndexError(f"face indices {idx} out of range for {n} crops")
    f = f.to(dev).contiguous()
    idx_t = torch.tensor(idx, dtype=torch.int32, device=dev)
    out = torch.empty((len(idx), 3, 224, 224), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().agb_stage_faces(f.data_ptr(), n, h, w, idx_t.data_ptr(), len(idx), out.data_ptr(),
                                               _stream(dev)), None, "agb_stage_faces")
    return out


def load_face(faces: Union[str, np.ndarray, torch.Tensor], n_frms: int = 8, height: int = 224, width: int = 224,
              sampling_mode: str = "uniform", device="cuda:0", rng=None) -> torch.Tensor:
    """Drop-in for the reference's ``load_face``: -> [3, n_frms, 224, 224] float32 holding 0..255, on ``device``."""
    if (height, width) != (224, 224):
        raise ValueError("the encode path is fed 224x224 crops (height=224, width=224)")
    if isinstance(faces, str):
        faces = np.load(faces)
    vlen = len(faces)
    if sampling_mode == "uniform":
        idx = sampling.uniform_frame_indices(vlen, n_frms)
    elif sampling_mode == "headtail":
        idx = sampling.headtail_frame_indices(vlen, n_frms, rng)
    else:
        raise NotImplementedError(sampling_mode)                     # video_processor.py:284-285
    u8 = stage_faces(faces, idx, device)                             # [T,3,224,224]
    return u8.permute(1, 0, 2, 3).to(torch.float32)                  # '(C, T, H, W)' float, :292-293


def stage_audio(waveform: torch.Tensor, sample_rate: int = 16000, clip_duration: int = 2, clips_per_video: int = 8,
                device="cuda:0") -> torch.Tensor:
    """waveform [channels, n] float32 at ``sample_rate`` (already resampled) -> [clips,1,clip_duration*sample_rate] on
    ``device`` (data.py:188-212: two channels are averaged, waveforms shorter than one clip are zero-padded)."""
    dev = _cuda(device)
    if waveform.dim() != 2:
        raise ValueError(f"expected [channels, samples], got {tuple(waveform.shape)}")
    ch, n = waveform.shape
    wav = waveform.to(device=dev, dtype=torch.float32).contiguous()
    clip_len = clip_duration * sample_rate
    bounds = sampling.audio_segment_bounds(n, sample_rate, clip_duration, clips_per_video)
    starts = torch.tensor([a for a, _ in bounds], dtype=torch.int64, device=dev)
    out = torch.empty((clips_per_video, 1, clip_len), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().agb_stage_audio(wav.data_ptr(), ch, n, starts.data_ptr(), clips_per_video, clip_len,
                                               out.data_ptr(), _stream(dev)), None, "agb_stage_audio")
    return out
