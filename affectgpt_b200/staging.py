"""GPU-side input staging with the reference's loader signatures.

Reference: ``load_face`` (A/my_affectgpt/processors/video_processor.py:262-296) and ``load_audio``
(A/my_affectgpt/models/ImageBind/data.py:170-214).  Decoding (np.load / torchaudio / decord) stays with the caller; what
runs here, as CUDA kernels behind the C ABI, is everything after it: the bilinear resize of the SAMPLED face crops only
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


_BANKS: dict = {}


def read_wav(audio_path: str):
    """-> (float32 [channels, n], sample rate of the file) with torchaudio.load semantics (int PCM scaled to [-1, 1));
    resampling to 16 kHz (data.py:186-187) happens on the GPU in ``staging.stage_audio``."""
    from scipy.io import wavfile
    sr, data = wavfile.read(audio_path)
    if data.dtype == np.int16:
        wav = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        wav = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        wav = (data.astype(np.float32) - 128.0) / 128.0
    else:
        wav = data.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(wav.T if wav.ndim == 2 else wav[None])), int(sr)


def resample_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The windowed-sinc filter bank of ``torchaudio.functional.resample`` (sinc_interp_hann, its defaults - what
    ``load_audio`` calls at ImageBind/data.py:186-187), restated with the same float32 torch operations in the same
    order: -> (bank [new, 2*width + orig] float32 on the host, orig // gcd, new // gcd, width)."""
    import math
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("sample rates must be positive integers")
    key = (int(orig_freq), int(new_freq), lowpass_filter_width, rolloff)
    if key in _BANKS:
        return _BANKS[key]
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    dt = torch.float32                                                  # resample() passes waveform.dtype
    idx = torch.arange(-width, width + orig, dtype=dt)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=dt)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    out = (kernels[:, 0, :].contiguous(), orig, new, width)
    _BANKS[key] = out
    return out


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int = 16000, device="cuda:0") -> torch.Tensor:
    """``torchaudio.functional.resample(waveform, orig_freq, new_freq)`` on the GPU: [channels, n] -> [channels,
    ceil(new * n / orig)] float32 on ``device`` (same filter bank, fp32 FIR; differs from torchaudio's conv1d only by
    the summation order)."""
    import math
    dev = _cuda(device)
    if waveform.dim() != 2:
        raise ValueError(f"expected [channels, samples], got {tuple(waveform.shape)}")
    wav = waveform.to(device=dev, dtype=torch.float32).contiguous()
    if int(orig_freq) == int(new_freq):
        return wav
    bank, orig, new, width = resample_bank(orig_freq, new_freq)
    ch, n = wav.shape
    n_out = int(math.ceil(new * n / orig))
    bank_d = bank.to(dev)
    out = torch.empty((ch, n_out), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().agb_resample_audio(wav.data_ptr(), ch, n, bank_d.data_ptr(), orig, new, width, bank.shape[1],
                                                  out.data_ptr(), n_out, _stream(dev)), None, "agb_resample_audio")
    return out


def stage_audio(waveform: torch.Tensor, sample_rate: int = 16000, clip_duration: int = 2, clips_per_video: int = 8,
                device="cuda:0", orig_sample_rate: int = None) -> torch.Tensor:
    """waveform [channels, n] float32 -> [clips,1,clip_duration*sample_rate] on ``device`` (data.py:186-212: resampled to
    ``sample_rate`` when ``orig_sample_rate`` differs, two channels averaged, waveforms shorter than one clip
    zero-padded)."""
    dev = _cuda(device)
    if waveform.dim() != 2:
        raise ValueError(f"expected [channels, samples], got {tuple(waveform.shape)}")
    if orig_sample_rate is not None and int(orig_sample_rate) != int(sample_rate):
        waveform = resample(waveform, int(orig_sample_rate), int(sample_rate), dev)
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


def load_audio(audio_paths, device, sample_rate: int = 16000, clip_duration: int = 2, clips_per_video: int = 8):
    """Drop-in for the reference's ``load_audio`` (ImageBind/data.py:170-214): wav files -> [n, clips, 1,
    clip_duration * sample_rate] float32 on ``device``.  The files are decoded on the host (scipy; the reference uses
    torchaudio.load); resampling, mono mix-down, padding and slicing run on the GPU."""
    if audio_paths is None:
        return None
    outs = []
    for path in audio_paths:
        wav, sr = read_wav(path)
        outs.append(stage_audio(wav, sample_rate, clip_duration, clips_per_video, device, orig_sample_rate=sr))
    return torch.stack(outs, dim=0)
