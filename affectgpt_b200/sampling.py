"""Host-side integer contracts that sit just before the encoders: which frames / face crops are sampled and
where the eight 2-second audio segments start.  These must match the reference bit-for-bit.

Reference:
  frame / face indices  A/my_affectgpt/processors/video_processor.py:210-244 (uniform, emotion_peak pad rule),
                        :59-164 (_calculate_smart_frame_indices)
  audio segmentation    A/my_affectgpt/models/ImageBind/data.py:70-78, 170-214 (+ pytorchvideo 0.1.5
                        ConstantClipsPerVideoSampler, exact Fraction arithmetic on the float duration)
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import List, Optional, Sequence, Tuple

import torch


def uniform_frame_indices(vlen: int, n_frms: int = 8) -> List[int]:
    """``np.arange(0, vlen, vlen / n).astype(int)`` with n = min(n_frms, vlen), padded by repeating the last
    index.  np.arange with a float step yields ceil(vlen / step) points start + i*step in float64; truncation
    toward zero is the ``astype(int)``."""
    if vlen <= 0:
        raise ValueError("empty video")
    n = min(n_frms, vlen)
    step = vlen / n                                   # float64, as numpy computes it
    count = int(math.ceil((vlen - 0) / step))
    idx = [int(0 + i * step) for i in range(count)]
    while len(idx) < n_frms:
        idx.append(idx[-1])
    return idx


def headtail_frame_indices(vlen: int, n_frms: int = 8, rng=None) -> List[int]:
    """sampling == 'headtail' (video_processor.py:213-216): n//2 random frames from each half, drawn with Python's
    ``random.sample`` in this order (first half, then second half) so that a caller who seeds ``random`` the way the
    reference process does gets the reference's indices."""
    import random as _random
    rng = rng or _random
    n = min(n_frms, vlen)
    head = sorted(rng.sample(range(vlen // 2), n // 2))
    tail = sorted(rng.sample(range(vlen // 2, vlen), n // 2))
    idx = head + tail
    while len(idx) < n_frms:
        idx.append(idx[-1])
    return idx


def emotion_peak_frame_indices(vlen: int, au_info: Optional[dict] = None, n_frms: int = 8) -> List[int]:
    """sampling == 'emotion_peak' (video_processor.py:217-244): without au_info the middle frame repeated."""
    idx = smart_frame_indices(au_info, vlen) if au_info else [vlen // 2]
    while len(idx) < n_frms:
        idx.append(idx[-1])
    return idx


def smart_frame_indices(au_info: Optional[dict], total_video_frames: int) -> List[int]:
    """Peak frame +- neighbours, then an even fill from the unused frames (video_processor.py:59-164)."""
    if not au_info or not au_info.get("peak_frames"):
        n = total_video_frames - 1
        # np.linspace(0, n, 8).astype(int): i * (n / 7) evaluated as numpy does (step first, endpoint exact)
        step = n / 7.0
        out = [int(i * step) for i in range(7)] + [int(n)]
        return sorted(out)
    pk = au_info["peak_frames"][0]
    p, before, after = pk["peak_index"], pk["frames_before_peak"], pk["frames_after_peak"]
    total = au_info["total_frames"]
    chosen = {p}
    lo = [p - k for k in (1, 2) if p >= k]
    hi = [p + k for k in (1, 2) if p + k < total]
    if before >= 2 and after >= 2:
        chosen.update(lo); chosen.update(hi)
    elif before == 1 and after >= 2:
        chosen.update(lo[:1]); chosen.update(hi)
    elif before >= 2 and after == 1:
        chosen.update(hi[:1]); chosen.update(lo)
    elif before == 1 and after == 1:
        chosen.update(lo[:1]); chosen.update(hi[:1])
    elif before == 0:
        chosen.update(hi)
    elif after == 0:
        chosen.update(lo)
    need = 8 - len(chosen)
    if need > 0:
        avail = [i for i in range(total) if i not in chosen]
        if avail:
            if len(avail) <= need:
                chosen.update(avail)
            else:
                step = len(avail) / need
                for i in range(need):
                    j = int(i * step)
                    if j < len(avail):
                        chosen.add(avail[j])
    while len(chosen) < 8 and len(chosen) < total:
        rest = [i for i in range(total) if i not in chosen]
        if not rest:
            break
        chosen.add(rest[0])
    res = sorted(chosen)
    if len(res) < 8:
        base = list(res)
        k = 0
        while len(res) < 8:
            res.append(base[k % len(base)])
            k += 1
        res.sort()
    return res[:8]


def audio_segment_bounds(num_samples: int, sample_rate: int = 16000, clip_duration: int = 2,
                         clips_per_video: int = 8) -> List[Tuple[int, int]]:
    """[start, end) sample indices of the clips ``load_audio`` cuts (after its zero padding of short waveforms).
    start_i = int(i * Fraction(max(D - clip, 0)) / (clips - 1) * sr) with D = n / sr as a Python float."""
    n = max(num_samples, clip_duration * sample_rate)      # zero-pad rule, data.py:195-199
    duration = n / sample_rate                             # float, data.py:203
    max_start = Fraction(max(duration - Fraction(clip_duration), 0))
    step = Fraction(max_start, max(clips_per_video - 1, 1))
    bounds = []
    for i in range(clips_per_video):
        s = step * i
        bounds.append((int(s * sample_rate), int((s + clip_duration) * sample_rate)))
        # the sampler stops early only if the next start would exceed max_start, impossible for i < clips
    return bounds


def segment_audio(waveform: torch.Tensor, sample_rate: int = 16000, clip_duration: int = 2,
                  clips_per_video: int = 8) -> torch.Tensor:
    """waveform [channels, n] at ``sample_rate`` -> [clips, 1, clip_duration*sample_rate] (data.py:188-212)."""
    if waveform.shape[0] == 2:
        waveform = waveform.mean(dim=0, keepdim=True)
    ch, n = waveform.shape
    need = clip_duration * sample_rate
    if n < need:
        pad = torch.zeros((ch, need), dtype=waveform.dtype, device=waveform.device)
        pad[:, :n] = waveform
        waveform = pad
    clips = [waveform[:, a:b] for a, b in audio_segment_bounds(waveform.shape[1], sample_rate, clip_duration, clips_per_video)]
    return torch.stack(clips, dim=0)
