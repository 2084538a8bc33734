"""Integer contracts of the path (TEST INFRASTRUCTURE ONLY): frame indices, audio segment
boundaries, token splice positions.  Literal restatements of the reference's Python.

* ``uniform_indices``                <- A/my_affectgpt/processors/video_processor.py:210-216, 243-244
* ``emotion_peak_indices``           <- video_processor.py:217-237 (+ pad :243-244)
* ``smart_frame_indices``            <- video_processor.py:59-164
* ``ConstantClipsPerVideoSampler``   <- pytorchvideo==0.1.5 (A/environment.yml:250;
  un-vendored: restated from its published source, pytorchvideo/data/clip_sampling.py)
* ``clip_timepoints`` / ``audio_segments`` <- A/my_affectgpt/models/ImageBind/data.py:70-78, 170-214
* ``splice_positions`` / ``splice``  <- A/my_affectgpt/models/affectgpt.py:956-1009

Pinned by the known answers of SURVEY.md 8a (rows a1, a4) in tests/test_sampling.py.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def uniform_indices(vlen: int, n_frms: int = 8) -> List[int]:
    n = min(n_frms, vlen)
    indices = np.arange(0, vlen, vlen / n).astype(int).tolist()
    while len(indices) < n_frms:
        indices.append(indices[-1])
    return indices


def headtail_indices(vlen: int, n_frms: int = 8, rng=None) -> List[int]:
    """video_processor.py:213-216 + pad :243-244; ``rng`` defaults to the ``random`` module like the reference."""
    import random as rnd
    rng = rng or rnd
    n = min(n_frms, vlen)
    indices_h = sorted(rng.sample(range(vlen // 2), n // 2))
    indices_t = sorted(rng.sample(range(vlen // 2, vlen), n // 2))
    indices = indices_h + indices_t
    while len(indices) < n_frms:
        indices.append(indices[-1])
    return indices


def smart_frame_indices(au_info: Optional[dict], total_video_frames: int) -> List[int]:
    if not au_info or "peak_frames" not in au_info or len(au_info["peak_frames"]) == 0:
        return sorted(np.linspace(0, total_video_frames - 1, 8).astype(int).tolist())
    peak = au_info["peak_frames"][0]
    p = peak["peak_index"]
    before = peak["frames_before_peak"]
    after = peak["frames_after_peak"]
    total = au_info["total_frames"]
    sel = {p}

    def add_prev(k):
        if p >= k:
            sel.add(p - k)

    def add_next(k):
        if p + k < total:
            sel.add(p + k)

    if before >= 2 and after >= 2:
        add_prev(1); add_prev(2); add_next(1); add_next(2)
    elif (before == 1 and after >= 2) or (before >= 2 and after == 1):
        if before == 1:
            add_prev(1); add_next(1); add_next(2)
        else:
            add_next(1); add_prev(1); add_prev(2)
    elif before == 1 and after == 1:
        add_prev(1); add_next(1)
    elif before == 0 or after == 0:
        if before == 0:
            add_next(1); add_next(2)
        else:
            add_prev(1); add_prev(2)
    remaining = 8 - len(sel)
    if remaining > 0:
        avail = [i for i in range(total) if i not in sel]
        if len(avail) > 0:
            if len(avail) <= remaining:
                sel.update(avail)
            else:
                step = len(avail) / remaining
                for i in range(remaining):
                    idx = int(i * step)
                    if idx < len(avail):
                        sel.add(avail[idx])
    while len(sel) < 8 and len(sel) < total:
        avail = [i for i in range(total) if i not in sel]
        if avail:
            sel.add(avail[0])
        else:
            break
    res = sorted(sel)
    if len(res) < 8:
        orig = list(res)
        while len(res) < 8:
            for idx in orig:
                if len(res) >= 8:
                    break
                res.append(idx)
        res.sort()
    return res[:8]


def emotion_peak_indices(vlen: int, au_info: Optional[dict], n_frms: int = 8) -> List[int]:
    if au_info:
        indices = smart_frame_indices(au_info, vlen)
    else:
        indices = [vlen // 2]
    while len(indices) < n_frms:
        indices.append(indices[-1])
    return indices


class ConstantClipsPerVideoSampler:
    """pytorchvideo 0.1.5 sampler: ``clips_per_video`` evenly spaced clips of ``clip_duration``."""

    def __init__(self, clip_duration, clips_per_video: int, augs_per_clip: int = 1):
        self._clip_duration = Fraction(clip_duration)
        self._clips_per_video = clips_per_video
        self._augs_per_clip = augs_per_clip
        self._current_clip_index = 0
        self._current_aug_index = 0

    def __call__(self, last_clip_time, video_duration, annotation=None):
        max_possible_clip_start = Fraction(max(video_duration - self._clip_duration, 0))
        uniform_clip = Fraction(max_possible_clip_start, max(self._clips_per_video - 1, 1))
        clip_start_sec = uniform_clip * self._current_clip_index
        clip_index = self._current_clip_index
        aug_index = self._current_aug_index
        self._current_aug_index += 1
        if self._current_aug_index >= self._augs_per_clip:
            self._current_clip_index += 1
            self._current_aug_index = 0
        is_last_clip = False
        if (self._current_clip_index >= self._clips_per_video
                or uniform_clip * self._current_clip_index > max_possible_clip_start):
            self._current_clip_index = 0
            is_last_clip = True
        if is_last_clip:
            self._current_clip_index = 0
            self._current_aug_index = 0
        return clip_start_sec, clip_start_sec + self._clip_duration, clip_index, aug_index, is_last_clip


def clip_timepoints(duration: float, clip_duration=2, clips_per_video: int = 8):
    sampler = ConstantClipsPerVideoSampler(clip_duration, clips_per_video)
    out, is_last, end = [], False, 0.0
    while not is_last:
        start, end, _, _, is_last = sampler(end, duration, annotation=None)
        out.append((start, end))
    return out


def audio_segments(waveform: torch.Tensor, sample_rate: int = 16000, clip_duration=2,
                   clips_per_video: int = 8) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
    """data.py:170-214 after decode/resample: ``waveform`` [channels, n] at ``sample_rate``."""
    if waveform.shape[0] == 2:
        waveform = waveform.mean(dim=0, keepdim=True)
    ch, n = waveform.shape
    if n < 2 * sample_rate:
        tmp = torch.zeros([ch, 2 * sample_rate])
        tmp[:, :n] = waveform
        waveform = tmp
    bounds, clips = [], []
    for s, e in clip_timepoints(waveform.size(1) / sample_rate, clip_duration, clips_per_video):
        a, b = int(s * sample_rate), int(e * sample_rate)
        bounds.append((a, b))
        clips.append(waveform[:, a:b])
    return torch.stack(clips, dim=0), bounds


def splice_positions(input_ids: Sequence[int], token_ids: Dict[str, int],
                     n_query: Dict[str, int], present: Sequence[str]) -> Dict[str, int]:
    """affectgpt.py:968-1005: for (FRAME, FACE, AUDIO, MULTI, IMAGE) in that order."""
    ids = torch.as_tensor(list(input_ids))
    out = {}
    for name in ("frame", "face", "audio", "multi", "image"):
        tid = token_ids.get(name)
        if tid is None or int((ids == tid).sum()) == 0:
            continue
        if name not in present:
            raise AssertionError(f"Some input info is missing: {name.upper()} embeds is None.")
        nq = n_query[name]
        if int((ids == tid).sum()) != nq:
            raise ValueError("The number of audio patch tokens should be the same as the number of audio patches.")
        masked = torch.where(ids == tid)[0]
        start = int(masked[0])
        if (masked != torch.arange(start, start + nq)).any():
            raise ValueError("The image patch tokens should be consecutive.")
        out[name] = start
    return out


def splice(input_embeds: torch.Tensor, input_ids: torch.Tensor, token_ids: Dict[str, int],
           n_query: Dict[str, int], embeds: Dict[str, torch.Tensor]) -> torch.Tensor:
    """affectgpt.py:956-1009 for a batch: input_embeds [b,L,D], embeds[name] [b,n_q,D]."""
    outs = []
    for i in range(input_ids.shape[0]):
        cur = input_embeds[i]
        pos = splice_positions(input_ids[i].tolist(), token_ids, n_query, list(embeds.keys()))
        for name in ("frame", "face", "audio", "multi", "image"):
            if name in pos:
                s, nq = pos[name], n_query[name]
                cur = torch.cat((cur[:s], embeds[name][i], cur[s + nq:]), dim=0)
        outs.append(cur)
    return torch.stack(outs, dim=0)
