"""Token layout: where the fused tokens go inside the prompt embedding.

Reference: AffectGPT.forward, A/my_affectgpt/models/affectgpt.py:956-1009 (same rule in Chat.answer_sample,
conversation/conversation_video.py:330-360): per sample, for each of FRAME, FACE, AUDIO, MULTI, IMAGE in that order,
if the placeholder id occurs in the sample it must occur exactly ``n_query`` times, consecutively; that run is
overwritten by the modality's tokens.  ``placeholder_positions`` is the integer contract (host logic, bit-exact);
``splice_tokens`` runs the scan + validation + copy as one CUDA kernel over the whole batch (``agb_splice_tokens``)
instead of the reference's per-sample torch.where / torch.cat chain, and raises the reference's exceptions from the
kernel's verdict.  There is no CPU path for the copy.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import torch

from . import _lib

ORDER = ("frame", "face", "audio", "multi", "image")
_MSG_COUNT = "The number of audio patch tokens should be the same as the number of audio patches."   # :997-998 (sic)
_MSG_CONSEC = "The image patch tokens should be consecutive."                                           # :1001-1002


def placeholder_positions(input_ids: torch.Tensor, token_ids: Dict[str, int], n_query: Dict[str, int],
                          present: Sequence[str]) -> Dict[str, torch.Tensor]:
    """input_ids [b, L] -> {name: start position per sample [b], -1 where the sample has no such placeholder} for every
    modality whose placeholder occurs at least once.  Raises what the reference raises (:984-1002)."""
    out: Dict[str, torch.Tensor] = {}
    b, L = input_ids.shape
    ar = torch.arange(L, device=input_ids.device)
    for name in ORDER:
        tid = token_ids.get(name)
        if tid is None:
            continue
        mask = input_ids == tid
        counts = mask.sum(dim=1)
        has = counts > 0
        if not bool(has.any()):
            continue
        if name not in present:
            raise AssertionError(f"Some input info is missing: {name.upper()} embeds is None.")
        nq = n_query[name]
        if bool((has & (counts != nq)).any()):
            raise ValueError(_MSG_COUNT)
        first = torch.where(mask, ar, torch.full_like(ar, L)).min(dim=1).values            # [b]
        expect = (ar.unsqueeze(0) >= first.unsqueeze(1)) & (ar.unsqueeze(0) < (first + nq).unsqueeze(1))
        if bool(((expect != mask) & has.unsqueeze(1)).any()):
            raise ValueError(_MSG_CONSEC)
        out[name] = torch.where(has, first, torch.full_like(first, -1))
    return out


def splice_tokens(input_embeds: torch.Tensor, input_ids: torch.Tensor, token_ids: Dict[str, int],
                  n_query: Dict[str, int], embeds: Dict[str, torch.Tensor], inplace: bool = False) -> torch.Tensor:
    """input_embeds [b,L,D] fp32 / bf16 on a CUDA device (embedding of the ids with placeholders zeroed),
    embeds[name] [b,n_q,D] -> spliced tensor (a copy unless ``inplace``)."""
    if not input_embeds.is_cuda:
        raise RuntimeError("splice_tokens runs on CUDA tensors only; affectgpt_b200 has no CPU fallback")
    if input_embeds.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"input_embeds must be float32 or bfloat16, got {input_embeds.dtype}")
    dev = input_embeds.device
    b, L, D = input_embeds.shape
    ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
    assert tuple(ids.shape) == (b, L)
    names = [n for n in ORDER if n in token_ids]
    if not names:
        return input_embeds if inplace else input_embeds.clone()
    if inplace and input_embeds.is_contiguous():
        # in place there is no copy to fall back on: validate every sample first (same exceptions, same order), so that a
        # ValueError / AssertionError never leaves a half-spliced tensor behind
        placeholder_positions(ids, token_ids, n_query, [n for n in names if embeds.get(n) is not None])
        out = input_embeds
    else:
        out = input_embeds.contiguous().clone()
    keep = []                                                   # staging tensors must outlive the launch
    tok = (C.c_int64 * len(names))(*[int(token_ids[n]) for n in names])
    nq = (C.c_int32 * len(names))(*[int(n_query.get(n, 0)) for n in names])
    ptrs = (C.c_void_p * len(names))()
    for i, n in enumerate(names):
        e = embeds.get(n)
        if e is None:
            ptrs[i] = None
            continue
        e = e.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(e.shape) != (b, int(n_query[n]), D):
            raise ValueError(f"{n} tokens have shape {tuple(e.shape)}, expected {(b, int(n_query[n]), D)}")
        keep.append(e)
        ptrs[i] = e.data_ptr()
    status = torch.empty((b, len(names), 3), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.check(_lib.load().agb_splice_tokens(out.data_ptr(), 0 if out.dtype == torch.float32 else 2, ids.data_ptr(), b, L,
                                                 D, len(names), tok, nq, ptrs, status.data_ptr(), int(stream.cuda_stream)),
                   None, "agb_splice_tokens")
        st = status.cpu()                                        # synchronises: the verdict decides the exceptions
    for i, n in enumerate(names):                                # same order and messages as the reference's loop
        cnt, consecutive = st[:, i, 0], st[:, i, 2]
        has = cnt > 0
        if not bool(has.any()):
            continue
        if embeds.get(n) is None:
            raise AssertionError(f"Some input info is missing: {n.upper()} embeds is None.")
        if bool((has & (cnt != int(n_query[n]))).any()):
            raise ValueError(_MSG_COUNT)
        if bool((has & (consecutive == 0)).any()):
            raise ValueError(_MSG_CONSEC)
    del keep
    return out
