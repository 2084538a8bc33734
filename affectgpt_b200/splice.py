"""Token layout: where the fused tokens go inside the prompt embedding.

Reference: AffectGPT.forward, A/my_affectgpt/models/affectgpt.py:956-1009 (same rule in Chat.answer_sample,
conversation/conversation_video.py:330-360): for each of FRAME, FACE, AUDIO, MULTI, IMAGE in that order, the
placeholder id must occur exactly ``n_query`` times, consecutively; that run is overwritten by the modality's
tokens.  Positions are an integer contract (bit-exact); the copy itself is a batched index_copy on the GPU
instead of the reference's per-sample torch.cat chain.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

ORDER = ("frame", "face", "audio", "multi", "image")


def placeholder_positions(input_ids: torch.Tensor, token_ids: Dict[str, int], n_query: Dict[str, int],
                          present: Sequence[str]) -> Dict[str, torch.Tensor]:
    """input_ids [b, L] -> {name: start positions [b]} for every modality whose placeholder occurs.
    Raises ValueError / AssertionError on the conditions the reference raises on (:984-1002)."""
    out: Dict[str, torch.Tensor] = {}
    b, L = input_ids.shape
    ar = torch.arange(L, device=input_ids.device)
    for name in ORDER:
        tid = token_ids.get(name)
        if tid is None:
            continue
        mask = input_ids == tid
        counts = mask.sum(dim=1)
        if int(counts.max()) == 0:
            continue
        if name not in present:
            raise AssertionError(f"Some input info is missing: {name.upper()} embeds is None.")
        nq = n_query[name]
        if bool((counts != nq).any()):
            raise ValueError("The number of audio patch tokens should be the same as the number of audio patches.")
        first = torch.where(mask, ar, torch.full_like(ar, L)).min(dim=1).values            # [b]
        expect = (ar.unsqueeze(0) >= first.unsqueeze(1)) & (ar.unsqueeze(0) < (first + nq).unsqueeze(1))
        if bool((expect != mask).any()):
            raise ValueError("The image patch tokens should be consecutive.")
        out[name] = first
    return out


def splice_tokens(input_embeds: torch.Tensor, input_ids: torch.Tensor, token_ids: Dict[str, int],
                  n_query: Dict[str, int], embeds: Dict[str, torch.Tensor]) -> torch.Tensor:
    """input_embeds [b,L,D] (embedding of ids with placeholders zeroed), embeds[name] [b,n_q,D] -> spliced copy."""
    pos = placeholder_positions(input_ids, token_ids, n_query, list(embeds.keys()))
    out = input_embeds.clone()
    b = input_ids.shape[0]
    rows = torch.arange(b, device=input_ids.device)
    for name in ORDER:
        if name not in pos:
            continue
        nq = n_query[name]
        e = embeds[name].to(out.dtype)
        for j in range(nq):
            out[rows, pos[name] + j] = e[:, j]
    return out
