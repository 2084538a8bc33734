"""Data-parallel sharding of clips across ranks (one process + one engine replica per GPU).

Clips are independent: rank r takes clips r, r+W, r+2W, ...; there is no data-path collective.
The only exchange step is the optional gather of fused tokens to the ranks that consume them
(``gather_clip_rows``: one all_gather over NCCL/NVLink, or gloo in CPU tests).  The reference has no analogue
for extraction (single GPU); its training uses DistributedSampler (A/my_affectgpt/runners/runner_base.py:548).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_indices(num_clips: int, rank: int, world_size: int) -> List[int]:
    """Strided split: clip i belongs to rank i % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("bad rank")
    return list(range(rank, num_clips, world_size))


def shard_sizes(num_clips: int, world_size: int) -> List[int]:
    return [len(range(r, num_clips, world_size)) for r in range(world_size)]


def gather_clip_rows(local_rows: torch.Tensor, num_clips: int, group=None) -> torch.Tensor:
    """local_rows [n_local, ...] for this rank's strided shard -> [num_clips, ...] on every rank, rows back in
    clip order.  Uneven shards are padded to the largest shard for the collective and trimmed after."""
    if not dist.is_available() or not dist.is_initialized():
        if local_rows.shape[0] != num_clips:
            raise ValueError("without a process group the local shard must be the whole batch")
        return local_rows
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(num_clips, world)
    if local_rows.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {sizes[rank]}")
    pad = max(sizes)
    buf = local_rows.new_zeros((pad,) + tuple(local_rows.shape[1:]))
    buf[: sizes[rank]] = local_rows
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf.contiguous(), group=group)
    out = local_rows.new_empty((num_clips,) + tuple(local_rows.shape[1:]))
    for r in range(world):
        out[r::world] = gathered[r][: sizes[r]]
    return out
