"""Data-parallel sharding of clips across ranks (one process + one engine replica per GPU).

Clips are independent: rank r takes clips r, r+W, r+2W, ...; there is no data-path collective.
The only exchange step is the optional gather of fused tokens to the ranks that consume them: one
``all_gather_into_tensor`` over NCCL/NVLink (gloo in the CPU tests) into pre-allocated buffers, followed by ONE strided
copy that puts the rows back in clip order.  ``ClipGather`` issues both on a side stream: the collective of step i
overlaps the encoders of step i+1 and never paces the compute stream (a blocking rendez-vous on the compute stream
makes every step wait for the slowest power-capped GPU of the node).  The reference has no analogue for extraction
(single GPU); its training uses DistributedSampler (A/my_affectgpt/runners/runner_base.py:548).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(num_clips: int, rank: int, world_size: int) -> List[int]:
    """Strided split: clip i belongs to rank i % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("bad rank")
    return list(range(rank, num_clips, world_size))


def shard_sizes(num_clips: int, world_size: int) -> List[int]:
    return [len(range(r, num_clips, world_size)) for r in range(world_size)]


def balanced_counts(num_clips: int, speeds: Sequence[float], max_shift: Optional[int] = None) -> List[int]:
    """Split ``num_clips`` over the ranks in proportion to their measured speed (clips per second, any unit).

    The GPUs of one node do not run at one speed: every B200 sits at its own point of the 1 kW power cap (SM clocks of
    1.26-1.42 GHz were seen on one box), and with equal shards the job runs at the pace of the slowest one.  Largest-
    remainder rounding; ``max_shift`` bounds how far a rank's share may move from the even split (None = unbounded).
    Deterministic for given speeds, so every rank computes the same list from the all-gathered speeds."""
    w = len(speeds)
    if w == 0 or any(not (s > 0) for s in speeds):
        raise ValueError("speeds must be positive")
    tot = float(sum(speeds))
    even = num_clips / w
    ideal = [num_clips * s / tot for s in speeds]
    if max_shift is not None:
        ideal = [min(max(x, even - max_shift), even + max_shift) for x in ideal]
        scale = num_clips / sum(ideal)
        ideal = [x * scale for x in ideal]
    counts = [int(x) for x in ideal]
    order = sorted(range(w), key=lambda i: (-(ideal[i] - counts[i]), i))
    for i in order[: num_clips - sum(counts)]:
        counts[i] += 1
    return counts


def _all_gather_flat(recv: torch.Tensor, send: torch.Tensor, group) -> None:
    """recv [W * n] <- every rank's send [n] (rank-major)."""
    dist.all_gather_into_tensor(recv, send, group=group)


class ClipGather:
    """Gather of per-clip rows for one fixed shape: buffers allocated once, collective + reorder on a side stream.

        g = ClipGather(num_clips, row_shape, dtype, device)
        g.start(local_rows)          # returns immediately; the calling stream is not blocked
        ... next step's kernels ...
        rows = g.wait()              # [num_clips, *row_shape] in clip order; the calling stream now waits for the gather

    ``start`` snapshots ``local_rows`` (device-to-device copy on the side stream after the producing work on the calling
    stream), so the caller may overwrite them with the next step.  A second ``start`` before ``wait`` queues behind the
    first one on the side stream and reuses the buffers: ``wait`` then returns the latest result.  On CPU tensors (gloo)
    everything runs synchronously."""

    def __init__(self, num_clips: int, row_shape: Sequence[int], dtype: torch.dtype, device, group=None,
                 counts: Optional[Sequence[int]] = None):
        """Default sharding: strided (clip i on rank i % W).  With ``counts`` (e.g. from ``balanced_counts``) rank r holds
        the contiguous block of ``counts[r]`` clips that starts at ``sum(counts[:r])``."""
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ClipGather needs an initialised process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.num_clips = int(num_clips)
        if counts is not None:
            if len(counts) != self.world or sum(counts) != self.num_clips or min(counts) < 0:
                raise ValueError(f"counts {list(counts)} do not split {self.num_clips} clips over {self.world} ranks")
            self.sizes = [int(c) for c in counts]
        else:
            self.sizes = shard_sizes(self.num_clips, self.world)
        self.contiguous = counts is not None
        self.pad = max(self.sizes) if self.sizes else 0
        self.row_shape = tuple(int(d) for d in row_shape)
        dev = torch.device(device)
        self.device = dev
        self.send = torch.zeros((self.pad,) + self.row_shape, dtype=dtype, device=dev)       # rows >= my shard stay zero
        self.recv = torch.empty((self.world, self.pad) + self.row_shape, dtype=dtype, device=dev)
        # clip i = j * W + r sits at recv[r, j]: viewed as [pad, W, ...] the output is one strided copy of recv^T;
        # the padding rows of uneven shards all land behind row num_clips
        self.out_padded = torch.empty((self.pad * self.world,) + self.row_shape, dtype=dtype, device=dev)
        self.index = None
        if self.contiguous:                                     # clip i of rank r's block sits at recv[r, i - offset_r]
            idx, off = [], 0
            for r, n in enumerate(self.sizes):
                idx.extend(r * self.pad + j for j in range(n))
                off += n
            self.index = torch.tensor(idx, dtype=torch.int64, device=dev)
        self.cuda = dev.type == "cuda"
        if self.cuda:
            self.stream = torch.cuda.Stream(device=dev)
            self.ready = torch.cuda.Event()
            self.done = torch.cuda.Event()
        self._pending = False

    def start(self, local_rows: torch.Tensor) -> None:
        n = self.sizes[self.rank]
        if tuple(local_rows.shape) != (n,) + self.row_shape:
            raise ValueError(f"rank {self.rank} holds rows of shape {tuple(local_rows.shape)}, expected {(n,) + self.row_shape}")
        if not self.cuda:
            self._run(local_rows)
            self._pending = True
            return
        cur = torch.cuda.current_stream(self.device)
        self.ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready)
            local_rows.record_stream(self.stream)
            self._run(local_rows)
            self.done.record(self.stream)
        self._pending = True

    def _run(self, local_rows: torch.Tensor) -> None:
        self.send[: self.sizes[self.rank]].copy_(local_rows, non_blocking=True)
        _all_gather_flat(self.recv.view(-1), self.send.view(-1), self.group)
        if self.contiguous:
            torch.index_select(self.recv.view((self.world * self.pad,) + self.row_shape), 0, self.index,
                               out=self.out_padded[: self.num_clips])
        else:
            self.out_padded.view((self.pad, self.world) + self.row_shape).copy_(self.recv.transpose(0, 1), non_blocking=True)

    def wait(self) -> torch.Tensor:
        if not self._pending:
            raise RuntimeError("ClipGather.wait() without a start()")
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(self.done)
        return self.out_padded[: self.num_clips]


def gather_clip_rows(local_rows: torch.Tensor, num_clips: int, group=None,
                     gatherer: Optional[ClipGather] = None) -> torch.Tensor:
    """local_rows [n_local, ...] for this rank's strided shard -> [num_clips, ...] on every rank, rows back in
    clip order (blocking convenience form of ``ClipGather``: start + wait).  Uneven shards are padded to the largest
    shard for the collective and trimmed after.  Pass a ``gatherer`` to reuse its buffers across calls."""
    if not dist.is_available() or not dist.is_initialized():
        if local_rows.shape[0] != num_clips:
            raise ValueError("without a process group the local shard must be the whole batch")
        return local_rows
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(num_clips, world)
    if local_rows.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {sizes[rank]}")
    if gatherer is None:
        gatherer = ClipGather(num_clips, local_rows.shape[1:], local_rows.dtype, local_rows.device, group)
    gatherer.start(local_rows.contiguous())
    return gatherer.wait()
