"""N>1 host path on CPU: strided clip sharding + the one gather, world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from affectgpt_b200 import parallel as P


def test_shard_indices_cover_all_clips_once():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 4, 8):
            parts = [P.shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert [len(p) for p in parts] == P.shard_sizes(n, w)


def test_balanced_counts_follow_speed_and_sum_up():
    assert P.balanced_counts(512, [1.0] * 8) == [64] * 8
    c = P.balanced_counts(512, [544, 539, 525, 537, 536, 542, 537, 530])          # clips/s of one 8 x B200 box
    assert sum(c) == 512 and c[0] == max(c) and c[2] == min(c) and max(c) - min(c) <= 4
    c = P.balanced_counts(128, [1.0, 3.0], max_shift=8)
    assert c == [56, 72]
    assert P.balanced_counts(7, [1.0, 1.0, 1.0]) in ([3, 2, 2],)
    with pytest.raises(ValueError):
        P.balanced_counts(8, [1.0, 0.0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(num_clips * 4 * 6, dtype=torch.float32).reshape(num_clips, 4, 6)   # [clips, tokens, D]
        mine = full[P.shard_indices(num_clips, rank, world)]
        got = P.gather_clip_rows(mine, num_clips)
        ok = bool(torch.equal(got, full))
        # the pre-allocated form, reused across steps (what bench.py / a consumer loop does)
        g = P.ClipGather(num_clips, (4, 6), torch.float32, "cpu")
        for step in range(3):
            g.start(mine + step)
            ok = ok and bool(torch.equal(g.wait(), full + step))
        # speed-balanced contiguous blocks (uneven on purpose)
        counts = P.balanced_counts(num_clips, [1.0, 1.6])
        off = sum(counts[:rank])
        gb = P.ClipGather(num_clips, (4, 6), torch.float32, "cpu", counts=counts)
        gb.start(full[off:off + counts[rank]].clone())
        ok = ok and bool(torch.equal(gb.wait(), full)) and sum(counts) == num_clips
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_clips", [8, 7, 1])
def test_gather_restores_clip_order_world2(num_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_without_process_group_is_identity():
    x = torch.randn(5, 3)
    assert P.gather_clip_rows(x, 5) is x
    with pytest.raises(ValueError):
        P.gather_clip_rows(x, 6)
