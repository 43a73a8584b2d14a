"""Host-side logic of the multi-GPU path on CPU: two `gloo` ranks shard a batch of
stereo pairs, each processes only its slice, and the gathered result equals the unsharded one.
(The per-pair work here is a stand-in pure function of the pair; the CUDA path itself has no CPU
form and is covered by the `-m gpu` tests.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dsmnet_b200.shard import gather_pairs, max_over_ranks, shard_pairs, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _per_pair(left, right):
    """deterministic stand-in for 'disparity of each pair': depends on both images of the pair only"""
    return (left * 2.0 - right).flatten(1).cumsum(1)[:, ::7]


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        left = torch.randn(n_pairs, 3, 8, 12, generator=g); right = torch.randn(n_pairs, 3, 8, 12, generator=g)
        l, r = shard_pairs(left, right, rank, world)
        local = _per_pair(l, r)
        full = gather_pairs(local, n_pairs)
        ok = torch.equal(full, _per_pair(left, right))
        slow = max_over_ranks(1.0 + rank)
        q.put((rank, bool(ok), l.shape[0], slow))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 5, 1])
def test_two_rank_sharding_gloo(n_pairs):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert sum(n for _, _, n, _ in res) == n_pairs                    # every pair owned exactly once
    assert all(abs(s - 2.0) < 1e-12 for _, _, _, s in res)            # max over ranks seen by both


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n and (e - b) in (n // world, n // world + 1)
                cover += list(range(b, e))
            assert cover == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
