"""Multi-GPU sharding of the hot path: stereo pairs are independent units.

One process per GPU (``torchrun``); the batch of pairs is split into contiguous slices, weights are
replicated, and there is NO collective on the inference data path — each rank keeps the disparities
of its own pairs.  ``gather_pairs`` exists for callers that want all results on every rank (evaluation
scripts); it is the only place a collective appears and it is not part of the timed path in
``bench.py``.  The reference itself is single-process (stereo.py:32-34).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the pairs rank ``rank`` owns: contiguous slices whose sizes differ by at most one
    (the first ``n_pairs % world`` ranks get the extra pair); every pair is owned by exactly one rank."""
    if world <= 0 or not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("shard_range: bad rank/world/n_pairs")
    base, extra = divmod(n_pairs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_pairs(left: torch.Tensor, right: torch.Tensor, rank: int, world: int):
    """Slice a (B, ...) left/right batch to this rank's pairs (views, no copy)."""
    if left.shape[0] != right.shape[0]:
        raise ValueError("shard_pairs: left and right batches differ")
    b, e = shard_range(left.shape[0], rank, world)
    return left[b:e], right[b:e]


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1-process defaults)."""
    import os
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gather_pairs(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """All ranks' per-pair results concatenated in global pair order on every rank.
    ``local`` is this rank's (b_local, ...) tensor; slices may differ in size by one, so shorter ones
    are padded for the fixed-size all_gather and trimmed afterwards."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_pairs, r, world) for r in range(world)]
    mx = max(e - b for b, e in sizes)
    b, e = sizes[rank]
    if local.shape[0] != e - b:
        raise ValueError("gather_pairs: rank %d holds %d pairs, expected %d" % (rank, local.shape[0], e - b))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - b] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """Max of a per-rank scalar (bench timing: the job is as slow as its slowest rank)."""
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
