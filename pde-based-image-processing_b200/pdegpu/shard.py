"""Data-parallel sharding of independent problems (frame pairs, stereo pairs, PDE frames) over ranks.

The solver path has no exchange step for a batch: every rank relaxes its own contiguous
block of the batch, results are gathered on request. torch.distributed is plumbing only
(nccl on the GPU box, gloo in the CPU tests); there is no data-path collective.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def partition(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Balanced contiguous blocks: the first n_items % world ranks get one extra item."""
    if world < 1:
        raise ValueError("world must be >= 1")
    q, r = divmod(n_items, world)
    out, start = [], 0
    for k in range(world):
        stop = start + q + (1 if k < r else 0)
        out.append((start, stop))
        start = stop
    return out


def my_block(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    return partition(n_items, world)[rank]


def run_sharded(items: Sequence, fn: Callable, group=None) -> list:
    """Apply `fn` to this rank's block of `items` and return the results of ALL items, in order,
    on every rank (all_gather_object). Without an initialised process group: plain map."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [fn(x) for x in items]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = my_block(len(items), rank, world)
    local = [fn(x) for x in items[a:b]]
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    return [r for block in gathered for r in block]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """max over ranks of a per-rank scalar (elapsed device time): how multi-GPU numbers are reported."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
