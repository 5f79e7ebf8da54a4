"""Query sharding helpers for one-process-per-GPU runs (bench.py under torchrun).

Queries are independent, so the only cross-rank traffic is bookkeeping: the max over ranks of the
device-timed step and a host gather of per-rank counters.  No data-path collective exists.
``shard_range`` is the same contiguous split the library applies over the devices of one context
(``shard_queries`` in csrc/api.cu), so a multi-process run and a multi-device context agree on
which rank/device owns which query.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(Q: int, world: int, rank: int) -> Tuple[int, int]:
    """(q0, nq) of rank ``rank``: contiguous ranges, the first Q % world ranks get one extra query."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(Q, world)
    q0 = rank * base + min(rank, rem)
    return q0, base + (1 if rank < rem else 0)


def all_ranges(Q: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(Q, world, r) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the step time in ms); identity without an initialised group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj):
    """Host gather of small per-rank records (counters, stage times); list indexed by rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
