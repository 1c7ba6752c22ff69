"""Rod-index sharding across ranks (one process per GPU) and the only collective of the whole path.

Rods are independent: rank g of G integrates the contiguous rod-index block
[floor(g*B/G), floor((g+1)*B/G)) and no data-path collective exists.  The Newton static-shape driver adds one
all-reduce of two scalars per iteration (sum of squared residuals, max |residual|) -- NCCL on the GPUs, gloo in the
CPU tests of this host-side logic.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open rod-index range owned by `rank`; the ranges of all ranks tile [0, total) exactly."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request: total={total} rank={rank} world={world}")
    return (rank * total) // world, ((rank + 1) * total) // world


def allreduce_residual(norm2_and_max, group=None):
    """In-place global reduction of a 2-element tensor [sum(rho^2), max|rho|] -> (sum over ranks, max over ranks).

    Works on CUDA tensors (NCCL) and CPU tensors (gloo).  With a single process / no initialised process group it is
    the identity."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return norm2_and_max
    s = norm2_and_max[0:1]
    m = norm2_and_max[1:2]
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return norm2_and_max
