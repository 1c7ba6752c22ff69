"""Rod-index sharding across ranks (one process per GPU) and the only collective of the whole path.

Rods are independent: rank g of G integrates the contiguous rod-index block
[floor(g*B/G), floor((g+1)*B/G)) and no data-path collective exists.  The Newton static-shape driver adds one
all-reduce of two scalars per iteration (sum of squared residuals, max |residual|) -- NCCL on the GPUs, gloo in the
CPU tests of this host-side logic.  (The native driver, sri_newton_static_shape, does the same reduction on the device
through sri_nccl_init; this module serves the torch-tensor driver in newton.py.)
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

    ONE collective: the pairs of all ranks are all-gathered (16 bytes each) and folded locally in rank order, so every
    rank obtains bit-identical norms.  Works on CUDA tensors (NCCL) and CPU tensors (gloo).  With a single process / no
    initialised process group it is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return norm2_and_max
    world = dist.get_world_size(group)
    gathered = torch.empty((world, 2), dtype=norm2_and_max.dtype, device=norm2_and_max.device)
    dist.all_gather_into_tensor(gathered, norm2_and_max.reshape(1, 2).contiguous(), group=group)
    total = gathered[0, 0].clone()
    for r in range(1, world):  # rank order, not a tree: identical bits everywhere
        total = total + gathered[r, 0]
    norm2_and_max[0] = total
    norm2_and_max[1] = gathered[:, 1].max()
    return norm2_and_max
