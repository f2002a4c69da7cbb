"""Instance sharding across the GPUs of one box.

Robot instances are independent, so a batch of N instances is split into contiguous ranges, one per rank, and
every rank runs the same kernels on its own range with no collective on the data path.  The optional result
gather is the only place NCCL appears (``gather_results``), and it is off the timed path.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank ``rank``: [r*N/G, (r+1)*N/G) with integer floors (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def gather_results(local, group=None):
    """Optional: all-gather per-rank result slabs (tiled-SoA tensors of equal shape) with torch.distributed
    (NCCL on GPUs, gloo on CPU).  Returns the list of per-rank tensors in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    outs = [local.new_empty(local.shape) for _ in range(world)]
    dist.all_gather(outs, local.contiguous(), group=group)
    return outs
