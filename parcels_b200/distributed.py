"""Multi-GPU mode R (replicas): one process per GPU, the field replicated in every GPU's HBM,
particles sharded contiguously by rank.  Particles do not interact on this path, so the step loop
needs NO collective; ``torch.distributed`` is used only to assemble results / timings
(NCCL on GPUs, gloo in the CPU tests).  (the reference has no distributed layer.)
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [lo, hi) ranges: the first n % world ranks get one extra particle."""
    base, extra = divmod(int(n), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_particles(pdata: dict, rank: int, world: int) -> dict:
    """This rank's slice of a particle SoA dict (copies, C-contiguous)."""
    lo, hi = shard_bounds(len(pdata["x"]), world)[rank]
    return {k: np.ascontiguousarray(v[lo:hi]) for k, v in pdata.items()}


def gather_particles(local: dict, dist, dst: int = 0):
    """Concatenate every rank's SoA on ``dst`` in rank order (deleted particles already compacted
    locally, so shards may have shrunk).  Returns the merged dict on ``dst`` and None elsewhere."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    return {k: np.concatenate([b[k] for b in bucket], axis=0) for k in local}


def allreduce_max(value: float, dist, device="cpu") -> float:
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(value: float, dist, device="cpu") -> float:
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
