"""Batch sharding across GPUs: independent MPC solves shard with NO data-path
collective — contiguous batch slices per rank, model tables replicated per device.  The only
cross-rank traffic is the timing reduction (max over ranks) and an optional gather of results."""
from __future__ import annotations

import numpy as np


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous slice of the global batch owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, device=None) -> float:
    """Max over ranks of a per-rank duration (torch.distributed; identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_batches(local: np.ndarray, global_batch: int):
    """all_gather of per-rank result slices into the global batch order (results only; not on the
    data path of the solve)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local))
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == global_batch
    return out
