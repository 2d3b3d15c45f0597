"""Host-side logic of the multi-GPU path (one process per GPU).

Images are independent, so the path shards with no data-path collective:
  * `shard_range`      -- contiguous split of a global batch over ranks;
  * `broadcast_arena`  -- the ONE collective of a job: rank 0's prepared weight arena to every rank, at init;
  * `max_over_ranks`   -- device-time reduction used for reporting (time of the slowest rank);
  * `gather_to_rank0`  -- optional collection of per-rank detection tensors on rank 0.
All of it works on the gloo backend too (tests/test_parallel_gloo.py, world size 2, CPU).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the images rank `rank` owns: contiguous, sizes differ by at most one, every image exactly once."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_arena(arena, src: int = 0):
    """Broadcast the engine's weight arena (a uint8 torch tensor viewing `yb_network_weight_arena`) from rank `src`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def arena_tensor(ptr: int, nbytes: int, device):
    """Zero-copy torch view of a raw device allocation (CUDA array interface)."""
    import torch

    class _Arena:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Arena(), device=device)


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(local: np.ndarray, global_batch: int, device=None) -> Optional[np.ndarray]:
    """Concatenate per-rank result blocks (axis 0 = images of that rank's shard) on rank 0, in image order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(global_batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((most,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    bufs: List = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([bufs[r].cpu().numpy()[:hi - lo] for r, (lo, hi) in enumerate(sizes)], axis=0)
