"""Multi-GPU plumbing: arenas never interact (the reference is one arena per process,
gameplay.hpp:47-55), so the batch shards into contiguous ranges of global arena ids, one
``sf_handle`` per GPU, with NO collective on the step path.  The only exchange is the
reduction of the episode statistics (``SF_FIELD_STATS``, int64[16]) -- NCCL all-reduce on the
GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard(n_total, rank, world):
    """Contiguous range of global arena ids owned by ``rank``: (env_id_base, n_local).  Seeds,
    levels and action streams depend on the global id only, so results do not depend on
    ``world``."""
    base = n_total // world
    extra = n_total % world
    n_local = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, n_local


def reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-handle statistics over all ranks (in place); a no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device) -> float:
    """Device timings are reported as the maximum over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
