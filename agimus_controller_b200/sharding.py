"""Batch sharding across the GPUs of one box: contiguous slabs, no data-path collective.

Problems never interact, so rank ``r`` of ``G`` owns problems ``shard_range(B, G, r)`` and the
only communication is one gather of the per-problem statistics after the solve (NCCL on GPUs; the CPU tests
run the same code over gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, world: int, rank: int) -> range:
    """Contiguous slab of problem indices owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(B, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_stats(cost: torch.Tensor, iters: torch.Tensor, status: torch.Tensor, B: int) -> dict:
    """All ranks receive ``cost[B] f64, iters[B] i32, status[B] i32`` in global problem order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(cost=cost, iters=iters, status=status)
    world = dist.get_world_size()
    sizes = [len(shard_range(B, world, r)) for r in range(world)]
    pad = max(sizes)
    mine = torch.zeros(pad, 3, dtype=torch.float64, device=cost.device)
    n = cost.shape[0]
    mine[:n, 0] = cost
    mine[:n, 1] = iters.to(torch.float64)
    mine[:n, 2] = status.to(torch.float64)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    allv = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    return dict(cost=allv[:, 0].contiguous(), iters=allv[:, 1].to(torch.int32), status=allv[:, 2].to(torch.int32))
