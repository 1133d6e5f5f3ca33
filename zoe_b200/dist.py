"""Multi-GPU plumbing for the SW path: one process per GPU, contiguous index shards, no collective on the
data path (every (streamed, profiled) pair is independent -- SURVEY.md 8(e)).  torch.distributed is used only
for the barrier and for reducing timings / counters across ranks."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, last) slice of ``n`` items owned by ``rank`` (same rule as the C library's
    in-process multi-device split: first = n*rank//world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(local, rank: int, world: int):
    """Gather per-rank result arrays to rank 0 in index order (used only off the timed path)."""
    import torch.distributed as dist
    if world == 1:
        return [local]
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    return out
