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


def shard_batch(buf, offs, rank: int, world: int):
    """This rank's contiguous shard of a packed batch: (first index, buffer slice, offsets rebased to 0).

    ``buf`` / ``offs`` are the concatenated bytes and the n + 1 uint64 offsets of the whole (fixed) set; the shard's
    results land at pair indices ``[first * n_profiled, ...)`` of the whole job."""
    import numpy as np
    n = len(offs) - 1
    lo, hi = shard_range(n, rank, world)
    b0, b1 = int(offs[lo]), int(offs[hi])
    sbuf = np.ascontiguousarray(buf[b0:b1]) if b1 > b0 else np.zeros(1, dtype=np.uint8)
    soffs = (np.asarray(offs[lo:hi + 1]) - offs[lo]).astype(np.uint64)
    return lo, sbuf, soffs


def sum_u64_over_ranks(x: int, device=None) -> int:
    """Exact sum (mod 2^64) of per-rank 64-bit checksums: two 32-bit halves through an int64 all-reduce."""
    import torch
    import torch.distributed as dist
    x = int(x) & 0xFFFFFFFFFFFFFFFF
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x & 0xFFFFFFFF, x >> 32], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (int(t[0].item()) + (int(t[1].item()) << 32)) & 0xFFFFFFFFFFFFFFFF


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
