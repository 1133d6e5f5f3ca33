"""CPU-only, world_size 2 over gloo: the sharding rule and the cross-rank reductions bench.py relies on."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from zoe_b200.dist import shard_range


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from zoe_b200 import dist as zd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    a, b = zd.shard_range(n, rank, world)
    # each rank "scores" its shard: a stand-in result that depends only on the global index
    local = (np.arange(a, b, dtype=np.int64) * 7 + 3) % 311
    ms = zd.max_over_ranks(10.0 + rank)
    cells = zd.sum_over_ranks(float(b - a))
    dist.barrier()
    parts = zd.gather_results(local, rank, world)
    if rank == 0:
        full = np.concatenate(parts)
        q.put((ms, cells, bool(np.array_equal(full, (np.arange(n, dtype=np.int64) * 7 + 3) % 311))))
    dist.destroy_process_group()


def test_two_rank_gloo_reductions_and_ordering():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, cells, ordered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 11.0 and cells == 1001.0 and ordered
