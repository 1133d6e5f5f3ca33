"""CPU-only, world_size 2 over gloo: the sharding rule, the shard slicing and the cross-rank reductions bench.py
relies on -- exercised on real Smith-Waterman results (the CPU port scores each rank's shard; the GPU is not needed
to show that results and checksums do not depend on how the reads are split)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from zoe_b200.dist import shard_batch, shard_range


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_slices_ragged_batches():
    rng = np.random.default_rng(0)
    seqs = [rng.integers(65, 70, int(k), dtype=np.uint8) for k in rng.integers(0, 9, 23)]
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.concatenate(seqs)
    for world in (1, 2, 3, 8, 40):
        seen = 0
        for r in range(world):
            first, sbuf, soffs = shard_batch(buf, offs, r, world)
            assert first == seen and soffs[0] == 0
            for i in range(len(soffs) - 1):
                assert np.array_equal(sbuf[int(soffs[i]):int(soffs[i + 1])], seqs[first + i])
            seen += len(soffs) - 1
        assert seen == len(seqs)


def _workload():
    from zoe_b200 import WeightMatrix, synth
    wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
    targets, reads = synth.config2(n_reads=301, seed_reads=3)
    buf, offs = synth.fixed_len_batch(reads)
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    return wm, pbuf, poff, buf, offs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import cpu_baseline as CB
    from zoe_b200 import dist as zd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wm, pbuf, poff, buf, offs = _workload()
    first, sbuf, soffs = zd.shard_batch(buf, offs, rank, world)
    score, status, tier = CB.score_batch(pbuf, poff, sbuf, soffs, wm.weights, wm.mapping.index_map, -10, -1, n_threads=1)
    checksum = zd.sum_u64_over_ranks(int(score.sum(dtype=np.uint64)) + (1 << 40) * 3)  # exercises the high half too
    ms = zd.max_over_ranks(10.0 + rank)
    cells = zd.sum_over_ranks(float(int(soffs[-1]) * int(poff[-1])))
    dist.barrier()
    parts = zd.gather_results((first, score), rank, world)
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        q.put((ms, cells, checksum, np.concatenate([p[1] for p in parts])))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_scores_equal_unsharded():
    from oracle import cpu_baseline as CB
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, cells, checksum, sharded = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wm, pbuf, poff, buf, offs = _workload()
    full, _, _ = CB.score_batch(pbuf, poff, buf, offs, wm.weights, wm.mapping.index_map, -10, -1, n_threads=2)
    assert ms == 11.0
    assert cells == float(int(offs[-1]) * int(poff[-1]))
    assert np.array_equal(sharded, full)
    assert checksum == (int(full.sum(dtype=np.uint64)) + 2 * 3 * (1 << 40)) & 0xFFFFFFFFFFFFFFFF
