"""Several B200s driven from one process (zoe_cuda_create(n_devices)): results must be independent of the device
count -- contiguous streamed-index shards, no collective.  Skipped on a single-GPU box."""
import os

import numpy as np
import pytest

from zoe_b200 import CudaProfiles, WeightMatrix, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_devices_equal_one_device_score_align_ranges():
    targets, reads = synth.config3(ROOT, n_reads=20_001, seed=5)  # odd count: uneven shards
    buf, offs = synth.fixed_len_batch(reads)
    res = []
    for nd in (1, 2):
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1, n_devices=nd)
        score = prof.sw_score_arrays(buf, offs)
        aln = prof.align_arrays(buf, offs)
        rng = prof.ranges_arrays(buf, offs)
        st = prof.last_stats()
        tp = prof.align_arrays(buf, offs, three_pass=True)
        st_tp = prof.last_stats()
        prof.close()
        res.append((score, aln, rng, st, tp, st_tp))
    (s1, a1, r1, st1, t1, stt1), (s2, a2, r2, st2, t2, stt2) = res
    n3 = int(t1["cigar_off"][-1])
    assert n3 == int(t2["cigar_off"][-1])
    for k in ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(t1[k], t2[k]), ("3pass", k)
    assert np.array_equal(t1["cigar"][:n3], t2["cigar"][:n3])
    assert (stt1["tp_nogaps"], stt1["tp_banded"], stt1["tp_scalar"]) == (stt2["tp_nogaps"], stt2["tp_banded"], stt2["tp_scalar"])
    for x, y in zip(s1, s2):
        assert np.array_equal(x, y)
    n_words = int(a1["cigar_off"][-1])
    assert n_words == int(a2["cigar_off"][-1])
    for k in ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(a1[k], a2[k]), k
    assert np.array_equal(a1["cigar"][:n_words], a2["cigar"][:n_words])
    for k in r1:
        assert np.array_equal(r1[k], r2[k]), k
    assert st1["tier8"] == st2["tier8"] and st1["tier16"] == st2["tier16"]


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_devices_sneaky_snake():
    from oracle import oracle as O
    from zoe_b200 import SneakySnake
    rng = np.random.default_rng(12)
    refs, qs = [], []
    for _ in range(1001):
        ref = rng.choice(list(b"ACGT"), 100).astype(np.uint8)
        q = ref.copy()
        for pos in rng.integers(0, 100, int(rng.integers(0, 20))):
            q[pos] = rng.choice(list(b"ACGT"))
        refs.append(bytes(ref))
        qs.append(bytes(q))
    snake = SneakySnake(n_devices=2)
    assert snake.sneaky_snake_batch(refs, qs, 0.1) == [O.sneaky_snake(r, q, 0.1) for r, q in zip(refs, qs)]
    assert snake.sneaky_snake_batch(refs[:1], qs[:1], 0.1) == [O.sneaky_snake(refs[0], qs[0], 0.1)]  # fewer pairs than devices
    snake.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_devices_pipelined_score_batch():
    # sub-batches alternate between the two streams of each GPU: same scores as one device, one shot
    rng = np.random.default_rng(3)
    target = synth.random_dna(rng, 150)
    n = 540_001
    reads = synth.random_dna(rng, n * 64).reshape(n, 64)
    own = rng.random(n) < 0.5
    idx = rng.integers(0, 80, n)[:, None] + np.arange(64)[None, :]
    reads[own] = target[idx[own]]
    buf, offs = synth.fixed_len_batch(reads.astype(np.uint8))
    out = []
    for nd, env in ((1, "ZOE_CUDA_NO_PIPELINE"), (2, "ZOE_CUDA_PIPELINE")):
        prof = CudaProfiles.new_with_w256([bytes(target)], W25, -10, -1, n_devices=nd)
        os.environ[env] = "1"
        try:
            out.append((prof.sw_score_arrays(buf, offs), prof.last_stats()))
        finally:
            del os.environ[env]
        prof.close()
    for x, y in zip(out[0][0], out[1][0]):
        assert np.array_equal(x, y)
    assert out[0][1] == out[1][1]
