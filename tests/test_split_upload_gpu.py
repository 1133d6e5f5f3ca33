"""GPU: the split upload of the end-to-end calls (first part of the batch, first launch of the call's first kernel on it
while the rest arrives, second launch on the remaining tasks; zoe_cuda.cu stage_device / first_part_tasks) must give the
results of the single upload, entry point by entry point, and the oracle's on a sample.  The library splits shards of
64 MB and more; the tests lower that to 4 MB (ZOE_CUDA_SPLIT_UPLOAD_MIN_MB) so a 10-MB ragged batch is split, with the
split point inside the task pairing."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import CudaProfiles, WeightMatrix, synth

pytestmark = pytest.mark.gpu
W = WeightMatrix.new_dna_matrix(2, -5, b"N")


@pytest.fixture(autouse=True)
def _small_split_threshold():
    old = os.environ.get("ZOE_CUDA_SPLIT_UPLOAD_MIN_MB")
    os.environ["ZOE_CUDA_SPLIT_UPLOAD_MIN_MB"] = "4"
    yield
    if old is None:
        os.environ.pop("ZOE_CUDA_SPLIT_UPLOAD_MIN_MB", None)
    else:
        os.environ["ZOE_CUDA_SPLIT_UPLOAD_MIN_MB"] = old


def _batch(n=80_000, seed=11):
    rng = np.random.default_rng(seed)
    targets = [synth.random_dna(rng, L) for L in (900, 433)]
    reads = synth.illumina_reads(rng, targets, n)
    # ragged lengths (the split point is an even sequence index, not a byte position)
    cut = rng.integers(100, 151, size=n)
    seqs = [np.asarray(r[: int(c)], dtype=np.uint8) for r, c in zip(reads, cut)]
    buf, offs = synth.pack(seqs)
    return targets, seqs, buf, offs


def _with_env(flag, fn):
    old = os.environ.pop("ZOE_CUDA_NO_SPLIT_UPLOAD", None)
    try:
        if flag:
            os.environ["ZOE_CUDA_NO_SPLIT_UPLOAD"] = "1"
        return fn()
    finally:
        os.environ.pop("ZOE_CUDA_NO_SPLIT_UPLOAD", None)
        if old is not None:
            os.environ["ZOE_CUDA_NO_SPLIT_UPLOAD"] = old


def test_split_upload_matches_single_upload_and_oracle():
    targets, seqs, buf, offs = _batch()
    assert buf.nbytes >= (4 << 20) and len(seqs) >= 4096
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W, -10, -1, profiled_is_query=False)
    try:
        runs = {}
        for single in (False, True):
            runs[single] = _with_env(single, lambda: (
                prof.sw_score_arrays(buf, offs), prof.last_timing()["kernel_launches"],
                prof.align_arrays(buf, offs), prof.ranges_arrays(buf, offs), prof.align_arrays(buf, offs, three_pass=True)))
        (s0, l0, a0, r0, t0), (s1, l1, a1, r1, t1) = runs[False], runs[True]
        assert l0 == l1 + 1, "the split call launches its first kernel twice"
        for x, y in zip(s0, s1):
            assert np.array_equal(x, y)
        for d0, d1 in ((a0, a1), (r0, r1), (t0, t1)):
            for k in d0:
                if k == "cigar":
                    nw = int(d0["cigar_off"][-1])
                    assert nw == int(d1["cigar_off"][-1]) and np.array_equal(d0[k][:nw], d1[k][:nw])
                else:
                    assert np.array_equal(d0[k], d1[k]), k
        # the oracle on sequences either side of the split point (1/8 of the batch, at least one grid trip) and at the ends
        sc = O.Scoring(W.weights, W.mapping.index_map, -10, -1)
        n = len(seqs)
        probe = sorted(set([0, 1, n // 8 - 2, n // 8 - 1, n // 8, n // 8 + 1, 28414, 28415, 28416, 28417, n // 2, n - 2, n - 1]))
        score = s0[0]
        for i in probe:
            for j, t in enumerate(targets):
                rc, want, _ = O.sw_score_from(bytes(t), bytes(seqs[i]), sc)
                if rc == O.SOME:
                    assert int(score[i, j]) == want, (i, j)
                rc, wa, _ = O.sw_align_from(bytes(t), bytes(seqs[i]), sc, streamed_is_query=True)
                p = i * len(targets) + j
                if rc == O.SOME:
                    assert (int(a0["score"][p]), int(a0["ref_start"][p]), int(a0["ref_end"][p]), int(a0["query_start"][p]),
                            int(a0["query_end"][p])) == (wa.score, wa.ref_range[0], wa.ref_range[1], wa.query_range[0],
                                                         wa.query_range[1]), (i, j)
    finally:
        prof.close()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2")
def test_split_upload_two_devices_in_one_process():
    """Each device's shard is large enough to be split: the per-device copy streams, events and second launches must
    not depend on one another (zoe_cuda_create(n_devices = 2), one host thread per device)."""
    targets, seqs, buf, offs = _batch(n=170_000, seed=12)
    res = []
    for nd in (1, 2):
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W, -10, -1, profiled_is_query=False, n_devices=nd)
        try:
            res.append((prof.sw_score_arrays(buf, offs), prof.align_arrays(buf, offs), prof.ranges_arrays(buf, offs)))
        finally:
            prof.close()
    (s1, a1, r1), (s2, a2, r2) = res
    for x, y in zip(s1, s2):
        assert np.array_equal(x, y)
    nw = int(a1["cigar_off"][-1])
    assert nw == int(a2["cigar_off"][-1]) and np.array_equal(a1["cigar"][:nw], a2["cigar"][:nw])
    for k in a1:
        if k != "cigar":
            assert np.array_equal(a1[k], a2[k]), k
    for k in r1:
        assert np.array_equal(r1[k], r2[k]), k
