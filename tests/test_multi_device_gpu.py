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
        prof.close()
        res.append((score, aln, rng, st))
    (s1, a1, r1, st1), (s2, a2, r2, st2) = res
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
