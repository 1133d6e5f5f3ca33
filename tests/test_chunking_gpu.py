"""The align / ranges / 3-pass pipelines process a batch in chunks when their scratch (checkpoints, direction-bit windows,
box scratch) exceeds the memory budget.  The benchmark configs fit one chunk on a 180 GB part, so the multi-chunk paths are
driven here with tiny budgets (zoe_cuda_set_memory_budget): every output must be identical to the single-chunk run."""
import os

import numpy as np
import pytest

from zoe_b200 import CudaProfiles, WeightMatrix, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
KEYS = ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end")


def _same(a, b, n_pairs, with_cigar):
    for k in KEYS:
        assert np.array_equal(a[k][:n_pairs], b[k][:n_pairs]), k
    if with_cigar:
        assert np.array_equal(a["cigar_off"][:n_pairs + 1], b["cigar_off"][:n_pairs + 1])
        nw = int(a["cigar_off"][n_pairs])
        assert np.array_equal(a["cigar"][:nw], b["cigar"][:nw])


@pytest.mark.parametrize("budget", [1 << 20, 5 << 20, 40 << 20])
def test_chunked_equals_unchunked_config3(budget):
    targets, reads = synth.config3(ROOT, n_reads=3001, seed=17)
    buf, offs = synth.fixed_len_batch(reads)
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
    n_pairs = reads.shape[0]
    ref_align = prof.align_arrays(buf, offs)
    ref_3p = prof.align_arrays(buf, offs, three_pass=True)
    ref_rng = prof.ranges_arrays(buf, offs)
    prof.set_memory_budget(budget)
    _same(prof.align_arrays(buf, offs), ref_align, n_pairs, True)
    _same(prof.align_arrays(buf, offs, three_pass=True), ref_3p, n_pairs, True)
    _same(prof.ranges_arrays(buf, offs), ref_rng, n_pairs, False)
    for mode in (CudaProfiles.ALIGN_FULL, CudaProfiles.ALIGN_WINDOW):
        prof.set_align_options(mode)
        _same(prof.align_arrays(buf, offs), ref_align, n_pairs, True)
    prof.close()


def test_chunked_panel_of_references():
    # several profiled sequences: pair index = read * n_profiled + j must survive chunking
    rng = np.random.default_rng(23)
    targets = [synth.random_dna(rng, int(L)) for L in (700, 1500, 333, 2048, 90)]
    reads = []
    for _ in range(801):
        t = targets[int(rng.integers(0, len(targets)))]
        k = min(120, len(t) - 2)
        st = int(rng.integers(0, len(t) - k + 1))
        frag = synth._mutate(rng, t[st:st + k], 0.04, 0.02, 0.02, np.frombuffer(b"ACGT", dtype=np.uint8))
        reads.append(frag[:140])
    buf = np.concatenate(reads)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
    n_pairs = len(reads) * len(targets)
    ref_align = prof.align_arrays(buf, offs)
    ref_3p = prof.align_arrays(buf, offs, three_pass=True)
    ref_rng = prof.ranges_arrays(buf, offs)
    prof.set_memory_budget(3 << 20)
    _same(prof.align_arrays(buf, offs), ref_align, n_pairs, True)
    _same(prof.align_arrays(buf, offs, three_pass=True), ref_3p, n_pairs, True)
    _same(prof.ranges_arrays(buf, offs), ref_rng, n_pairs, False)
    prof.close()
