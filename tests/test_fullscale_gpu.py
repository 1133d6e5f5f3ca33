"""GPU parity at BASELINE.json's full sizes, through size-independent properties: two independently written CUDA
paths for the same zoe function must agree on every pair (a checksum of checksums is not enough: arrays are
compared element-wise), and a seeded sample of each is checked against the CPU restatement."""
import os

import numpy as np
import pytest

from oracle import cpu_baseline as CB
from zoe_b200 import BLOSUM_62, CudaProfiles, WeightMatrix, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")


def _scores(targets, wm, buf, offs, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
        out = prof.sw_score_arrays(buf, offs)
        stats = prof.last_stats()
        prof.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out, stats


def _cpu_sample(targets, wm, buf, offs, n_sample):
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    o = offs[: n_sample + 1]
    return CB.score_batch(pbuf, poff, buf[: int(o[-1])], o, wm.weights, wm.mapping.index_map, -10, -1, width_bits=256,
                          n_threads=CB.hardware_threads())


def _assert_matches_cpu(gpu, cpu, n_sample):
    (score, status, tier), (c_score, c_status, c_tier) = gpu, cpu
    assert np.array_equal(status[:n_sample], c_status)
    some = c_status == 0
    assert np.array_equal(score[:n_sample][some], c_score[some])
    assert np.array_equal(tier[:n_sample][some], c_tier[some])


def test_cfg2_full_size_two_kernels_agree_and_match_cpu_sample():
    targets, reads = synth.config2(n_reads=1_000_000)
    buf, offs = synth.fixed_len_batch(reads)
    two, stats = _scores(targets, W25, buf, offs)                                 # two column streams per sweep
    one, _ = _scores(targets, W25, buf, offs, {"ZOE_CUDA_ONE_STREAM": "1"})       # single-stream instantiation
    for a, b in zip(two, one):
        assert np.array_equal(a, b)
    assert stats["tier8"] + stats["tier16"] + stats["unmapped"] == 8_000_000 and stats["tier16"] > 400_000
    _assert_matches_cpu(two, _cpu_sample(targets, W25, buf, offs, 20_000), 20_000)


def test_cfg3_full_size_window_pipeline_equals_full_matrix_pipeline():
    targets, reads = synth.config3(ROOT, n_reads=1_000_000)
    buf, offs = synth.fixed_len_batch(reads)
    outs = []
    for mode in (CudaProfiles.ALIGN_WINDOW, CudaProfiles.ALIGN_FULL):
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
        prof.set_align_options(mode, 7, 16)
        outs.append(prof.align_arrays(buf, offs, cigar_cap=8 * 1_000_000 + 1024))
        prof.close()
    a, b = outs
    n_words = int(a["cigar_off"][-1])
    assert n_words == int(b["cigar_off"][-1]) and n_words > 1_000_000
    for k in ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cigar"][:n_words], b["cigar"][:n_words])
    # every CIGAR consumes exactly its read (query side) and its reference span
    ops, lens = a["cigar"][:n_words] & 15, a["cigar"][:n_words] >> 4
    pair = np.repeat(np.arange(1_000_000), np.diff(a["cigar_off"]).astype(np.int64))
    q_consumed = np.bincount(pair, weights=lens * np.isin(ops, (0, 1, 4)), minlength=1_000_000)
    r_consumed = np.bincount(pair, weights=lens * np.isin(ops, (0, 2)), minlength=1_000_000)
    some = a["status"] == 0
    assert np.array_equal(q_consumed[some], np.full(int(some.sum()), 150.0))
    assert np.array_equal(r_consumed[some], (a["ref_end"][some] - a["ref_start"][some]).astype(np.float64))


def test_cfg5_full_size_transposed_kernel_equals_per_task_table_kernel():
    targets, q = synth.config5(n_queries=1_000_000)
    buf, offs = synth.fixed_len_batch(q)
    rows, stats = _scores(targets, BLOSUM_62, buf, offs)
    per_task, _ = _scores(targets, BLOSUM_62, buf, offs, {"ZOE_CUDA_NO_ROWS_KERNEL": "1"})
    for a, b in zip(rows, per_task):
        assert np.array_equal(a, b)
    assert stats["tier16"] > 500_000 and stats["tier8"] > 300_000
    _assert_matches_cpu(rows, _cpu_sample(targets, BLOSUM_62, buf, offs, 20_000), 20_000)


def test_cfg4_slice_long_rows_match_cpu():
    targets, reads = synth.config4(n_reads=2_000)
    buf, offs = synth.pack(reads)
    gpu, stats = _scores(targets, W25, buf, offs)
    assert stats["tier16"] > 500
    _assert_matches_cpu(gpu, _cpu_sample(targets, W25, buf, offs, 400), 400)
