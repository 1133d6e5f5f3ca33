"""GPU parity of the literal striped emulation kernels (sw_exact_fast_kernel + walk, sw_align_exact_kernel) on workloads
where many pairs need them: small weights and cheap gaps (E == H == F ties on the walk), gap_open == 0 (every pair),
packed overflow.  Checked pair by pair against the vectorised CPU port of sw_simd_align (itself pinned to the plain-C
oracle in tests/test_cpu_baseline.py), through the C ABI."""
import os

import numpy as np
import pytest

from oracle import cpu_baseline as CB
from oracle import oracle as O
from zoe_b200 import BLOSUM_62, CudaProfiles, DNA_PROFILE_MAP, SeqSrc, WeightMatrix, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")
WIDTH = {(16, 8, 4): 128, (32, 16, 8): 256, (64, 32, 16): 512}


def run_and_compare(targets, reads, wm, go, ge, lanes=(32, 16, 8), env=None, align_options=None):
    """(stats, mismatches) of zoe_cuda_sw_align_batch against the CPU port on a packed batch."""
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    buf, offs = synth.pack([np.asarray(r, dtype=np.uint8) for r in reads])
    want = CB.align_batch(pbuf, poff, buf, offs, wm.weights, wm.mapping.index_map, go, ge, width_bits=WIDTH[lanes],
                          streamed_is_query=True)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        prof = CudaProfiles([bytes(t) for t in targets], wm, go, ge, lanes=lanes)
        if align_options:
            prof.set_align_options(*align_options)
        got = prof.align_arrays(buf, offs)
        stats = prof.last_stats()
        prof.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return stats, CB.compare_alignments(got, want, (len(offs) - 1) * len(targets)), got


def hazard_reads(n, seed, targets):
    rng = np.random.default_rng(seed)
    return list(synth.illumina_reads(rng, targets, n, sub=0.03, indel=0.004, random_frac=0.2))


@pytest.mark.parametrize("go,ge", [(-3, -1), (-2, -2), (-4, -1)])
@pytest.mark.parametrize("lanes", [(32, 16, 8), (16, 8, 4), (64, 32, 16)])
def test_hazard_heavy_reads_window_pipeline(go, ge, lanes):
    ha = synth.golden_ha(ROOT)
    reads = hazard_reads(1500, 11, [ha])
    stats, mism, _ = run_and_compare([ha], reads, W42, go, ge, lanes=lanes)
    assert mism == 0, stats
    assert stats["hazard"] > 15, stats  # the workload does exercise the literal path


@pytest.mark.parametrize("lanes", [(32, 16, 8), (16, 8, 4)])
def test_cta_per_pair_kernel_on_every_hazard_pair(lanes):
    # the latency kernel (one CTA per pair, F by a max-plus scan) forced onto whole hazard lists
    ha = synth.golden_ha(ROOT)
    reads = hazard_reads(700, 19, [ha])
    for go, ge in ((-3, -1), (-2, -2), (-10, -1)):
        stats, mism, _ = run_and_compare([ha], reads, W42, go, ge, lanes=lanes, env={"ZOE_CUDA_EXACT_CTA": "1"})
        assert mism == 0, (go, ge, stats)
    rng = np.random.default_rng(20)
    targets = [synth.random_dna(rng, 700), synth.random_dna(rng, 333), synth.random_dna(rng, 64)]
    stats, mism, _ = run_and_compare(targets, hazard_reads(300, 21, targets[:2]), W42, 0, 0, lanes=lanes,
                                     env={"ZOE_CUDA_EXACT_CTA": "1"})
    assert mism == 0, stats


def test_fast_and_slow_literal_kernels_agree():
    ha = synth.golden_ha(ROOT)
    reads = hazard_reads(800, 12, [ha])
    s1, m1, a = run_and_compare([ha], reads, W42, -3, -1)
    s2, m2, b = run_and_compare([ha], reads, W42, -3, -1, env={"ZOE_CUDA_EXACT_SLOW": "1"})
    assert m1 == 0 and m2 == 0 and s1["hazard"] == s2["hazard"] > 0
    assert CB.compare_alignments(a, b, len(reads)) == 0


def test_every_pair_literal_gap_open_zero_and_all_exact_env():
    rng = np.random.default_rng(13)
    targets = [synth.random_dna(rng, 700), synth.random_dna(rng, 333), synth.random_dna(rng, 64)]
    reads = hazard_reads(300, 14, targets[:2]) + [synth.random_dna(rng, int(k)) for k in rng.integers(1, 200, 60)]
    for go, ge in ((0, 0),):
        stats, mism, _ = run_and_compare(targets, reads, W42, go, ge)
        assert mism == 0, (go, ge, stats)
    # the window pipeline with every mapped pair forced through the literal kernels: end cells unknown to them
    ha = synth.golden_ha(ROOT)
    reads = hazard_reads(400, 15, [ha])
    stats, mism, _ = run_and_compare([ha], reads, W42, -3, -1, env={"ZOE_CUDA_ALL_EXACT": "1"}, align_options=(2, 6, 16))
    assert mism == 0, stats


def test_literal_kernels_multi_profiled_ragged_protein():
    rng = np.random.default_rng(16)
    tp, q = synth.config5(n_queries=200)
    t2 = synth._AA20[rng.integers(0, 20, 411)]
    # BLOSUM62 with cheap gaps: many E == H == F ties
    stats, mism, _ = run_and_compare([tp[0], t2], list(q), BLOSUM_62, -2, -1)
    assert mism == 0, stats
    assert stats["hazard"] > 0


def test_packed_overflow_pairs_take_the_literal_path():
    # match 127: scores beyond the packed 16-bit lanes -> exact 32-bit scores, literal kernels (i16 and i32 tiers)
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    rng = np.random.default_rng(17)
    t = synth.random_dna(rng, 900)
    reads = [t[:n].copy() for n in (2, 100, 255, 258, 300, 516, 600, 899)] + [synth.random_dna(rng, 300)]
    stats, mism, got = run_and_compare([t], reads, w, -10, -1)
    assert mism == 0, stats
    assert stats["tier32"] > 0 and stats["tier16"] > 0


def test_forced_literal_pairs_against_a_100kb_profiled_sequence():
    # the literal kernel's scratch must respect the budget when the profiled sequence is long
    rng = np.random.default_rng(18)
    genome = synth.random_dna(rng, 120_000)
    reads = [genome[s:s + 150].copy() for s in (0, 5000, 60_000, 119_850)] + [synth.random_dna(rng, 150)]
    for r in reads[:4]:
        r[rng.integers(0, 150, 4)] = synth.random_dna(rng, 4)
    sc = O.Scoring(W42.weights, W42.mapping.index_map, -3, -1)
    prof = CudaProfiles([bytes(genome)], W42, -3, -1)
    prof.set_memory_budget(1 << 30)
    os.environ["ZOE_CUDA_ALL_EXACT"] = "1"
    try:
        got = prof.sw_align_batch(SeqSrc.Query([bytes(r) for r in reads]))
    finally:
        os.environ.pop("ZOE_CUDA_ALL_EXACT", None)
    prof.close()
    for i, r in enumerate(reads):
        rc, want, _ = O.sw_align_from(bytes(genome), bytes(r), sc, streamed_is_query=True)
        assert got[i][0].status.value == rc
        if rc == 0:
            a = got[i][0].unwrap()
            assert (a.score, a.ref_range, a.query_range, a.states) == (want.score, want.ref_range, want.query_range, want.cigar)
