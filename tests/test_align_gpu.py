"""GPU parity: zoe_cuda_sw_align_batch (through the C ABI) vs the CPU oracle -- scores, ranges and CIGARs
including tie-breaking, bit-exact, in both SeqSrc orientations."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import BLOSUM_62, CudaProfiles, DNA_PROFILE_MAP, SeqSrc, WeightMatrix, synth
from zoe_b200.alignment import Status
from test_oracle_golden import ALN_CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")


def osc(wm, go, ge):
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


def check_align(targets, seqs, wm, go=-10, ge=-1, profiled_is_query=False, lanes=(32, 16, 8), expect_hazard=None,
                align_options=None):
    """Compare every (seq, target) pair with ProfileSets::sw_align_from_i8 of the oracle."""
    targets = [bytes(t) for t in targets]
    seqs = [bytes(s) for s in seqs]
    prof = CudaProfiles(targets, wm, go, ge, lanes=lanes, profiled_is_query=profiled_is_query)
    if align_options is not None:
        prof.set_align_options(*align_options)
    src = SeqSrc.Reference(seqs) if profiled_is_query else SeqSrc.Query(seqs)
    got = prof.sw_align_batch(src)
    stats = prof.last_stats()
    sc = osc(wm, go, ge)
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            rc, want, tier = O.sw_align_from(t, s, sc, lanes=lanes, streamed_is_query=not profiled_is_query)
            g = got[i][j]
            assert g.status.value == rc, (i, j, g, rc)
            if rc == O.SOME:
                a = g.unwrap()
                assert (a.score, a.ref_range, a.query_range, a.states, a.ref_len, a.query_len) == \
                       (want.score, want.ref_range, want.query_range, want.cigar, want.ref_len, want.query_len), \
                       (i, j, a, want, tier)
    prof.close()
    if expect_hazard is not None:
        assert (stats["hazard"] > 0) == expect_hazard, stats
    return stats


def test_doc_examples_both_orientations():
    # doc striped.rs:418-441 (6M2D9M3S), sw/mod.rs:164-188 (5M1D4M), types/test.rs:14-38 (+ inverted)
    prof = CudaProfiles.new_with_w256([b"CGTTCGCCATAAAGGGGG", b"CTCAGATTG", b"TCTCAGATTGCAGTTT"], W42, -3, -1,
                                      profiled_is_query=True)
    r = prof.sw_align_batch(SeqSrc.Reference([b"ATGCATCGATCGATCGATCGATCGATCGATGC", b"GGCCACAGGATTGAG", b"GGCCACAGGATTGAGC"]))
    a = r[0][0].unwrap()
    assert (a.states, a.score, a.query_range, a.ref_range) == ("6M2D9M3S", 26, (0, 15), (14, 31))
    a = r[1][1].unwrap()
    assert (a.states, a.score, a.ref_range[0]) == ("5M1D4M", 27, 3)
    a = r[2][2].unwrap()
    assert (a.states, a.ref_range, a.query_range) == ("1S5M1D4M1I2M3S", (3, 15), (1, 13))
    prof.close()
    prof = CudaProfiles.new_with_w256([b"TCTCAGATTGCAGTTT"], W42, -3, -1, profiled_is_query=False)
    a = prof.sw_align_batch(SeqSrc.Query([b"GGCCACAGGATTGAGC"]))[0][0].unwrap()
    # the profile is the reference here: same pair as above seen from the other side
    want = O.scalar_align(b"TCTCAGATTGCAGTTT", b"GGCCACAGGATTGAGC", osc(W42, -3, -1), streamed_is_query=True)[1]
    assert (a.states, a.ref_range, a.query_range) == (want.cigar, want.ref_range, want.query_range)
    assert a.states == "3S5M1I4M1D2M1S"
    prof.close()


def test_zoe_small_alignment_cases():
    # sw/test.rs:121-195 -- profile = query, streamed = reference
    for query, reference, lanes, cigar in ALN_CASES:
        prof = CudaProfiles([query], W25, -10, -1, lanes=(lanes, lanes, lanes), profiled_is_query=True)
        a = prof.sw_align_batch(SeqSrc.Reference([reference]))[0][0].unwrap()
        assert a.states == cigar, (query, reference, a)
        prof.close()


def test_h5_h1_golden(seqs):
    # sw/test.rs:116-119 (336S29M1395S) and :71-79
    prof = CudaProfiles([seqs["H5_HA"]], W25, -10, -1, lanes=(8, 8, 8), profiled_is_query=True)
    a = prof.sw_align_batch(SeqSrc.Reference([seqs["H1_HA"]]))[0][0].unwrap()
    assert (a.states, a.score) == ("336S29M1395S", 37)
    prof.close()


def test_config3_sample_vs_oracle():
    targets, reads = synth.config3(ROOT, n_reads=400)
    stats = check_align(targets, list(reads), W25)
    assert stats["tier16"] > 100 and stats["tier8"] > 100


def test_random_pairs_many_scorings_default_orientation():
    rng = np.random.default_rng(17)
    n_hazard = 0
    for (ma, mi, go, ge) in [(2, -5, -10, -1), (4, -2, -3, -1), (1, -1, -4, -2), (3, -1, -4, -1), (2, -5, -10, -10),
                             (1, -1, -1, -1), (5, -4, -2, 0), (6, -2, -1, -1)]:
        wm = WeightMatrix.new_dna_matrix(ma, mi, b"N")
        targets = [synth.random_dna(rng, int(L)) for L in (37, 120, 64)]
        seqs = []
        for _ in range(60):
            L = int(rng.integers(4, 150))
            s = synth.random_dna(rng, L)
            if L > 24:
                t = targets[int(rng.integers(0, 3))]
                k = min(L - 4, len(t) - 2, 60)
                st = int(rng.integers(0, len(t) - k + 1))
                frag = synth._mutate(rng, t[st:st + k], 0.06, 0.05, 0.05, np.frombuffer(b"ACGT", dtype=np.uint8))
                k2 = min(len(frag), L - 2)
                s[2:2 + k2] = frag[:k2]
            seqs.append(s)
        stats = check_align(targets, seqs, wm, go, ge)
        n_hazard += stats["hazard"]
    assert n_hazard > 0  # the literal striped-emulation kernel was exercised


def test_random_pairs_profiled_is_query_and_lane_presets():
    rng = np.random.default_rng(23)
    wm = WeightMatrix.new_dna_matrix(4, -2, b"N")
    for lanes in [(16, 8, 4), (32, 16, 8), (64, 32, 16), (4, 4, 4)]:
        targets = [synth.random_dna(rng, int(L)) for L in (50, 90)]
        seqs = []
        for _ in range(40):
            L = int(rng.integers(10, 130))
            s = synth.random_dna(rng, L)
            t = targets[int(rng.integers(0, 2))]
            k = min(L, len(t)) - 4
            frag = synth._mutate(rng, t[:k], 0.08, 0.06, 0.06, np.frombuffer(b"AC", dtype=np.uint8))
            k2 = min(len(frag), L - 2)
            s[1:1 + k2] = frag[:k2]
            seqs.append(s)
        check_align(targets, seqs, wm, -3, -1, profiled_is_query=True, lanes=lanes)


def test_gap_open_zero_goes_through_literal_kernel():
    rng = np.random.default_rng(29)
    wm = WeightMatrix.new_dna_matrix(2, -3, b"N")
    targets = [synth.random_dna(rng, 60)]
    seqs = [synth.random_dna(rng, int(L)) for L in rng.integers(5, 80, 30)]
    for s in seqs[:15]:
        k = min(len(s), 40)
        s[:k] = targets[0][5:5 + k]
    stats = check_align(targets, seqs, wm, 0, 0)
    assert stats["hazard"] > 0


def test_protein_alignments_vs_oracle():
    targets, q = synth.config5(n_queries=40)
    check_align(targets, list(q), BLOSUM_62)


def test_align_edge_cases():
    wm = W25
    prof = CudaProfiles.new_with_w256([b"ACGTACGTAC", b"A"], wm, -10, -1)
    r = prof.sw_align_batch(SeqSrc.Query([b"", b"C", b"TTTT", b"ACGTACGTAC", b"A"]))
    assert r[0][0].status is Status.Unmapped and r[0][1].status is Status.Unmapped  # empty streamed sequence
    assert r[2][0].unwrap().score == 2                                               # single-base match
    assert r[2][1].status is Status.Unmapped
    a = r[3][0].unwrap()
    assert (a.states, a.score, a.ref_range, a.query_range) == ("10M", 20, (0, 10), (0, 10))
    assert r[4][1].unwrap().states == "1M"
    prof.close()
    # scores beyond the packed 16-bit range: exact 32-bit score + literal kernel (i32 tier)
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    stats = check_align([b"A" * 600], [b"A" * 600, b"A" * 300, b"A" * 100 + b"C" + b"A" * 100], w)
    assert stats["tier32"] == 1 and stats["rerun_wide"] >= 1


# ---- the checkpointed-window pipeline (sw_align_win.cuh) must give the same answers as the full-matrix one ----
WINDOW = CudaProfiles.ALIGN_WINDOW


def _related_seqs(rng, targets, n, lo=4, hi=150, alphabet=b"ACGT"):
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi))
        s = synth.random_dna(rng, L)
        if L > 24:
            t = targets[int(rng.integers(0, len(targets)))]
            k = min(L - 4, len(t) - 2, 120)
            st = int(rng.integers(0, len(t) - k + 1))
            frag = synth._mutate(rng, t[st:st + k], 0.06, 0.05, 0.05, np.frombuffer(alphabet, dtype=np.uint8))
            k2 = min(len(frag), L - 2)
            s[2:2 + k2] = frag[:k2]
        seqs.append(s)
    return seqs


@pytest.mark.parametrize("cb_log2,slack", [(2, 1), (3, 2), (4, 8), (5, 16), (7, 16)])
def test_window_pipeline_random_pairs_vs_oracle(cb_log2, slack):
    rng = np.random.default_rng(101 + cb_log2)
    fallbacks = 0
    for (ma, mi, go, ge) in [(2, -5, -10, -1), (4, -2, -3, -1), (1, -1, -4, -2), (3, -1, -4, -1), (1, -1, -1, -1),
                             (5, -4, -2, 0)]:
        wm = WeightMatrix.new_dna_matrix(ma, mi, b"N")
        targets = [synth.random_dna(rng, int(L)) for L in (37, 300, 64, 513)]
        seqs = _related_seqs(rng, targets, 70)
        stats = check_align(targets, seqs, wm, go, ge, align_options=(WINDOW, cb_log2, slack))
        fallbacks += stats["window_fallback"]
    if slack <= 2:
        assert fallbacks > 0  # walks longer than the slack were handed to the literal kernel


def test_window_pipeline_profiled_is_query_and_lane_presets():
    rng = np.random.default_rng(131)
    wm = WeightMatrix.new_dna_matrix(4, -2, b"N")
    for lanes in [(16, 8, 4), (32, 16, 8), (64, 32, 16)]:
        targets = [synth.random_dna(rng, int(L)) for L in (250, 90)]
        seqs = _related_seqs(rng, targets, 50, lo=10, hi=130, alphabet=b"AC")
        check_align(targets, seqs, wm, -3, -1, profiled_is_query=True, lanes=lanes, align_options=(WINDOW, 4, 6))


def test_window_pipeline_config3_equals_full_matrix_pipeline():
    """5000 cfg-3 reads: the two pipelines must agree on every output array, CIGAR words included."""
    targets, reads = synth.config3(ROOT, n_reads=5000)
    buf, offs = synth.fixed_len_batch(reads)
    outs = []
    for mode in (CudaProfiles.ALIGN_FULL, CudaProfiles.ALIGN_WINDOW):
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
        prof.set_align_options(mode, 7, 16)
        outs.append(prof.align_arrays(buf, offs))
        st = prof.last_stats()
        prof.close()
    a, b = outs
    n_words = int(a["cigar_off"][-1])
    assert n_words == int(b["cigar_off"][-1]) and n_words > 5000
    for k in ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cigar"][:n_words], b["cigar"][:n_words])
    assert st["window_fallback"] < 50  # the default slack covers all but the rare long-gap reads


def test_window_pipeline_config3_sample_vs_oracle():
    targets, reads = synth.config3(ROOT, n_reads=300, seed=77)
    check_align(targets, list(reads), W25, align_options=(WINDOW, 7, 16))


def test_window_pipeline_protein_and_overflow():
    targets, q = synth.config5(n_queries=30)
    check_align(targets, list(q), BLOSUM_62, align_options=(WINDOW, 5, 12))
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    stats = check_align([b"A" * 600], [b"A" * 600, b"A" * 300, b"A" * 100 + b"C" + b"A" * 100], w,
                        align_options=(WINDOW, 6, 8))
    assert stats["tier32"] == 1 and stats["rerun_wide"] >= 1


def test_window_pipeline_with_profiled_set_larger_than_shared_memory():
    """> 96 KB of profiled sequences: pass A reads the symbol codes from global memory (sw_align_scan_kernel<.., false>)."""
    rng = np.random.default_rng(211)
    targets = [synth.random_dna(rng, int(L)) for L in rng.integers(2300, 2900, 40)]
    assert sum(len(t) for t in targets) > 96 * 1024
    seqs = _related_seqs(rng, targets, 40, lo=60, hi=150)
    stats = check_align(targets, seqs, W25, align_options=(CudaProfiles.ALIGN_AUTO, 6, 16))
    assert stats["tier16"] + stats["tier8"] > 0
    # and the two pipelines agree on every array for a larger batch
    reads = _related_seqs(rng, targets, 600, lo=100, hi=150)
    buf, offs = synth.pack(reads)
    outs = []
    for mode in (CudaProfiles.ALIGN_WINDOW, CudaProfiles.ALIGN_FULL):
        prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
        prof.set_align_options(mode, 6, 16)
        outs.append(prof.align_arrays(buf, offs))
        prof.close()
    a, b = outs
    n_words = int(a["cigar_off"][-1])
    assert n_words == int(b["cigar_off"][-1])
    for k in ("score", "status", "tier", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cigar"][:n_words], b["cigar"][:n_words])
