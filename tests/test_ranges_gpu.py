"""GPU parity: zoe_cuda_sw_score_ranges_batch (through the C ABI) vs the CPU oracle's literal restatement of
sw_simd_score_ranges (striped.rs:355-388) with the ProfileSets escalation (profile_set.rs:313-359)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import BLOSUM_62, CudaProfiles, DNA_PROFILE_MAP, SeqSrc, WeightMatrix, synth
from zoe_b200.alignment import Status

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")


def osc(wm, go, ge):
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


def check_ranges(targets, seqs, wm, go=-10, ge=-1, profiled_is_query=False, lanes=(32, 16, 8), policy=None):
    targets = [bytes(t) for t in targets]
    seqs = [bytes(s) for s in seqs]
    prof = CudaProfiles(targets, wm, go, ge, lanes=lanes, profiled_is_query=profiled_is_query)
    if policy:
        prof.set_width_policy(*policy)
    src = SeqSrc.Reference(seqs) if profiled_is_query else SeqSrc.Query(seqs)
    got = prof.sw_score_ranges_batch(src)
    stats = prof.last_stats()
    sc = osc(wm, go, ge)
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            if policy and policy[0] == policy[1]:
                rc, score, rr, qr = O.striped_score_ranges(t, s, sc, policy[0], lanes[0], signed=not policy[2],
                                                           streamed_is_query=not profiled_is_query)
            else:
                rc, score, rr, qr, _ = O.sw_score_ranges_from(t, s, sc, lanes=lanes, first_bits=policy[0] if policy else 8,
                                                              streamed_is_query=not profiled_is_query)
            g = got[i][j]
            assert g.status.value == rc, (i, j, g, rc, score)
            if rc == O.SOME:
                a = g.unwrap()
                assert (a.score, a.ref_range, a.query_range) == (score, rr, qr), (i, j, a, score, rr, qr)
    prof.close()
    return stats


def test_doc_example_and_zoe_tests():
    # doc profile_set.rs:293-310 (score 26, query 0..15, ref 14..31); sw/test.rs:197-262
    prof = CudaProfiles.new_with_w256([b"CGTTCGCCATAAAGGGGG", b"GGGGGGGCCCCCAAAA", b"CCCCA"], W42, -3, -1, profiled_is_query=True)
    r = prof.sw_score_ranges_batch(SeqSrc.Reference([b"ATGCATCGATCGATCGATCGATCGATCGATGC"]))
    a = r[0][0].unwrap()
    assert (a.score, a.query_range, a.ref_range) == (26, (0, 15), (14, 31))
    prof.close()
    check_ranges([b"GGGGGGGCCCCCAAAA", b"CCCCA"], [b"TTTTTTCCTTTTTTTTCCCCCTTTTT", b"TAAAA"], W25, profiled_is_query=True,
                 lanes=(8, 8, 8))


def test_ranges_agree_with_alignment_ranges_on_config3_sample():
    """zoe's own test asserts ranges == the alignment's ranges for its examples; on the benchmark reads both GPU
    paths (ranges without a matrix, align with traceback) are additionally checked against the oracle."""
    targets, reads = synth.config3(ROOT, n_reads=300, seed=91)
    stats = check_ranges(targets, list(reads), W25)
    assert stats["tier16"] > 50 and stats["tier8"] > 50


def test_random_pairs_many_scorings_both_orientations():
    rng = np.random.default_rng(71)
    for (ma, mi, go, ge) in [(2, -5, -10, -1), (4, -2, -3, -1), (1, -1, -4, -2), (3, -1, -4, -1), (1, -1, -1, -1),
                             (5, -4, 0, 0)]:
        wm = WeightMatrix.new_dna_matrix(ma, mi, b"N")
        targets = [synth.random_dna(rng, int(L)) for L in (37, 300, 64, 513)]
        seqs = []
        for _ in range(61):
            L = int(rng.integers(1, 150))
            s = synth.random_dna(rng, L)
            if L > 24:
                t = targets[int(rng.integers(0, 4))]
                k = min(L - 4, len(t) - 2, 120)
                st = int(rng.integers(0, len(t) - k + 1))
                frag = synth._mutate(rng, t[st:st + k], 0.06, 0.05, 0.05, np.frombuffer(b"ACGT", dtype=np.uint8))
                k2 = min(len(frag), L - 2)
                s[2:2 + k2] = frag[:k2]
            seqs.append(s)
        seqs.append(np.zeros(0, dtype=np.uint8))
        check_ranges(targets, seqs, wm, go, ge)
        check_ranges(targets[:2], seqs[:30], wm, go, ge, profiled_is_query=True, lanes=(16, 8, 4))


def test_repeats_and_ties():
    # low-complexity sequences: many equal-score cells, the (min row, min column) rules decide both passes
    wm = WeightMatrix.new_dna_matrix(1, -1, b"N")
    targets = [b"ACACACACACACACACACAC", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"ACGTACGTACGTACGTACGTACGT"]
    seqs = [b"ACACAC", b"CACACACA", b"AAAAAAAAAA", b"A", b"ACGTACGT", b"GTACGTAC", b"TTTTTT", b"ACGTTTACGT", b"CACAAACACA"]
    check_ranges(targets, seqs, wm, -2, -1)
    check_ranges(targets, seqs, wm, -2, -1, profiled_is_query=True)


def test_protein_and_width_policies():
    targets, q = synth.config5(n_queries=40)
    check_ranges(targets, list(q), BLOSUM_62)
    rng = np.random.default_rng(5)
    t = [synth.random_dna(rng, 200)]
    seqs = [t[0][10:150].copy(), synth.random_dna(rng, 90), t[0][:64].copy()]
    for policy in [(16, 32, False), (8, 8, False), (8, 8, True), (16, 16, True)]:
        check_ranges(t, seqs, W25, policy=policy, profiled_is_query=True, lanes=(16, 16, 16) if policy[0] == policy[1] else (32, 16, 8))
    # scores beyond the packed 16-bit lanes: the 32-bit instantiation of both passes
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    stats = check_ranges([b"A" * 600], [b"A" * 600, b"A" * 300, b"A" * 100 + b"C" + b"A" * 100], w)
    assert stats["tier32"] == 1


def test_edge_cases():
    prof = CudaProfiles.new_with_w256([b"ACGTACGTAC", b"A"], W25, -10, -1)
    r = prof.sw_score_ranges_batch(SeqSrc.Query([b"", b"C", b"TTTT", b"ACGTACGTAC", b"A"]))
    assert r[0][0].status is Status.Unmapped and r[0][1].status is Status.Unmapped
    assert r[1][1].status is Status.Unmapped
    a = r[3][0].unwrap()
    assert (a.score, a.ref_range, a.query_range) == (20, (0, 10), (0, 10))
    a = r[4][1].unwrap()
    assert (a.score, a.ref_range, a.query_range) == (2, (0, 1), (0, 1))
    prof.close()
