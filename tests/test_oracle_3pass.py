"""Pins the oracle's restatement of zoe's banded alignment and 3-pass algorithm (src/alignment/sw/banded.rs,
three_pass.rs, profile_set.rs:185-290) on the known answers zoe's own tests / doc-tests hold, and on the properties
those tests state.  CPU only."""
import re

import numpy as np

from oracle import oracle as O
from zoe_b200.matrices import WeightMatrix

W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")
W21 = WeightMatrix.new_dna_matrix(2, -1, b"N")


def sc_of(wm, go, ge):
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


def rescore(aln: O.Aln, ref: bytes, query: bytes, wm: WeightMatrix, go: int, ge: int) -> int:
    """Score implied by a CIGAR over its ranges (M = substitution weight, gap of length L = go + (L-1)*ge)."""
    r, q = aln.ref_range[0], None
    total = 0
    ops = [(int(n), op) for n, op in re.findall(r"(\d+)([MIDS])", aln.cigar)]
    q = 0
    first = True
    for n, op in ops:
        if op == "S":
            if first:
                q += n
        elif op == "M":
            for _ in range(n):
                total += wm.get_weight(ref[r], query[q])
                r += 1
                q += 1
        elif op == "D":
            total += go + (n - 1) * ge
            r += n
        elif op == "I":
            total += go + (n - 1) * ge
            q += n
        first = False
    assert r == aln.ref_range[1]
    return total


def test_banded_known_answers():
    # sw/test.rs:353-369 test_banded_sw_align_simple
    rc, aln = O.banded_align(b"AACCGG", b"AAACCCGGG", sc_of(W21, -2, -1), 3)
    assert rc == O.SOME and aln.score == 10
    assert aln.ref_range[1] > aln.ref_range[0] and aln.query_range[1] > aln.query_range[0]
    # sw/test.rs:371-380 test_banded_empty_sequences
    assert O.banded_align(b"", b"ACGT", sc_of(W21, -2, -1), 3)[0] == O.ERR_EMPTY_SEQUENCE
    assert O.banded_align(b"ACGT", b"", sc_of(W21, -2, -1), 3)[0] == O.UNMAPPED
    # banded.rs:22-38 doc example and sw/test.rs:314-348 sw_banded_comparison: Some, positive score
    sc = sc_of(W42, -3, -1)
    rc, aln = O.banded_align(b"CTCAGATTG", b"GGCCACAGGATTGAG", sc, 5)
    assert rc == O.SOME and aln.score > 0
    rc, wide = O.banded_align(b"CTCAGATTG", b"GGCCACAGGATTGAG", sc, 15)
    assert rc == O.SOME and wide.score == 27  # the scalar score of this pair (sw/mod.rs:164-188)
    assert O.banded_align(b"CTCAGATTG", b"GGCCACAGGATTGAG", sc, 3)[0] in (O.SOME, O.UNMAPPED)


def test_banded_full_band_equals_scalar():
    # with a band that covers the whole table the banded recurrence is sw_scalar_align's (sw/test.rs:36-49 runs this
    # configuration next to the scalar alignment)
    rng = np.random.default_rng(11)
    sc = sc_of(W25, -10, -1)
    for _ in range(300):
        m, n = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        p = bytes(rng.choice(list(b"ACGT"), m).astype(np.uint8))
        s = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
        a = O.banded_align(p, s, sc, m + n)
        b = O.scalar_align(p, s, sc)
        assert a == b, (p, s)


def test_3pass_doc_example():
    # profile_set.rs:192-208: LocalProfiles::new_with_w256(query,..).sw_align_from_i8_3pass(SeqSrc::Reference(reference))
    sc = sc_of(W42, -3, -1)
    ref, query = b"ATGCATCGATCGATCGATCGATCGATCGATGC", b"CGTTCGCCATAAAGGGGG"
    rc, aln, tier, path = O.sw_align_3pass_from(query, ref, sc, lanes=(32, 16, 8))
    assert rc == O.SOME and aln.score == 26 and tier == 8
    # same pair through sw_align_from_i8 (doc profile_set.rs:293-310): same score and ranges
    rc2, full, _ = O.sw_align_from(query, ref, sc, lanes=(32, 16, 8))
    assert (aln.ref_range, aln.query_range) == (full.ref_range, full.query_range) == ((14, 31), (0, 15))
    assert rescore(aln, ref, query, W42, -3, -1) == 26


def _mutate(rng, seq, sub, indel):
    out = []
    for b in seq:
        u = rng.random()
        if u < indel:
            continue
        if u < 2 * indel:
            out.append(int(rng.choice(list(b"ACGT"))))
        out.append(int(rng.choice(list(b"ACGT"))) if rng.random() < sub else b)
    return bytes(out)


def test_3pass_properties():
    # three_pass.rs: the score and ranges are those of sw_simd_score_ranges (== the scalar alignment's,
    # sw/test.rs:197-262); the states must be a valid path of that score over those ranges.
    rng = np.random.default_rng(5)
    paths = {0: 0, 1: 0, 2: 0}
    for trial in range(400):
        wm, go, ge = [(W25, -10, -1), (W42, -3, -1), (W21, -2, -1), (W42, -4, -4)][trial % 4]
        sc = sc_of(wm, go, ge)
        m = int(rng.integers(20, 120))
        target = bytes(rng.choice(list(b"ACGT"), m).astype(np.uint8))
        a, b = sorted(int(x) for x in rng.integers(0, m + 1, 2))
        read = _mutate(rng, target[a:b], 0.03, 0.03) or b"A"
        if trial % 5 == 0:
            read = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 60))).astype(np.uint8))
        for inv in (False, True):
            rc, aln, tier, path = O.sw_align_3pass_from(target, read, sc, lanes=(32, 16, 8), streamed_is_query=inv)
            rs = O.sw_score_ranges_from(target, read, sc, lanes=(32, 16, 8), streamed_is_query=inv)
            assert rc == rs[0], (target, read)
            if rc != O.SOME:
                continue
            assert (aln.score, aln.ref_range, aln.query_range, tier) == (rs[1], rs[2], rs[3], rs[4]), (target, read, aln)
            ref, query = (target, read) if inv else (read, target)
            assert (aln.ref_len, aln.query_len) == (len(ref), len(query))
            assert rescore(aln, ref, query, wm, go, ge) == aln.score, (target, read, aln, path)
            ops = [(int(n), op) for n, op in re.findall(r"(\d+)([MIDS])", aln.cigar)]
            assert sum(n for n, op in ops if op in "MI") == aln.query_range[1] - aln.query_range[0]
            assert sum(n for n, op in ops if op in "MIS") == len(query)
            paths[path & 0xff] += 1
    assert paths[0] > 0 and paths[1] > 0 and paths[2] > 0, paths
