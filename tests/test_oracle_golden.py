"""Pins the CPU oracle (oracle/zoe_sw_oracle.c) against every known answer zoe's own tests and
doc-tests hold for the striped SW path.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200.matrices import DNA_PROFILE_MAP, ByteIndexMap, WeightMatrix

GAP_OPEN, GAP_EXTEND = -10, -1  # src/alignment/sw/mod.rs:469-470


def sc_of(wm: WeightMatrix, go=GAP_OPEN, ge=GAP_EXTEND) -> O.Scoring:
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")


def test_validate_profile_args():
    # src/alignment/profile.rs:32-44
    assert O.validate_profile_args(0, -10, -1) == O.ERR_EMPTY_SEQUENCE
    assert O.validate_profile_args(5, -128, -1) == O.ERR_GAP_OPEN_RANGE
    assert O.validate_profile_args(5, 1, -1) == O.ERR_GAP_OPEN_RANGE
    assert O.validate_profile_args(5, -10, 1) == O.ERR_GAP_EXTEND_RANGE
    assert O.validate_profile_args(5, -1, -10) == O.ERR_BAD_GAP_WEIGHTS
    assert O.validate_profile_args(5, -10, -1) == 0
    assert O.validate_profile_args(5, 0, 0) == 0
    assert O.validate_profile_args(5, -127, -127) == 0


def test_sw_scalar(seqs):
    # sw/test.rs:71-84
    sc = sc_of(W25)
    rc, aln = O.scalar_align(seqs["H1_HA"], seqs["H5_HA"], sc)
    assert rc == O.SOME and (aln.ref_range[0], aln.score) == (336, 37)
    assert aln.cigar == "267S29M685S" and aln.ref_range == (336, 365) and aln.query_range == (267, 296)
    assert O.scalar_score(seqs["H1_HA"], seqs["H5_HA"], sc) == (O.SOME, 37)
    v = b"A" * 100
    assert O.scalar_score(v, v, sc) == (O.SOME, 200)


def test_sw_t_u_check():
    # sw/test.rs:87-100
    sc = sc_of(W25)
    assert O.scalar_score(b"ACGTUNacgtun", b"ACGTTNACGTTN", sc) == (O.SOME, 20)
    assert O.striped_score(b"ACGTUNacgtun", b"ACGTTNACGTTN", sc, 16, 16, signed=False) == (O.SOME, 20)
    assert O.striped_score(b"ACGTUNacgtun", b"ACGTTNACGTTN", sc, 16, 16, signed=True) == (O.SOME, 20)


def test_sw_simd(seqs):
    # sw/test.rs:102-114
    sc = sc_of(W25)
    assert O.striped_score(seqs["H5_HA"], seqs["H1_HA"], sc, 8, 16, signed=False) == (O.SOME, 37)
    assert O.striped_score(seqs["H5_HA"], seqs["H1_HA"], sc, 16, 16, signed=True) == (O.SOME, 37)


def _test_sw_simd_align(profile_seq, other_seq, bits, lanes):
    """The reference's macro test_sw_simd_align! (sw/test.rs:7-51)."""
    sc = sc_of(W25)
    rc, score = O.striped_score(profile_seq, other_seq, sc, bits, lanes, signed=True)
    rc_s, aln_scalar = O.scalar_align(profile_seq, other_seq, sc)
    rc_v, aln_simd = O.striped_align(profile_seq, other_seq, sc, bits, lanes, signed=True)
    assert rc == rc_s == rc_v == O.SOME
    assert score == aln_scalar.score
    assert aln_scalar == aln_simd
    rc_u, aln_u = O.striped_align(profile_seq, other_seq, sc, bits, lanes, signed=False)
    assert rc_u == O.SOME and aln_scalar == aln_u
    rc_e, s_e, ref_end, query_end = O.striped_score_ends(profile_seq, other_seq, sc, bits, lanes, signed=False)
    assert rc_e == O.SOME and s_e == aln_scalar.score
    assert aln_scalar.ref_range[1] == ref_end and aln_scalar.query_range[1] == query_end
    return aln_scalar


def test_sw_simd_aln(seqs):
    # sw/test.rs:116-119
    aln = _test_sw_simd_align(seqs["H5_HA"], seqs["H1_HA"], 16, 8)
    assert aln.cigar == "336S29M1395S"


# sw/test.rs:121-195 ; CIGARs from SURVEY.md section 8(c)
ALN_CASES = [
    (b"TTTAG", b"AAACTA", 2, "2S2M1S"),
    (b"AAAAAATAAA", b"AAAAAAAAAA", 4, "10M"),
    (b"CCCCA", b"TAAAA", 4, "4S1M"),
    (b"CCCCC", b"TCCCC", 4, "4M1S"),
    (b"CCCCT", b"GCTTTTC", 4, "3S2M"),
    (b"TTTTTGTTTTCTTTTTTGTTTA", b"TTGTTTTTTTTTGTT", 16, "3S7M2I8M2S"),
    (b"TTGTTTTGGGGAAAAA", b"TTTTTGTTTGGGAAAAATTCTT", 8, "6M2I8M"),
    (b"TTTTTTTCTTGTTTTTG", b"TTTTTGTTTTCTTGGT", 16, "3S8M6S"),
    (b"TTTTTTTTACTATTTTTAAATTTATGTTTTGTTA", b"TTTTTTTTTTTTAAAATTTGTAAACGTTTTGTTA", 8, "8M4I13M4D9M"),
    (b"TTTTTTTTTTTTTTTTTTTCCTTTTTTTTTTTTTTTTTTTTTTTTTTCCCCCCTTTA",
     b"TTTATTTTTTTTTTTTTTCCCCCCCTTTTTTTTTTTTTTTTTCCCCCCTTT", 8, "21S9M7D26M1S"),
    (b"TTTTTTTTTTTTTTTCCTTTTTTTTTTTTTTTTTTTCCCCCCCCCTA", b"TTTTTTTTTTTTTTTCCCCCTTTTTTTTTTCCCCCCCCCTT", 8,
     "9S6M2I9M5D20M1S"),
]


@pytest.mark.parametrize("query,reference,lanes,cigar", ALN_CASES)
def test_sw_simd_aln_small(query, reference, lanes, cigar):
    aln = _test_sw_simd_align(query, reference, 8, lanes)
    assert aln.cigar == cigar


def test_sw_simd_poly_a_single_profile_set(seqs):
    # sw/test.rs:264-280, 303-311
    sc = sc_of(W25)
    v = b"A" * 100
    assert O.striped_score(v, v, sc, 16, 16, signed=False) == (O.SOME, 200)
    cy = seqs["CY137594"]
    assert O.striped_score(cy, cy, sc, 16, 16, signed=False) == (O.SOME, 3372)
    rc, score, tier = O.sw_score_from(cy, cy, sc, lanes=(16, 8, 4))  # new_with_w128
    assert (rc, score, tier) == (O.SOME, 3372, 16)


def test_sw_simd_regression():
    # sw/test.rs:282-290 : lazy-F with gap_open == gap_extend
    wm = WeightMatrix.new(DNA_PROFILE_MAP, 10, -10, b"N")
    assert O.striped_score(b"AGA", b"AA", sc_of(wm, -5, -5), 16, 4, signed=False) == (O.SOME, 15)


def test_sw_simd_overflow_check():
    # sw/test.rs:292-301
    wm = WeightMatrix.new(DNA_PROFILE_MAP, 127, 0, b"N")
    rc, _ = O.striped_score(b"AAAA", b"AAAA", sc_of(wm), 8, 8, signed=False)
    assert rc == O.OVERFLOWED


def test_profile_set_w256_profile_equality(seqs):
    # profile_set.rs:704-712 : LocalProfiles::new_with_w256(..).get_i8() == StripedProfile::<i8,32,5>::new
    sc = sc_of(W25)
    cy = seqs["CY137594"]
    prof = O.striped_profile(cy, sc, 8, 32)
    S, nv, N = prof.shape
    assert (S, nv, N) == (5, (len(cy) + 31) // 32, 32)
    idx = DNA_PROFILE_MAP.index_map
    for v in (0, 1, nv - 1):
        for lane in (0, 5, 31):
            q = v + lane * nv
            for a in range(5):
                want = int(W25.weights[a, idx[cy[q]]]) if q < len(cy) else 0
                assert prof[a, v, lane] == want


def test_doc_striped():
    # doc striped.rs:44-56 and :418-441
    sc = sc_of(W42, -3, -1)
    reference, query = b"ATGCATCGATCGATCGATCGATCGATCGATGC", b"CGTTCGCCATAAAGGGGG"
    assert O.striped_score(query, reference, sc, 8, 32, signed=False) == (O.SOME, 26)
    rc, aln = O.striped_align(query, reference, sc, 8, 8, signed=False)
    assert rc == O.SOME and aln.cigar == "6M2D9M3S" and aln.score == 26


def test_doc_profile_set():
    # doc profile_set.rs:293-310 (w256)
    sc = sc_of(W42, -3, -1)
    reference, query = b"ATGCATCGATCGATCGATCGATCGATCGATGC", b"CGTTCGCCATAAAGGGGG"
    rc, aln, tier = O.sw_align_from(query, reference, sc, lanes=(32, 16, 8))
    assert rc == O.SOME and tier == 8
    assert aln.score == 26 and aln.query_range == (0, 15) and aln.ref_range == (14, 31)


def test_doc_sw_mod():
    # doc sw/mod.rs:164-188, 220-247
    sc = sc_of(W42, -3, -1)
    reference, query = b"GGCCACAGGATTGAG", b"CTCAGATTG"
    rc, aln = O.striped_align(query, reference, sc, 8, 32, signed=True)
    assert rc == O.SOME and aln.score == 27 and aln.cigar == "5M1D4M"
    rc, aln, _ = O.sw_align_from(query, reference, sc, lanes=(32, 16, 8))
    assert rc == O.SOME and aln.score == 27 and aln.cigar == "5M1D4M" and aln.ref_range[0] == 3
    rc, aln = O.scalar_align(query, reference, sc)
    assert aln.ref_range[0] == 3 and aln.cigar == "5M1D4M" and aln.score == 27
    # doc sw/mod.rs:190-218 : custom alphabet ABCD
    mapping = ByteIndexMap(b"ABCD", b"A", ignore_case=False)
    wm = WeightMatrix.new(mapping, 1, -1, None)
    sc2 = O.Scoring(wm.weights, mapping.index_map, -4, -2)
    rc, aln = O.striped_align(b"AABDDAB", b"BDAACAABDDDB", sc2, 8, 32, signed=True)
    assert rc == O.SOME and aln.score == 5 and aln.cigar == "5M2S"


def test_types_invert():
    # types/test.rs:14-38
    sc = sc_of(W42, -3, -1)
    reference, query = b"GGCCACAGGATTGAGC", b"TCTCAGATTGCAGTTT"
    rc, aln = O.scalar_align(query, reference, sc)
    assert rc == O.SOME
    assert aln.ref_range == (3, 15) and aln.query_range == (1, 13) and aln.cigar == "1S5M1D4M1I2M3S"
    rc, inv = O.scalar_align(query, reference, sc, streamed_is_query=True)
    assert inv.cigar == "3S5M1I4M1D2M1S"
    assert inv.ref_range == (1, 13) and inv.query_range == (3, 15)
    assert (inv.ref_len, inv.query_len) == (len(query), len(reference))


def test_tier_boundaries():
    # striped.rs:608-633: i8 valid 1..=254, i16 valid 1..=65534
    wm = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    sc = sc_of(wm)
    assert O.striped_score(b"AA", b"AA", sc, 8, 32) == (O.SOME, 254)
    wm1 = WeightMatrix.new(DNA_PROFILE_MAP, 85, -5, b"N")
    assert O.striped_score(b"AAA", b"AAA", sc_of(wm1), 8, 32)[0] == O.OVERFLOWED  # 255
    assert O.sw_score_from(b"AAA", b"AAA", sc_of(wm1)) == (O.SOME, 255, 16)
    assert O.striped_score(b"CCCC", b"AAAA", sc, 8, 32)[0] == O.UNMAPPED
    rc, aln = O.striped_align(b"ACGT", b"", sc, 8, 32)
    assert rc == O.UNMAPPED
    # i16 -> i32: 600 x match 127 = 76200 > 65534
    big = b"A" * 600
    assert O.sw_score_from(big, big, sc) == (O.SOME, 76200, 32)


def test_striped_equals_scalar_random():
    rng = np.random.default_rng(7)
    sc = sc_of(W25)
    for _ in range(150):
        m, n = rng.integers(1, 70, 2)
        p = bytes(rng.choice(list(b"ACGT"), m).tolist())
        r = bytearray(rng.choice(list(b"ACGT"), n).tolist())
        if m > 12 and n > 12:  # plant a noisy copy so alignments are non-trivial
            k = int(min(m, n) * 0.8)
            r[:k] = p[:k]
            for _ in range(3):
                r[rng.integers(0, k)] = b"ACGT"[rng.integers(0, 4)]
        r = bytes(r)
        rc_s, a_s = O.scalar_align(p, r, sc)
        for bits, lanes in ((8, 32), (16, 16), (32, 8), (8, 4)):
            rc_v, a_v = O.striped_align(p, r, sc, bits, lanes)
            assert rc_v == rc_s
            if rc_s == O.SOME:
                assert a_v.score == a_s.score and a_v.ref_range == a_s.ref_range and a_v.query_range == a_s.query_range
                assert O.striped_score(p, r, sc, bits, lanes) == (O.SOME, a_s.score)


def test_sw_simd_locations_and_ranges():
    """sw/test.rs:197-262 (sw_simd_locations, sw_simd_ranges) and the doc example profile_set.rs:293-310."""
    sc = sc_of(W25)
    for query, reference in [(b"GGGGGGGCCCCCAAAA", b"TTTTTTCCTTTTTTTTCCCCCTTTTT"), (b"CCCCA", b"TAAAA")]:
        rc_s, aln_scalar = O.scalar_align(query, reference, sc)
        rc, score, ref_range, query_range = O.striped_score_ranges(query, reference, sc, 8, 8, signed=False)
        assert rc == rc_s == O.SOME and score == aln_scalar.score
        assert ref_range == tuple(aln_scalar.ref_range) and query_range == tuple(aln_scalar.query_range)
        rc_e, s_e, ref_end, query_end = O.striped_score_ends(query, reference, sc, 8, 8, signed=False)
        assert (rc_e, s_e, ref_end, query_end) == (O.SOME, score, ref_range[1], query_range[1])
    sc42 = O.Scoring(W42.weights, W42.mapping.index_map, -3, -1)
    rc, score, ref_range, query_range, tier = O.sw_score_ranges_from(
        b"CGTTCGCCATAAAGGGGG", b"ATGCATCGATCGATCGATCGATCGATCGATGC", sc42, lanes=(32, 16, 8))
    assert (rc, score, ref_range, query_range, tier) == (O.SOME, 26, (14, 31), (0, 15), 8)
    # SeqSrc::Query(streamed): the ranges swap sides (alignment/mod.rs:176-190)
    rc, score, ref_range, query_range, _ = O.sw_score_ranges_from(
        b"CGTTCGCCATAAAGGGGG", b"ATGCATCGATCGATCGATCGATCGATCGATGC", sc42, streamed_is_query=True)
    assert (rc, score, ref_range, query_range) == (O.SOME, 26, (0, 15), (14, 31))
