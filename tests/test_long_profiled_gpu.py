"""Profiled sequences longer than 131 070 residues: pass A of the windowed pipeline keeps step-pair indices in 16-bit
halves, so such sequences must take another route (or be refused) -- never a silently wrong answer."""
import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import CudaProfiles, SeqSrc, WeightMatrix, synth

pytestmark = pytest.mark.gpu
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")


@pytest.mark.parametrize("length", [131_000, 140_000, 270_000])
def test_reads_against_a_very_long_profiled_sequence(length):
    rng = np.random.default_rng(length)
    target = synth.random_dna(rng, length)
    reads = []
    for st in (0, 77, 65_000, 131_050, length - 150, length - 151, length // 2):
        st = min(st, length - 150)
        r = target[st:st + 150].copy()
        r[10] = ord("A") if r[10] != ord("A") else ord("C")
        reads.append(r)
    reads.append(np.concatenate([target[length - 100:length - 50], target[length - 45:length]]))  # a deletion near the end
    reads.append(synth.random_dna(rng, 150))
    tb, sb = bytes(target), [bytes(r) for r in reads]
    sc = O.Scoring(W25.weights, W25.mapping.index_map, -10, -1)
    prof = CudaProfiles.new_with_w256([tb], W25, -10, -1)
    got_s = prof.sw_score_batch(sb)
    got_a = prof.sw_align_batch(SeqSrc.Query(sb))
    got_r = prof.sw_score_ranges_batch(SeqSrc.Query(sb))
    got_3 = prof.sw_align_3pass_batch(SeqSrc.Query(sb))
    prof.close()
    for i, s in enumerate(sb):
        rc, score, _ = O.sw_score_from(tb, s, sc)
        assert got_s[i][0].status.value == rc and (rc != 0 or got_s[i][0].unwrap() == score), ("score", i)
        rc, aln, _ = O.sw_align_from(tb, s, sc, streamed_is_query=True)
        assert got_a[i][0].status.value == rc, ("align", i)
        if rc == 0:
            a = got_a[i][0].unwrap()
            assert (a.score, a.ref_range, a.query_range, a.states) == (aln.score, aln.ref_range, aln.query_range, aln.cigar), ("align", i, a, aln)
        rc, score, rr, qr, _ = O.sw_score_ranges_from(tb, s, sc, streamed_is_query=True)
        assert got_r[i][0].status.value == rc, ("ranges", i)
        if rc == 0:
            a = got_r[i][0].unwrap()
            assert (a.score, a.ref_range, a.query_range) == (score, rr, qr), ("ranges", i, a, score, rr, qr)
        rc, aln, _, _ = O.sw_align_3pass_from(tb, s, sc, streamed_is_query=True)
        assert got_3[i][0].status.value == rc, ("3pass", i)
        if rc == 0:
            a = got_3[i][0].unwrap()
            assert (a.score, a.ref_range, a.query_range, a.states) == (aln.score, aln.ref_range, aln.query_range, aln.cigar), ("3pass", i, a, aln)
