"""GPU parity for the integer-width policies: ProfileSets::sw_*_from_i16 / _from_i32 (profile_set.rs:71-179) and
standalone StripedProfile<T,N,S> for signed and unsigned T (profile.rs:440-446, 515-519; unsigned = biased matrix,
matrices/mod.rs:471-491; status rule striped.rs:608-633)."""
import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import CudaProfiles, SeqSrc, WeightMatrix, synth

pytestmark = pytest.mark.gpu
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")
W42 = WeightMatrix.new_dna_matrix(4, -2, b"N")


def osc(wm, go, ge):
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


def _inputs(seed, n=50):
    rng = np.random.default_rng(seed)
    targets = [synth.random_dna(rng, int(L)) for L in (180, 75)]
    seqs = []
    for k in range(n):
        L = int(rng.integers(8, 170))
        s = synth.random_dna(rng, L)
        t = targets[k % 2]
        w = min(L, len(t)) - 4
        if w > 10 and k % 5:
            st = int(rng.integers(0, len(t) - w + 1))
            keep = rng.random(w) > (0.02 if k % 3 else 0.2)
            s[2:2 + w] = np.where(keep, t[st:st + w], s[2:2 + w])
        seqs.append(s)
    return [bytes(t) for t in targets], [bytes(s) for s in seqs]


@pytest.mark.parametrize("first_bits", [16, 32])
def test_from_i16_and_from_i32_chains(first_bits):
    targets, seqs = _inputs(first_bits)
    sc = osc(W25, -10, -1)
    prof = CudaProfiles.new_with_w256(targets, W25, -10, -1, profiled_is_query=True)
    prof.set_width_policy(first_bits, 32, False)
    got = prof.sw_score_batch(seqs)
    buf, offs = synth.pack([np.frombuffer(s, dtype=np.uint8) for s in seqs])
    _, _, tier = prof.sw_score_arrays(buf, offs)
    aln = prof.sw_align_batch(SeqSrc.Reference(seqs))
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            rc, score, want_tier = O.sw_score_from(t, s, sc, first_bits=first_bits)
            assert got[i][j].status.value == rc
            if rc == O.SOME:
                assert got[i][j].unwrap() == score and int(tier[i, j]) == want_tier
            rc, want, _ = O.sw_align_from(t, s, sc, first_bits=first_bits, streamed_is_query=False)
            assert aln[i][j].status.value == rc
            if rc == O.SOME:
                a = aln[i][j].unwrap()
                assert (a.score, a.ref_range, a.query_range, a.states) == (want.score, want.ref_range, want.query_range, want.cigar)
    prof.close()


@pytest.mark.parametrize("bits,signed,lanes", [(8, True, 32), (16, True, 16), (8, False, 32), (16, False, 16), (32, False, 8),
                                               (8, False, 16), (32, True, 8)])
@pytest.mark.parametrize("wm,go,ge", [(W25, -10, -1), (W42, -3, -1)])
def test_standalone_striped_profile_types(bits, signed, lanes, wm, go, ge):
    targets, seqs = _inputs(1000 + bits + lanes + (0 if signed else 7))
    sc = osc(wm, go, ge)
    prof = CudaProfiles(targets, wm, go, ge, lanes=(lanes, lanes, lanes), profiled_is_query=True)
    prof.set_width_policy(bits, bits, not signed)
    got = prof.sw_score_batch(seqs)
    aln = prof.sw_align_batch(SeqSrc.Reference(seqs))
    n_over = 0
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            rc, score = O.striped_score(t, s, sc, bits, lanes, signed=signed)
            assert got[i][j].status.value == rc, (i, j, got[i][j], rc, score)
            n_over += rc == O.OVERFLOWED
            if rc == O.SOME:
                assert got[i][j].unwrap() == score
            rc, want = O.striped_align(t, s, sc, bits, lanes, signed=signed, streamed_is_query=False)
            assert aln[i][j].status.value == rc, (i, j, aln[i][j], rc)
            if rc == O.SOME:
                a = aln[i][j].unwrap()
                assert (a.score, a.ref_range, a.query_range, a.states) == (want.score, want.ref_range, want.query_range, want.cigar)
    stats = prof.last_stats()
    assert stats["overflowed"] == n_over
    if bits == 8:
        assert n_over > 0  # near-identical reads score beyond the 8-bit types
    prof.close()


@pytest.mark.parametrize("policy", [(8, 32), (16, 32), (32, 32)])
def test_tiers_agree_across_entry_points_including_unmapped(policy):
    # `tier` is the width of the type that produced the result -- for Unmapped too (decided in the FIRST tier of
    # the chain), and the same from score, align, ranges and 3-pass
    targets, seqs = _inputs(7, n=30)
    seqs += [b"NNNNNNNNNNNN", b"N"]  # score 0 against anything: Unmapped
    import numpy as _np
    buf = _np.frombuffer(b"".join(seqs), dtype=_np.uint8)
    offs = _np.zeros(len(seqs) + 1, dtype=_np.uint64)
    offs[1:] = _np.cumsum([len(s) for s in seqs])
    prof = CudaProfiles.new_with_w256(targets, W25, -10, -1)
    prof.set_width_policy(policy[0], policy[1])
    score, status, tier = prof.sw_score_arrays(buf, offs)
    al = prof.align_arrays(buf, offs)
    rg = prof.ranges_arrays(buf, offs)
    tp = prof.align_arrays(buf, offs, three_pass=True)
    prof.close()
    n_pairs = len(seqs) * len(targets)
    sc = osc(W25, -10, -1)
    for name, got in (("align", al), ("ranges", rg), ("3pass", tp)):
        assert _np.array_equal(got["status"][:n_pairs], status.ravel()), name
        assert _np.array_equal(got["tier"][:n_pairs], tier.ravel()), name
    assert (status.ravel() == 2).any()
    assert (tier.ravel()[status.ravel() == 2] == policy[0]).all()
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            rc, _, want_tier = O.sw_score_from(t, s, sc, first_bits=policy[0])
            assert int(status[i, j]) == rc
            if rc != 1:
                assert int(tier[i, j]) == want_tier, (i, j, rc, tier[i, j], want_tier)
