"""GPU parity: zoe_cuda_sw_score_batch (through the C ABI) vs the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from zoe_b200 import BLOSUM_62, CudaProfiles, DNA_PROFILE_MAP, WeightMatrix, synth
from zoe_b200.alignment import Status

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")


def osc(wm, go=-10, ge=-1):
    return O.Scoring(wm.weights, wm.mapping.index_map, go, ge)


def oracle_scores(targets, seqs, wm, go, ge, lanes=(32, 16, 8)):
    sc = osc(wm, go, ge)
    out = []
    for s in seqs:
        row = []
        for t in targets:
            row.append(O.sw_score_from(bytes(t), bytes(s), sc, lanes=lanes))
        out.append(row)
    return out


def check(targets, seqs, wm, go=-10, ge=-1):
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, go, ge)
    buf, offs = synth.pack([np.frombuffer(bytes(s), dtype=np.uint8) for s in seqs])
    score, status, tier = prof.sw_score_arrays(buf, offs)
    want = oracle_scores(targets, seqs, wm, go, ge)
    for i in range(len(seqs)):
        for j in range(len(targets)):
            rc, sc_, t_ = want[i][j]
            assert int(status[i, j]) == rc, (i, j, status[i, j], want[i][j])
            if rc == O.SOME:
                assert int(score[i, j]) == sc_, (i, j, score[i, j], want[i][j])
                assert int(tier[i, j]) == t_, (i, j, tier[i, j], want[i][j])
    prof.close()
    return score, status, tier


def test_golden_known_answers(seqs):
    # sw/test.rs known answers, through the CUDA path
    prof = CudaProfiles.new_with_w256([seqs["H5_HA"]], W25, -10, -1)
    r = prof.sw_score_batch([seqs["H1_HA"], b"A" * 100])
    assert r[0][0].unwrap() == 37
    prof.close()
    prof = CudaProfiles.new_with_w256([b"A" * 100, b"ACGTUNacgtun"], W25, -10, -1)
    r = prof.sw_score_batch([b"A" * 100, b"ACGTTNACGTTN", b"CCCC"])
    assert r[0][0].unwrap() == 200 and r[1][1].unwrap() == 20
    assert r[2][0].status is Status.Unmapped
    prof.close()
    cy = seqs["CY137594"]
    prof = CudaProfiles.new_with_w128([cy], W25, -10, -1)
    # 1686 rows > single-pass row capacity is allowed to be unsupported for now; 3372 via profile side
    prof.close()
    w42 = WeightMatrix.new_dna_matrix(4, -2, b"N")
    prof = CudaProfiles.new_with_w256([b"CGTTCGCCATAAAGGGGG", b"CTCAGATTG"], w42, -3, -1)
    r = prof.sw_score_batch([b"ATGCATCGATCGATCGATCGATCGATCGATGC", b"GGCCACAGGATTGAG"])
    assert r[0][0].unwrap() == 26 and r[1][1].unwrap() == 27
    prof.close()
    w = WeightMatrix.new(DNA_PROFILE_MAP, 10, -10, b"N")
    prof = CudaProfiles.new_with_w256([b"AGA"], w, -5, -5)
    assert prof.sw_score_batch([b"AA"])[0][0].unwrap() == 15
    prof.close()


def test_config1_sample_vs_oracle():
    targets, reads = synth.config1(ROOT, n_reads=600)
    score, status, tier = check(targets, list(reads), W25)
    assert (tier == 16).sum() > 200  # forward-strand true-origin reads escalate i8 -> i16 (half the reads are reverse strand)
    assert (tier == 8).sum() > 10


def test_config2_sample_vs_oracle():
    targets, reads = synth.config2(n_reads=160)
    check(targets, list(reads), W25)


def test_ragged_and_edge_lengths():
    rng = np.random.default_rng(11)
    target = synth.random_dna(rng, 300)
    seqs = []
    for L in [1, 2, 3, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 151, 152, 153, 200, 255, 256, 257, 300, 500, 777, 1024]:
        s = synth.random_dna(rng, L)
        k = min(L, 300)
        s[:k] = target[:k]  # related prefix so scores are non-trivial
        if L > 20:
            s[L // 2] = ord("N")
        seqs.append(s)
    seqs.append(np.zeros(0, dtype=np.uint8))  # empty streamed sequence -> Unmapped
    score, status, tier = check([target, target[:5], target[:64]], seqs, W25)
    assert int(status[-1, 0]) == O.UNMAPPED


def test_protein_blosum62_vs_oracle():
    targets, q = synth.config5(n_queries=120)
    score, status, tier = check(targets, list(q), BLOSUM_62)
    assert (tier == 16).any() and (tier == 8).any()


def test_protein_ragged_lengths_shared_rows_kernel():
    """The transposed (profiled-sequence-in-registers) kernel: ragged pairs, odd counts, pad columns, several
    targets of different lengths, unusual residues (B, Z, X, *), and a scoring whose packed lanes overflow."""
    rng = np.random.default_rng(41)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYBZX*", dtype=np.uint8)
    targets = [aa[rng.integers(0, 20, L)] for L in (566, 97, 1024, 33)]
    seqs = []
    for k in range(75):  # odd number of sequences: the last task / trip is partial
        L = int(rng.integers(64, 420)) if k else 420
        q = aa[rng.integers(0, 24, L)]
        t = targets[k % 4]
        w = min(L, len(t)) - 6
        st = int(rng.integers(0, len(t) - w + 1))
        keep = rng.random(w) > 0.25
        q[3:3 + w] = np.where(keep, t[st:st + w], q[3:3 + w])
        seqs.append(q)
    score, status, tier = check(targets, seqs, BLOSUM_62)
    assert (tier == 16).any() and (tier == 8).any()
    # packed overflow in the transposed kernel -> 32-bit re-run (i32 tier for the 600-mer)
    wide = WeightMatrix(BLOSUM_62.mapping, np.where(np.eye(25, dtype=bool), 120, -4).astype(np.int8))
    t600 = aa[rng.integers(0, 20, 600)]
    qs = [t600.copy(), t600[:300].copy(), aa[rng.integers(0, 20, 200)], t600[100:400].copy()]
    score, status, tier = check([t600], qs, wide)
    assert int(tier[0, 0]) == 32 and int(score[0, 0]) == 600 * 120


def test_i32_escalation():
    # match = 127: 600-mers score 76200 > 65534 -> i32 tier; 300-mers land in the i16 tier above 32767
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    a = b"A" * 600
    rng = np.random.default_rng(5)
    mixed = bytes(synth.random_dna(rng, 400))
    score, status, tier = check([a, mixed], [a, b"A" * 300, b"A" * 258, mixed, mixed[:100], b"C" * 50], w)
    assert int(score[0, 0]) == 76200 and int(tier[0, 0]) == 32
    assert int(score[1, 0]) == 38100 and int(tier[1, 0]) == 16


def test_various_scorings_random():
    rng = np.random.default_rng(3)
    for (ma, mi, go, ge) in [(4, -2, -3, -1), (1, -1, -4, -2), (3, -1, -4, -1), (2, -5, -10, -10), (5, -4, 0, 0)]:
        wm = WeightMatrix.new_dna_matrix(ma, mi, b"N")
        target = synth.random_dna(rng, 120)
        seqs = []
        for _ in range(40):
            L = int(rng.integers(5, 160))
            s = synth.random_dna(rng, L)
            if L > 40:
                st = int(rng.integers(0, 120 - 30))
                s[5:35] = target[st:st + 30]
            seqs.append(s)
        check([target], seqs, wm, go, ge)


def test_long_rows_golden_self_score(seqs):
    # sw/test.rs:272-280, 303-311: CY137594 (1686 nt) vs itself = 3372, i8 -> i16 escalation; 1686 rows take the
    # chunked long-row kernel
    cy = seqs["CY137594"]
    prof = CudaProfiles.new_with_w128([cy, seqs["H5_HA"]], W25, -10, -1)
    buf, offs = synth.pack([np.frombuffer(cy, dtype=np.uint8), np.frombuffer(seqs["H1_HA"], dtype=np.uint8)])
    score, status, tier = prof.sw_score_arrays(buf, offs)
    assert int(score[0, 0]) == 3372 and int(tier[0, 0]) == 16
    assert int(score[1, 1]) == 37 and int(tier[1, 1]) == 8
    prof.close()


def test_long_rows_vs_cpu_port_and_oracle():
    from oracle import cpu_baseline as CB
    targets, reads = synth.config4(n_reads=24, min_len=900, max_len=4200)
    genome = targets[0][:9000]
    reads = reads + [reads[0][:1024], reads[1][:1025], reads[2][:767], reads[3][:769], reads[4][:1]]
    prof = CudaProfiles.new_with_w256([bytes(genome), bytes(genome[:700])], W25, -10, -1)
    buf, offs = synth.pack(reads)
    score, status, tier = prof.sw_score_arrays(buf, offs)
    pbuf, poff = synth.pack([genome, genome[:700]])
    c_score, c_status, c_tier = CB.score_batch(pbuf, poff, buf, offs, W25.weights, W25.mapping.index_map, -10, -1)
    assert np.array_equal(status, c_status)
    assert np.array_equal(score[c_status == 0], c_score[c_status == 0])
    assert np.array_equal(tier[c_status == 0], c_tier[c_status == 0])
    assert (tier == 16).any()
    sc = osc(W25)
    for i in (0, 5, len(reads) - 2):  # spot-check the plain-C oracle as well
        rc, s_, t_ = O.sw_score_from(bytes(genome), bytes(reads[i]), sc)
        assert (int(status[i, 0]), int(score[i, 0]), int(tier[i, 0])) == (rc, s_, t_)
    prof.close()


def test_panel_larger_than_the_staging_area():
    """A profiled set beyond 96 KB of symbol codes is swept group by group (launch_score): every pair must still land
    in its own output slot with the score of the one-launch path.  Includes one sequence that alone exceeds the limit."""
    rng = np.random.default_rng(31)
    lens = [int(x) for x in rng.integers(900, 2400, 70)] + [100_000, 37, 1500]
    targets = [synth.random_dna(rng, L) for L in lens]
    assert sum(lens) > 2 * 96 * 1024
    reads = []
    for i in range(96):
        t = targets[(i * 7) % len(targets)]
        st = int(rng.integers(0, max(1, len(t) - 150)))
        r = t[st:st + 150].copy()
        flip = rng.random(len(r)) < 0.03
        r[flip] = synth.random_dna(rng, int(flip.sum()))
        reads.append(r)
    reads.append(synth.random_dna(rng, 150))
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
    buf, offs = synth.pack(reads)
    score, status, tier = prof.sw_score_arrays(buf, offs)
    assert prof.last_timing()["kernel_launches"] >= 3  # several group launches
    prof.close()
    sc = osc(W25)
    rs = np.random.default_rng(2)
    pairs = [(i, (i * 7) % len(targets)) for i in range(96)]  # the true-origin pair of every read
    pairs += [(int(rs.integers(0, len(reads))), int(rs.integers(0, len(targets)))) for _ in range(150)]
    pairs += [(i, 70) for i in range(0, 97, 16)]               # the 100-kb sequence (read from global memory)
    for i, j in pairs:
        rc, s_, t_ = O.sw_score_from(bytes(targets[j]), bytes(reads[i]), sc)
        assert (int(status[i, j]), int(score[i, j]) if rc == 0 else 0, int(tier[i, j]) if rc == 0 else 0) == \
               (rc, s_ if rc == 0 else 0, t_ if rc == 0 else 0), (i, j)


def _protein_batch(rng, target, n, min_len, max_len):
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    seqs = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        if i % 3 == 0:
            s = rng.choice(aa, L)
        else:
            st = int(rng.integers(0, max(1, len(target) - L)))
            s = target[st:st + L].copy()
            if len(s) < L:
                s = np.concatenate([s, rng.choice(aa, L - len(s))])
            flip = rng.random(L) < 0.2
            s[flip] = rng.choice(aa, int(flip.sum()))
        seqs.append(s.astype(np.uint8))
    return seqs


@pytest.mark.parametrize("min_len,max_len", [(300, 300), (64, 420), (64, 65)])
def test_streamed_trips_large_alphabet(min_len, max_len):
    """The shared-rows kernel sweeps a warp's trips back to back (sw_score_rows_stream_kernel): enough queries that every
    warp streams several trips, equal and ragged lengths, checked element-wise against the vectorised CPU port (which
    tests/test_cpu_baseline.py ties to the plain-C oracle)."""
    from oracle import cpu_baseline as CB
    rng = np.random.default_rng(41)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    target = rng.choice(aa, 566).astype(np.uint8)
    seqs = _protein_batch(rng, target, 30011, min_len, max_len)
    buf, offs = synth.pack(seqs)
    prof = CudaProfiles.new_with_w256([bytes(target)], BLOSUM_62, -10, -1)
    score, status, tier = prof.sw_score_arrays(buf, offs)
    stats = prof.last_stats()
    prof.close()
    pbuf, poff = synth.pack([target])
    c_score, c_status, c_tier = CB.score_batch(pbuf, poff, buf, offs, BLOSUM_62.weights, BLOSUM_62.mapping.index_map, -10, -1)
    assert np.array_equal(status, c_status)
    some = c_status == 0
    assert np.array_equal(score[some], c_score[some]) and np.array_equal(tier[some], c_tier[some])
    assert stats["tier16"] > 100 and stats["tier8"] > 100


def test_offsets_need_not_start_at_zero():
    # a batch may be a window of a larger buffer: offsets[0] != 0 (the shard's offsets are rebased on the device)
    targets, reads = synth.config1(ROOT, n_reads=257)
    buf, offs = synth.fixed_len_batch(reads)
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], W25, -10, -1)
    want = prof.sw_score_arrays(buf, offs)
    junk = np.frombuffer(b"GATTACA" * 11, dtype=np.uint8)
    buf2 = np.concatenate([junk, buf])
    offs2 = (offs + np.uint64(len(junk))).astype(np.uint64)
    got = prof.sw_score_arrays(buf2, offs2)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    a1 = prof.align_arrays(buf, offs)
    a2 = prof.align_arrays(buf2, offs2)
    for k in ("score", "status", "ref_start", "ref_end", "query_start", "query_end", "cigar_off"):
        assert np.array_equal(a1[k], a2[k]), k
    prof.close()


def test_pipelined_batch_equals_one_shot():
    """zoe_cuda_sw_score_batch cuts large batches into sub-batches on two streams (copies overlap kernels); outputs and
    the tier histogram must not depend on it."""
    rng = np.random.default_rng(77)
    target = synth.random_dna(rng, 240)
    n = 300_003
    reads = synth.random_dna(rng, n * 100).reshape(n, 100)
    own = rng.random(n) < 0.5
    starts = rng.integers(0, 140, n)
    idx = starts[:, None] + np.arange(100)[None, :]
    reads[own] = target[idx[own]]
    flip = rng.random(reads.shape) < 0.02
    reads = np.where(flip, synth.random_dna(rng, reads.size).reshape(reads.shape), reads).astype(np.uint8)
    buf, offs = synth.fixed_len_batch(reads)
    prof = CudaProfiles.new_with_w256([bytes(target), bytes(target[:90])], W25, -10, -1)
    os.environ["ZOE_CUDA_PIPELINE"] = "1"  # force it (the library pipelines only when copies are a visible share)
    try:
        a = prof.sw_score_arrays(buf, offs)
        st_a = prof.last_stats()
    finally:
        del os.environ["ZOE_CUDA_PIPELINE"]
    os.environ["ZOE_CUDA_NO_PIPELINE"] = "1"
    try:
        b = prof.sw_score_arrays(buf, offs)
        st_b = prof.last_stats()
    finally:
        del os.environ["ZOE_CUDA_NO_PIPELINE"]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert st_a == st_b and st_a["tier8"] > 1000 and st_a["unmapped"] + st_a["tier8"] + st_a["tier16"] == 2 * n
    sc = osc(W25)
    for i in list(range(0, n, 29989)) + [n - 1, n - 2, 75000, 75001, 150000, 150001, 225001]:
        for j, t in enumerate((target, target[:90])):
            rc, s_, t_ = O.sw_score_from(bytes(t), bytes(reads[i]), sc)
            assert (int(a[1][i, j]), int(a[0][i, j]) if rc == 0 else 0) == (rc, s_ if rc == 0 else 0), (i, j)
    prof.close()
