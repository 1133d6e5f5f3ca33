"""Randomised parity soak (GPU box): random scorings, lane presets, orientations, panels and ragged batches through
every batch entry point (score, align, ranges, 3-pass), element-wise against the plain-C oracle.  Prints one line per
case and a summary; exits non-zero on the first mismatch.  Usage: python scripts/soak.py [seconds] [seed]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from zoe_b200 import BLOSUM_62, CudaProfiles, SeqSrc, WeightMatrix, synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
t_end = time.time() + budget
cases = pairs_checked = 0
pool = ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 32)))
while time.time() < t_end:
    ma, mi = int(rng.integers(1, 6)), -int(rng.integers(1, 7))
    go = -int(rng.integers(0, 13))
    ge = -int(rng.integers(0, -go + 1))
    protein = rng.random() < 0.3
    wm = BLOSUM_62 if protein else WeightMatrix.new_dna_matrix(ma, mi, b"N")
    alpha = AA if protein else ACGT
    rand_seq = lambda L: rng.choice(alpha, L).astype(np.uint8)  # noqa: E731
    lanes = [(16, 8, 4), (32, 16, 8), (64, 32, 16)][int(rng.integers(0, 3))]
    pq = bool(rng.integers(0, 2))
    max_t = int(rng.choice([60, 300, 1500, 1500, 30000]))
    max_r = int(rng.choice([40, 150, 400, 1024, 1024, 2600]))  # beyond 1024: the long-row ranges / 3-pass / align paths
    targets = [rand_seq(int(rng.integers(1, max_t + 1))) for _ in range(int(rng.integers(1, 7 if max_t < 30000 else 5)))]
    seqs = []
    for _ in range(int(rng.integers(20, 60))):
        L = int(rng.integers(0 if rng.random() < 0.05 else 1, max_r + 1))
        s = rand_seq(L)
        t = targets[int(rng.integers(0, len(targets)))]
        if L > 8 and rng.random() < 0.7:
            k = min(L, len(t))
            st = int(rng.integers(0, len(t) - k + 1))
            frag = synth._mutate(rng, t[st:st + k], float(rng.choice([0.0, 0.02, 0.1])), float(rng.choice([0.0, 0.01, 0.05])),
                                 float(rng.choice([0.0, 0.01, 0.05])), alpha)
            k2 = min(len(frag), L)
            s[:k2] = frag[:k2]
        seqs.append(s)
    tb = [bytes(t) for t in targets]
    sb = [bytes(s) for s in seqs]
    sc = O.Scoring(wm.weights, wm.mapping.index_map, go, ge)
    # which integer types may answer: the from_i8 / _i16 / _i32 chains, or one standalone StripedProfile<T, N, S>
    policy = [None, None, None, (16, 32, False), (32, 32, False), (8, 8, False), (8, 8, True), (16, 16, False),
              (16, 16, True)][int(rng.integers(0, 9))]
    standalone = policy is not None and policy[0] == policy[1]
    first_bits = policy[0] if policy else 8
    if standalone:
        n_lanes = lanes[{8: 0, 16: 1, 32: 2}[policy[0]]]
        lanes = (n_lanes, n_lanes, n_lanes)
    prof = CudaProfiles(tb, wm, go, ge, lanes=lanes, profiled_is_query=pq)
    if policy:
        prof.set_width_policy(*policy)
    src = SeqSrc.Reference(sb) if pq else SeqSrc.Query(sb)
    g_score = prof.sw_score_batch(sb)
    g_align = prof.sw_align_batch(src)
    g_rng = prof.sw_score_ranges_batch(src)
    g_3p = prof.sw_align_3pass_batch(src)
    prof.close()

    def check(ij):
        i, j = ij
        t, s = tb[j], sb[i]
        if standalone:
            bits, signed, nl, inv = policy[0], not policy[2], lanes[0], not pq
            rc, score = O.striped_score(t, s, sc, bits, nl, signed=signed)
            r_al = O.striped_align(t, s, sc, bits, nl, signed=signed, streamed_is_query=inv)
            r_rg = O.striped_score_ranges(t, s, sc, bits, nl, signed=signed, streamed_is_query=inv)
            r_3p = O.striped_align_3pass(t, s, sc, bits, nl, signed=signed, streamed_is_query=inv)[:2]
        else:
            rc, score, _tier = O.sw_score_from(t, s, sc, lanes=lanes, first_bits=first_bits)
            r_al = O.sw_align_from(t, s, sc, lanes=lanes, first_bits=first_bits, streamed_is_query=not pq)[:2]
            r_rg = O.sw_score_ranges_from(t, s, sc, lanes=lanes, first_bits=first_bits, streamed_is_query=not pq)[:4]
            r_3p = O.sw_align_3pass_from(t, s, sc, lanes=lanes, first_bits=first_bits, streamed_is_query=not pq)[:2]
        g = g_score[i][j]
        assert g.status.value == rc and (rc != 0 or g.unwrap() == score), ("score", i, j, g, rc, score)
        rc, aln = r_al
        g = g_align[i][j]
        assert g.status.value == rc, ("align", i, j, g, rc)
        if rc == 0:
            a = g.unwrap()
            assert (a.score, a.ref_range, a.query_range, a.states) == (aln.score, aln.ref_range, aln.query_range, aln.cigar), ("align", i, j, a, aln)
        rc, score, rr, qr = r_rg
        g = g_rng[i][j]
        assert g.status.value == rc, ("ranges", i, j, g, rc)
        if rc == 0:
            a = g.unwrap()
            assert (a.score, a.ref_range, a.query_range) == (score, rr, qr), ("ranges", i, j, a, score, rr, qr)
        rc, aln = r_3p
        g = g_3p[i][j]
        assert g.status.value == rc, ("3pass", i, j, g, rc)
        if rc == 0:
            a = g.unwrap()
            assert (a.score, a.ref_range, a.query_range, a.states) == (aln.score, aln.ref_range, aln.query_range, aln.cigar), ("3pass", i, j, a, aln)
        return 1

    ij = [(i, j) for i in range(len(sb)) for j in range(len(tb))]
    try:
        pairs_checked += sum(pool.map(check, ij))
    except AssertionError as e:
        print("MISMATCH", dict(match=ma, mismatch=mi, go=go, ge=ge, lanes=lanes, profiled_is_query=pq), e.args[0][:3], flush=True)
        print(repr(e.args[0])[:2000])
        sys.exit(1)
    cases += 1
    print(f"case {cases}: {'BLOSUM62' if protein else 'DNA'} ({ma},{mi},{go},{ge}) lanes {lanes} policy {policy} pq {pq} targets {[len(t) for t in tb]} reads {len(sb)} max {max(len(s) for s in sb)} ok",
          flush=True)
print(f"soak ok: {cases} cases, {pairs_checked} pairs x 4 entry points, seed {seed}")
