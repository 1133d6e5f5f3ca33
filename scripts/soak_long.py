"""Randomised parity soak of the long-row score path (streamed sequences > 1024 residues, sw_score_long_kernel) and of
mixed batches, against the vectorised CPU port.  Usage: python scripts/soak_long.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import cpu_baseline as CB  # noqa: E402
from zoe_b200 import BLOSUM_62, CudaProfiles, WeightMatrix, synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
t_end = time.time() + budget
cases = pairs = 0
while time.time() < t_end:
    protein = rng.random() < 0.2
    ma, mi = int(rng.integers(1, 6)), -int(rng.integers(1, 7))
    go = -int(rng.integers(0, 13))
    ge = -int(rng.integers(0, -go + 1))
    wm = BLOSUM_62 if protein else WeightMatrix.new_dna_matrix(ma, mi, b"N")
    alpha = AA if protein else ACGT
    max_t = int(rng.choice([200, 3000, 40000]))
    targets = [rng.choice(alpha, int(rng.integers(1, max_t + 1))).astype(np.uint8) for _ in range(int(rng.integers(1, 4)))]
    seqs = []
    for _ in range(int(rng.integers(8, 40))):
        L = int(rng.integers(1, 9000)) if rng.random() < 0.8 else int(rng.integers(1, 1025))
        s = rng.choice(alpha, L).astype(np.uint8)
        t = targets[int(rng.integers(0, len(targets)))]
        if L > 8 and rng.random() < 0.7:
            k = min(L, len(t))
            st = int(rng.integers(0, len(t) - k + 1))
            frag = synth._mutate(rng, t[st:st + k], 0.06, 0.04, 0.04, alpha)
            k2 = min(len(frag), L)
            s[:k2] = frag[:k2]
        seqs.append(s)
    buf, offs = synth.pack(seqs)
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, go, ge)
    score, status, tier = prof.sw_score_arrays(buf, offs)
    prof.close()
    pbuf, poff = synth.pack(targets)
    c_score, c_status, c_tier = CB.score_batch(pbuf, poff, buf, offs, wm.weights, wm.mapping.index_map, go, ge)
    some = c_status == 0
    ok = np.array_equal(status, c_status) and np.array_equal(score[some], c_score[some]) and np.array_equal(tier[some], c_tier[some])
    cases += 1
    pairs += status.size
    print(f"case {cases}: {'BLOSUM62' if protein else 'DNA'} ({ma},{mi},{go},{ge}) targets {[len(t) for t in targets]} reads {len(seqs)} "
          f"max {max(len(s) for s in seqs)} {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = np.argwhere((status != c_status) | (some & ((score != c_score) | (tier != c_tier))))[:5]
        print("first mismatches (read, target):", bad.tolist())
        sys.exit(1)
print(f"soak_long ok: {cases} cases, {pairs} pairs")
