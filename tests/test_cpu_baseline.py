"""CPU-only: the vectorised CPU baseline (oracle/zoe_sw_cpu.cpp) agrees with the plain-C oracle."""
import os

import numpy as np

from oracle import cpu_baseline as CB
from oracle import oracle as O
from zoe_b200 import BLOSUM_62, WeightMatrix, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W25 = WeightMatrix.new_dna_matrix(2, -5, b"N")


def _check(targets, seqs, wm, go, ge, width, lanes, threads):
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    rbuf, roff = synth.pack([np.asarray(s, dtype=np.uint8) for s in seqs])
    score, status, tier = CB.score_batch(pbuf, poff, rbuf, roff, wm.weights, wm.mapping.index_map, go, ge,
                                         width_bits=width, n_threads=threads)
    sc = O.Scoring(wm.weights, wm.mapping.index_map, go, ge)
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            rc, sco, tr = O.sw_score_from(bytes(t), bytes(s), sc, lanes=lanes)
            assert int(status[i, j]) == rc
            if rc == O.SOME:
                assert (int(score[i, j]), int(tier[i, j])) == (sco, tr), (i, j)


def test_cpu_baseline_dna_w256_and_w512():
    targets, reads = synth.config1(ROOT, n_reads=60)
    _check(targets, list(reads), W25, -10, -1, 256, (32, 16, 8), 2)
    _check(targets, list(reads[:30]), W25, -10, -1, 512, (64, 32, 16), 1)
    _check(targets, list(reads[:30]), W25, -10, -1, 128, (16, 8, 4), 3)


def test_cpu_baseline_protein_and_i32():
    targets, q = synth.config5(n_queries=30)
    _check(targets, list(q), BLOSUM_62, -10, -1, 256, (32, 16, 8), 2)
    from zoe_b200 import DNA_PROFILE_MAP
    w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
    a = np.frombuffer(b"A" * 600, dtype=np.uint8)
    _check([a], [a, a[:300], a[:2]], w, -10, -1, 256, (32, 16, 8), 1)
