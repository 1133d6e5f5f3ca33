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


def _check_align(targets, seqs, wm, go, ge, width, lanes, threads, streamed_is_query=True):
    """The vectorised sw_simd_align port == the plain-C oracle, field by field and CIGAR by CIGAR."""
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    rbuf, roff = synth.pack([np.asarray(s, dtype=np.uint8) for s in seqs])
    out = CB.align_batch(pbuf, poff, rbuf, roff, wm.weights, wm.mapping.index_map, go, ge, width_bits=width,
                         n_threads=threads, streamed_is_query=streamed_is_query)
    sc = O.Scoring(wm.weights, wm.mapping.index_map, go, ge)
    n_prof = len(targets)
    for i, s in enumerate(seqs):
        for j, t in enumerate(targets):
            k = i * n_prof + j
            rc, want, tr = O.sw_align_from(bytes(t), bytes(s), sc, lanes=lanes, streamed_is_query=streamed_is_query)
            assert int(out["status"][k]) == rc, (i, j)
            lo, hi = int(out["cigar_off"][k]), int(out["cigar_off"][k + 1])
            if rc != O.SOME:
                assert lo == hi
                continue
            cig = "".join(f"{int(w) >> 4}{'MID?S'[int(w) & 15]}" for w in out["cigar"][lo:hi])
            got = (int(out["score"][k]), int(out["ref_start"][k]), int(out["ref_end"][k]), int(out["query_start"][k]),
                   int(out["query_end"][k]), cig, int(out["tier"][k]))
            assert got == (want.score, want.ref_range[0], want.ref_range[1], want.query_range[0], want.query_range[1],
                           want.cigar, tr), (i, j, got, want)


def test_cpu_align_port_dna_all_widths_both_orientations():
    targets, reads = synth.config3(ROOT, n_reads=80)
    _check_align(targets, list(reads), W25, -10, -1, 256, (32, 16, 8), 2)
    _check_align(targets, list(reads[:40]), W25, -10, -1, 512, (64, 32, 16), 1)
    _check_align(targets, list(reads[:40]), W25, -10, -1, 128, (16, 8, 4), 3)
    _check_align(targets, list(reads[:40]), W25, -10, -1, 256, (32, 16, 8), 2, streamed_is_query=False)


def test_cpu_align_port_tie_heavy_scorings_and_protein():
    # small weights and cheap gaps produce E == H == F ties, where zoe's lazy-F pass decides the CIGAR by lane layout
    rng = np.random.default_rng(5)
    w = WeightMatrix.new_dna_matrix(4, -2, b"N")
    targets = [synth.random_dna(rng, 180), synth.random_dna(rng, 75)]
    seqs = [synth.random_dna(rng, int(k)) for k in rng.integers(1, 90, 60)]
    seqs += [targets[0][10:70].copy(), targets[1][5:60].copy(), np.zeros(0, dtype=np.uint8)]
    for go, ge in ((-3, -1), (-2, -2), (0, 0), (-5, 0)):
        _check_align(targets, seqs, w, go, ge, 256, (32, 16, 8), 2)
        _check_align(targets, seqs[:20], w, go, ge, 128, (16, 8, 4), 1)
    tp, q = synth.config5(n_queries=12)
    _check_align(tp, list(q), BLOSUM_62, -10, -1, 256, (32, 16, 8), 2)


def test_compare_alignments_counts_differences():
    targets, reads = synth.config3(ROOT, n_reads=40)
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    rbuf, roff = synth.fixed_len_batch(reads)
    a = CB.align_batch(pbuf, poff, rbuf, roff, W25.weights, W25.mapping.index_map, -10, -1, n_threads=2)
    b = {k: v.copy() for k, v in a.items()}
    assert CB.compare_alignments(a, b, 40) == 0
    mapped = np.nonzero(a["status"] == 0)[0]
    b["score"][mapped[0]] += 1
    b["cigar"][int(a["cigar_off"][mapped[1]])] ^= 16
    b["status"][mapped[2]] = 2
    assert CB.compare_alignments(a, b, 40) == 3
