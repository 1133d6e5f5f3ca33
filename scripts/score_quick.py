"""Scratch GPU check: cfg-2 slice timing of the score kernel (not the bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zoe_b200 import CudaProfiles, WeightMatrix, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
targets, reads = synth.config2(n_reads=n)
prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
buf, offs = synth.fixed_len_batch(reads)
cells = reads.size * sum(len(t) for t in targets)
prof.stage(buf, offs)
best = 0
for rep in range(4):
    prof.run_score_staged()
    best = max(best, cells / (prof.last_timing()["dp_kernel_ms"] * 1e-3) / 1e9)
s, st, t = prof.fetch_scores()
print("cfg2 slice", n, "reads: best kernel GCUPS", round(best, 1), "checksum", int(s.sum()), "dpx", prof.dpx_peak(0)[0])
