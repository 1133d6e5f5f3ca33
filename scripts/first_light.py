"""Scratch GPU check: DPX issue-rate microbenchmarks + a timed config-2 slice (not the bench)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zoe_b200 import CudaProfiles, WeightMatrix, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
targets, reads = synth.config2(n_reads=n_reads)
prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
out = {}
for kind in (0, 1):
    g, ms = prof.dpx_peak(kind)
    out[f"dpx_peak_kind{kind}_Glane_instr_s"] = g
    out[f"dpx_peak_kind{kind}_ms"] = ms
buf, offs = synth.fixed_len_batch(reads)
cells = reads.size * sum(len(t) for t in targets)
for rep in range(3):
    t0 = time.time()
    score, status, tier = prof.sw_score_arrays(buf, offs)
    dt = time.time() - t0
    tm = prof.last_timing()
    out[f"e2e_rep{rep}"] = {"wall_s": dt, **tm, "gcups_e2e": cells / dt / 1e9,
                            "gcups_kernel": cells / (tm["dp_kernel_ms"] * 1e-3) / 1e9}
prof.stage(buf, offs)
for rep in range(3):
    prof.run_score_staged()
    tm = prof.last_timing()
    out[f"staged_rep{rep}"] = {**tm, "gcups_kernel": cells / (tm["dp_kernel_ms"] * 1e-3) / 1e9}
out["stats"] = prof.last_stats()
out["score_hist"] = np.bincount(np.minimum(score.reshape(-1), 310) // 31).tolist()
print(json.dumps(out, indent=1))
