"""Scratch GPU check: timed config-3 slice (align with traceback)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zoe_b200 import CudaProfiles, WeightMatrix, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
targets, reads = synth.config3(ROOT, n_reads=n_reads)
prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
buf, offs = synth.fixed_len_batch(reads)
cells = reads.size * sum(len(t) for t in targets)
out = {}
prof.stage(buf, offs)
for rep in range(3):
    t0 = time.time()
    prof.run_align_staged()
    dt = time.time() - t0
    tm = prof.last_timing()
    out[f"staged_rep{rep}"] = {"wall_s": dt, **tm, "gcups_total": cells / (tm["total_ms"] * 1e-3) / 1e9,
                               "gcups_fill": cells / (tm["dp_kernel_ms"] * 1e-3) / 1e9}
out["stats"] = prof.last_stats()
t0 = time.time()
a = prof.align_arrays(buf, offs)
out["e2e"] = {"wall_s": time.time() - t0, **prof.last_timing(), "cigar_words": int(a["cigar_off"][-1]),
              "gcups_e2e": cells / (time.time() - t0) / 1e9, "hazard": int(a["hazard"].sum())}
print(json.dumps(out, indent=1))
import ctypes as C
for rep in range(3):
    t0 = time.time()
    a = prof.align_arrays(buf, offs)
    t1 = time.time()
    print("align_arrays rep", rep, "wall_ms", (t1 - t0) * 1e3, prof.last_timing(), file=sys.stderr)
