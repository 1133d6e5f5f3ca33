"""A/B timing of the literal striped-emulation kernel: ZOE_CUDA_ALL_EXACT sends every pair through it, so a
one-read batch measures the latency of a single 150 x 1704 pair.  ZOE_LIB=<path> loads another build of the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoe_b200._lib as L  # noqa: E402

if os.environ.get("ZOE_LIB"):
    L.LIB_PATH = os.environ["ZOE_LIB"]
from zoe_b200 import CudaProfiles, WeightMatrix, synth  # noqa: E402

wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
targets, reads = synth.config3(ROOT, n_reads=64, seed=4)
os.environ["ZOE_CUDA_ALL_EXACT"] = "1"
prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
for n in (1, 64):
    buf, offs = synth.fixed_len_batch(reads[:n])
    ts = []
    for _ in range(4):
        out = prof.align_arrays(buf, offs)
        ts.append(prof.last_timing()["total_ms"])
    print(L.LIB_PATH, "n", n, "total_ms", ["%.3f" % t for t in ts], "hazard", int(out["hazard"].sum()),
          "checksum", int(out["cigar"][: int(out["cigar_off"][n])].sum()))
prof.close()
