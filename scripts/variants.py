import json, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from zoe_b200 import CudaProfiles, WeightMatrix, synth
    wm = WeightMatrix.new_dna_matrix(2, -5, b"N")
    targets, reads = synth.config2(n_reads=400000)
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], wm, -10, -1)
    buf, offs = synth.fixed_len_batch(reads)
    cells = reads.size * sum(len(t) for t in targets)
    prof.stage(buf, offs)
    best = 0
    for rep in range(4):
        prof.run_score_staged()
        best = max(best, cells / (prof.last_timing()["dp_kernel_ms"] * 1e-3) / 1e9)
    s, st, t = prof.fetch_scores()
    print(os.environ.get("ZOE_CUDA_SCORE_VARIANT", "default"), os.environ.get("ZOE_CUDA_ONE_STREAM", ""), round(best, 1), int(s.sum()))
    if os.environ.get("PEAKS"):
        for k in (0, 1, 2, 3):
            print("dpx kind", k, prof.dpx_peak(k))
else:
    for env in ({}, {"ZOE_CUDA_ONE_STREAM": "1"}, {"ZOE_CUDA_SCORE_VARIANT": "h"}, {"ZOE_CUDA_SCORE_VARIANT": "p"}, {"ZOE_CUDA_SCORE_VARIANT": "q"}):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e)
