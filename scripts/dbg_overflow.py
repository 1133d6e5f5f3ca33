import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoe_b200 import CudaProfiles, WeightMatrix, DNA_PROFILE_MAP, SeqSrc
w = WeightMatrix.new(DNA_PROFILE_MAP, 127, -5, b"N")
prof = CudaProfiles.new_with_w256([b"A" * 600], w, -10, -1)
try:
    r = prof.sw_align_batch(SeqSrc.Query([b"A" * 600, b"A" * 300, b"A" * 100 + b"C" + b"A" * 100]))
    print(r)
except Exception as e:
    print("ERR", e)
print(prof.last_stats())
