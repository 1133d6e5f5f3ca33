"""GPU parity at scale for protein alignments (config-5 alignments were checked on 70 queries only):
20 000 x 300-aa queries against the 566-aa target, BLOSUM62, every alignment against the vectorised CPU port of zoe's
striped sw_simd_align (pinned to the plain-C oracle on BLOSUM62 in tests/test_cpu_baseline.py)."""
import numpy as np
import pytest

from oracle import cpu_baseline as CB
from zoe_b200 import BLOSUM_62, CudaProfiles, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("go,ge", [(-10, -1), (-4, -1)])
def test_config5_alignments_vs_cpu_port(go, ge):
    targets, q = synth.config5(n_queries=20_000)
    buf, offs = synth.fixed_len_batch(q)
    pbuf, poff = synth.pack([np.asarray(t, dtype=np.uint8) for t in targets])
    want = CB.align_batch(pbuf, poff, buf, offs, BLOSUM_62.weights, BLOSUM_62.mapping.index_map, go, ge, streamed_is_query=True)
    prof = CudaProfiles.new_with_w256([bytes(t) for t in targets], BLOSUM_62, go, ge)
    got = prof.align_arrays(buf, offs)
    stats = prof.last_stats()
    prof.close()
    assert CB.compare_alignments(got, want, len(offs) - 1) == 0, stats
    assert stats["tier16"] > 0 and stats["tier8"] > 0
