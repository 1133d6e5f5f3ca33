"""CPU-only: SAM emission mirrors SamData::from_alignment / unmapped / Display (records/sam/mod.rs:201-245,
std_traits.rs:3-43)."""
import numpy as np

from zoe_b200 import Alignment, MaybeAligned
from zoe_b200.sam import SamData, sam_lines_from_arrays, sam_records


def test_from_alignment_fields_and_line():
    # types/test.rs:14-38: ref 3..15, query 1..13, 1S5M1D4M1I2M3S, score from the (4,-2,-3,-1) example
    aln = Alignment(27, (3, 15), (1, 13), "1S5M1D4M1I2M3S", 16, 16)
    rec = SamData.from_alignment(aln, "read1", 0, "ref1", 255, b"TCTCAGATTGCAGTTT", b"IIIIIIIIIIIIIIII")
    assert (rec.pos, rec.cigar, rec.rnext, rec.pnext, rec.tlen, rec.opt_fields) == (4, "1S5M1D4M1I2M3S", "*", 0, 0, ["AS:i:27"])
    assert str(rec) == "read1\t0\tref1\t4\t255\t1S5M1D4M1I2M3S\t*\t0\t0\tTCTCAGATTGCAGTTT\tIIIIIIIIIIIIIIII\tAS:i:27"


def test_unmapped_record():
    assert str(SamData.unmapped("r", "ref")) == "r\t4\tref\t0\t255\t*\t*\t0\t0\t*\t*"


def test_batch_helpers_agree():
    res = [[MaybeAligned.some(Alignment(20, (0, 10), (0, 10), "10M", 10, 10)), MaybeAligned.Unmapped],
           [MaybeAligned.Overflowed, MaybeAligned.some(Alignment(2, (0, 1), (3, 4), "3S1M", 1, 4))]]
    seqs = [b"ACGTACGTAC", b"TTTA"]
    recs = [str(r) for r in sam_records(res, ["q0", "q1"], ["t0", "t1"], seqs)]
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    offs = np.array([0, 10, 14], dtype=np.uint64)
    a = {"status": np.array([0, 2, 1, 0], dtype=np.uint8), "score": np.array([20, 0, 0, 2], dtype=np.uint32),
         "ref_start": np.array([0, 0, 0, 0], dtype=np.uint32),
         "cigar": np.array([(10 << 4) | 0, (3 << 4) | 4, (1 << 4) | 0], dtype=np.uint32),
         "cigar_off": np.array([0, 1, 1, 1, 3], dtype=np.uint64)}
    lines = list(sam_lines_from_arrays(a, 2, ["q0", "q1"], ["t0", "t1"], buf, offs))
    assert lines == recs
    assert lines[0] == "q0\t0\tt0\t1\t255\t10M\t*\t0\t0\tACGTACGTAC\t*\tAS:i:20"
    assert lines[3].split("\t")[5] == "3S1M"


def test_fastq_ingestion_to_packed_batch():
    import pytest
    from zoe_b200.fastq import pack_fastq, read_fastq
    data = b"@r1 desc\nACGT\n+\nIIII\n@r2\nTTTTTT\n+r2\n!!!!!!\r\n"
    names, buf, offs, quals = pack_fastq(data)
    assert names == ["r1 desc", "r2"] and bytes(buf) == b"ACGTTTTTTT" and offs.tolist() == [0, 4, 10]
    assert quals == [b"IIII", b"!!!!!!"]
    for bad, msg in [(b"r1\nACGT\n+\nIIII\n", "Missing '@'"), (b"@\nACGT\n+\nIIII\n", "Missing FASTQ header"),
                     (b"@r\n\n+\nIIII\n", "Missing FASTQ sequence"), (b"@r\nACGT\nIIII\n", r"Missing '\+' line"),
                     (b"@r\nACGT\n+\n\n", "Missing FASTQ quality"), (b"@r\nACGT\n+\nII\n", "length mismatch")]:
        with pytest.raises(ValueError, match=msg):
            list(read_fastq(bad))
