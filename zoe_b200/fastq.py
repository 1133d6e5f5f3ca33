"""FASTQ ingestion into the packed batch layout the C ABI takes (concatenated bytes + uint64 offsets).

Mirrors zoe's ``FastQReader`` (src/data/records/fastq/reader.rs:87-185): single-line records only, ``@`` header,
non-empty sequence, ``+`` separator, quality string of the sequence's length made of graphic ASCII; the same
conditions are errors here (``ValueError`` with zoe's messages).
"""
from __future__ import annotations

import io
from typing import BinaryIO, Iterable, Iterator, List, Tuple, Union

import numpy as np


def _chop(line: bytes) -> bytes:
    if line.endswith(b"\n"):
        line = line[:-1]
    if line.endswith(b"\r"):
        line = line[:-1]
    return line


def read_fastq(src: Union[str, bytes, BinaryIO, Iterable[bytes]]) -> Iterator[Tuple[str, bytes, bytes]]:
    """Yields ``(header, sequence, quality)`` per record."""
    if isinstance(src, str):
        fh: Iterable[bytes] = open(src, "rb")
    elif isinstance(src, (bytes, bytearray)):
        fh = io.BytesIO(bytes(src))
    else:
        fh = src
    it = iter(fh)
    for first in it:
        if not first.startswith(b"@"):
            raise ValueError("Missing '@' symbol at header line beginning! Ensure that the FASTQ file is not multi-line.")
        header = _chop(first[1:])
        if not header:
            raise ValueError("Missing FASTQ header!")
        name = header.decode("utf-8")
        seq = _chop(next(it, b""))
        if not seq:
            raise ValueError(f"Missing FASTQ sequence! See header: {name}")
        plus = next(it, b"")
        if not plus.startswith(b"+"):
            raise ValueError(f"Missing '+' line! Ensure that the FASTQ file is not multi-line. See header: {name}")
        qual = _chop(next(it, b""))
        if len(qual) != len(seq):
            if not qual:
                raise ValueError(f"Missing FASTQ quality scores! See header: {name}")
            raise ValueError(f"Sequence and quality score length mismatch ({len(seq)} ≠ {len(qual)})! See: {name}")
        if any(c < 33 or c > 126 for c in qual):
            raise ValueError(f"Quality scores must be graphic ASCII! See: {name}")
        yield name, seq, qual


def pack_fastq(src) -> Tuple[List[str], np.ndarray, np.ndarray, List[bytes]]:
    """``(headers, buf, offs, quals)``: the streamed-batch layout of ``zoe_cuda_sw_*_batch``."""
    names, seqs, quals = [], [], []
    for name, seq, qual in read_fastq(src):
        names.append(name)
        seqs.append(seq)
        quals.append(qual)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(1, dtype=np.uint8)
    return names, np.ascontiguousarray(buf), offs, quals
