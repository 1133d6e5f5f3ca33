"""Batched mirror of zoe's ``sneaky_snake(reference, query, threshold) -> Option<bool>``
(src/alignment/sneaky_snake.rs:78-131) on the GPU: ``zoe_cuda_sneaky_snake_batch``."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

_RESULT = {0: False, 1: True, 2: None}


def _pack(seqs: Sequence[bytes]):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bytes(s) for s in seqs) or b"\0", dtype=np.uint8).copy()
    return buf, offs


class SneakySnake:
    """Owns a CUDA context (``n_devices`` devices; pairs are split by contiguous index range)."""

    def __init__(self, n_devices: int = 1):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.zoe_cuda_create(C.byref(self._h), None, n_devices)
        if rc:
            raise RuntimeError("zoe_cuda_create failed (no usable CUDA device; there is no CPU fallback)")

    def filter_arrays(self, ref_buf, ref_offs, qry_buf, qry_offs, threshold: float) -> np.ndarray:
        """Raw result codes (0 = Some(false), 1 = Some(true), 2 = None) for packed batches."""
        n = len(ref_offs) - 1
        if len(qry_offs) - 1 != n:
            raise ValueError("references and queries must pair up")
        out = np.zeros(max(n, 1), dtype=np.uint8)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
        rc = self._lib.zoe_cuda_sneaky_snake_batch(self._h, p(ref_buf, C.c_uint8), p(ref_offs, C.c_uint64),
                                                   p(qry_buf, C.c_uint8), p(qry_offs, C.c_uint64), n,
                                                   C.c_float(threshold), p(out, C.c_uint8))
        if rc:
            raise RuntimeError(self._lib.zoe_cuda_last_error(self._h).decode(errors="replace"))
        return out[:n]

    def sneaky_snake_batch(self, references: Sequence[bytes], queries: Sequence[bytes], threshold: float) -> List[Optional[bool]]:
        """``out[i] == sneaky_snake(references[i], queries[i], threshold)``."""
        rb, ro = _pack(references)
        qb, qo = _pack(queries)
        return [_RESULT[int(v)] for v in self.filter_arrays(rb, ro, qb, qo, threshold)]

    def last_timing(self):
        """(total ms incl. copies, kernel ms) of the last call, from CUDA events on the library's stream."""
        t, k, n = C.c_float(0), C.c_float(0), C.c_uint32(0)
        self._lib.zoe_cuda_last_timing(self._h, C.byref(t), C.byref(k), C.byref(n))
        return t.value, k.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.zoe_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
