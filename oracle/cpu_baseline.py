"""ctypes front-end of ``libzoe_cpu.so`` -- the vectorised multi-threaded CPU restatement of zoe's
striped score path.  MEASUREMENT INFRASTRUCTURE ONLY (bench.py cpu_baseline / --impl reference, tests)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzoe_cpu.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "zoe_sw_cpu.cpp")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "libzoe_cpu.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB_PATH)
        _lib.zo_cpu_isa.restype = C.c_char_p
    return _lib


def hardware_threads() -> int:
    return int(lib().zo_cpu_hardware_threads())


def isa() -> str:
    return lib().zo_cpu_isa().decode()


def score_batch(prof_buf, prof_off, reads_buf, reads_off, weights, byte_to_index, gap_open, gap_extend,
                width_bits=256, n_threads=None):
    """Batched ``SharedProfiles::sw_score_from_i8``; returns (score, status, tier) as [n, n_prof] arrays."""
    n_prof = len(prof_off) - 1
    n = len(reads_off) - 1
    if n_threads is None:
        n_threads = hardware_threads()
    w = np.ascontiguousarray(weights, dtype=np.int8)
    lut = np.ascontiguousarray(byte_to_index, dtype=np.uint8)
    score = np.zeros(max(n * n_prof, 1), dtype=np.uint32)
    status = np.zeros(max(n * n_prof, 1), dtype=np.uint8)
    tier = np.zeros(max(n * n_prof, 1), dtype=np.uint8)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    prof_buf = np.ascontiguousarray(prof_buf, dtype=np.uint8)
    prof_off = np.ascontiguousarray(prof_off, dtype=np.uint64)
    reads_buf = np.ascontiguousarray(reads_buf, dtype=np.uint8)
    reads_off = np.ascontiguousarray(reads_off, dtype=np.uint64)
    rc = lib().zo_cpu_score_batch(p(prof_buf, C.c_uint8), p(prof_off, C.c_uint64), C.c_uint32(n_prof),
                                  p(reads_buf, C.c_uint8), p(reads_off, C.c_uint64), C.c_uint64(n),
                                  p(w, C.c_int8), w.shape[0], p(lut, C.c_uint8), int(gap_open), int(gap_extend),
                                  int(width_bits), int(n_threads), p(score, C.c_uint32), p(status, C.c_uint8),
                                  p(tier, C.c_uint8))
    if rc:
        raise ValueError(f"zo_cpu_score_batch failed: {rc}")
    shape = (n, n_prof)
    return score[: n * n_prof].reshape(shape), status[: n * n_prof].reshape(shape), tier[: n * n_prof].reshape(shape)
