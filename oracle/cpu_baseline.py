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


def align_batch(prof_buf, prof_off, reads_buf, reads_off, weights, byte_to_index, gap_open, gap_extend,
                width_bits=256, n_threads=None, streamed_is_query=True, cigar_cap=None):
    """Batched ``SharedProfiles::sw_align_from_i8`` (``SeqSrc::Query(read)`` when ``streamed_is_query``).

    Returns a dict of flat arrays (pair index = read * n_prof + profiled): score, status, tier, ref_start, ref_end,
    query_start, query_end, cigar_off (pairs + 1) and cigar (``len << 4 | op``, op 0 M / 1 I / 2 D / 4 S) -- the layout of
    the product's C ABI, so the two can be compared array against array."""
    n_prof = len(prof_off) - 1
    n = len(reads_off) - 1
    pairs = n * n_prof
    if n_threads is None:
        n_threads = hardware_threads()
    if cigar_cap is None:
        cigar_cap = pairs * 8 + 1024
    w = np.ascontiguousarray(weights, dtype=np.int8)
    lut = np.ascontiguousarray(byte_to_index, dtype=np.uint8)
    out = {k: np.zeros(max(pairs, 1), dtype=np.uint32) for k in ("score", "ref_start", "ref_end", "query_start", "query_end")}
    out.update({k: np.zeros(max(pairs, 1), dtype=np.uint8) for k in ("status", "tier")})
    out["cigar_off"] = np.zeros(pairs + 1, dtype=np.uint64)
    out["cigar"] = np.zeros(max(cigar_cap, 1), dtype=np.uint32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    prof_buf = np.ascontiguousarray(prof_buf, dtype=np.uint8)
    prof_off = np.ascontiguousarray(prof_off, dtype=np.uint64)
    reads_buf = np.ascontiguousarray(reads_buf, dtype=np.uint8)
    reads_off = np.ascontiguousarray(reads_off, dtype=np.uint64)
    for _ in range(2):
        rc = lib().zo_cpu_align_batch(p(prof_buf, C.c_uint8), p(prof_off, C.c_uint64), C.c_uint32(n_prof),
                                      p(reads_buf, C.c_uint8), p(reads_off, C.c_uint64), C.c_uint64(n),
                                      p(w, C.c_int8), w.shape[0], p(lut, C.c_uint8), int(gap_open), int(gap_extend),
                                      int(width_bits), int(n_threads), int(bool(streamed_is_query)),
                                      p(out["score"], C.c_uint32), p(out["status"], C.c_uint8), p(out["tier"], C.c_uint8),
                                      p(out["ref_start"], C.c_uint32), p(out["ref_end"], C.c_uint32),
                                      p(out["query_start"], C.c_uint32), p(out["query_end"], C.c_uint32),
                                      p(out["cigar"], C.c_uint32), p(out["cigar_off"], C.c_uint64),
                                      C.c_uint64(len(out["cigar"])))
        if rc != -6:
            break
        out["cigar"] = np.zeros(int(out["cigar_off"][pairs]) + 16, dtype=np.uint32)  # the needed size is known now
    if rc:
        raise ValueError(f"zo_cpu_align_batch failed: {rc}")
    for k in ("score", "ref_start", "ref_end", "query_start", "query_end", "status", "tier"):
        out[k] = out[k][:pairs]
    out["cigar"] = out["cigar"][: int(out["cigar_off"][pairs])]
    return out


def last_lazy_stats():
    """(lazy-F vector revisits, DP rows) of the last ``score_batch`` call."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib().zo_cpu_last_lazy_stats(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def compare_alignments(got: dict, want: dict, n_pairs: int) -> int:
    """Number of pairs whose (status, score, ranges, CIGAR) differ between two C-ABI-layout result sets.  ``tier`` is
    compared for mapped pairs.  Vectorised: CIGARs are compared through their offsets and a running word mismatch count."""
    gs, ws = np.asarray(got["status"][:n_pairs]), np.asarray(want["status"][:n_pairs])
    bad = gs != ws
    some = (ws == 0) & ~bad
    for k in ("score", "ref_start", "ref_end", "query_start", "query_end", "tier"):
        bad |= some & (np.asarray(got[k][:n_pairs]) != np.asarray(want[k][:n_pairs]))
    go, wo = np.asarray(got["cigar_off"][: n_pairs + 1]).astype(np.int64), np.asarray(want["cigar_off"][: n_pairs + 1]).astype(np.int64)
    glen, wlen = np.diff(go), np.diff(wo)
    bad |= glen != wlen
    ok = ~bad & (wlen > 0)
    if ok.any():
        # gather both CIGAR streams of the comparable pairs and count differing words per pair
        idx = np.nonzero(ok)[0]
        lens = wlen[idx]
        tot = int(lens.sum())
        starts_g, starts_w = go[idx], wo[idx]
        rep = np.repeat(np.arange(len(idx)), lens)
        within = np.arange(tot) - np.repeat(np.cumsum(lens) - lens, lens)
        dg = np.asarray(got["cigar"])[starts_g[rep] + within]
        dw = np.asarray(want["cigar"])[starts_w[rep] + within]
        diff = np.bincount(rep, weights=(dg != dw), minlength=len(idx)) > 0
        bad[idx[diff]] = True
    return int(bad.sum())
