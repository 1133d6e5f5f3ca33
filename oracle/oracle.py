"""ctypes front-end of ``libzoe_oracle.so`` (the plain-C restatement of zoe's striped SW).

TEST INFRASTRUCTURE ONLY -- see ``zoe_sw_oracle.c``.  The scoring arguments are raw arrays so
that this file has no dependency on the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzoe_oracle.so")

SOME, OVERFLOWED, UNMAPPED = 0, 1, 2
STATUS_NAMES = {0: "Some", 1: "Overflowed", 2: "Unmapped"}
ERR_EMPTY_SEQUENCE, ERR_GAP_OPEN_RANGE, ERR_GAP_EXTEND_RANGE, ERR_BAD_GAP_WEIGHTS = -1, -2, -3, -4


class _Scoring(C.Structure):
    _fields_ = [
        ("weights", C.POINTER(C.c_int8)),
        ("S", C.c_int),
        ("map", C.POINTER(C.c_uint8)),
        ("gap_open", C.c_int),
        ("gap_extend", C.c_int),
    ]


class _Alignment(C.Structure):
    _fields_ = [
        ("score", C.c_uint32),
        ("ref_start", C.c_uint64),
        ("ref_end", C.c_uint64),
        ("query_start", C.c_uint64),
        ("query_end", C.c_uint64),
        ("ref_len", C.c_uint64),
        ("query_len", C.c_uint64),
        ("n_ops", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle (and the vectorised CPU baseline) with gcc; returns the .so path."""
    srcs = [os.path.join(_HERE, "zoe_sw_oracle.c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "libzoe_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


@dataclass
class Scoring:
    """Raw scoring inputs: ``weights`` is S*S row-major ``weights[ref_idx][query_idx]``."""

    weights: np.ndarray
    byte_to_index: np.ndarray
    gap_open: int
    gap_extend: int

    def __post_init__(self):
        self.weights = np.ascontiguousarray(self.weights, dtype=np.int8)
        self.byte_to_index = np.ascontiguousarray(self.byte_to_index, dtype=np.uint8)
        assert self.byte_to_index.shape == (256,)
        self.S = self.weights.shape[0]

    def c(self) -> _Scoring:
        return _Scoring(
            self.weights.ctypes.data_as(C.POINTER(C.c_int8)),
            self.S,
            self.byte_to_index.ctypes.data_as(C.POINTER(C.c_uint8)),
            int(self.gap_open),
            int(self.gap_extend),
        )


@dataclass
class Aln:
    """Mirror of zoe's ``Alignment<u32>`` (output.rs:264-279) with the CIGAR as a string."""

    score: int
    ref_range: tuple
    query_range: tuple
    cigar: str
    ref_len: int
    query_len: int


def _buf(b: bytes):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


def _mk_aln(a: _Alignment, ops, lens) -> Aln:
    cigar = "".join(f"{lens[i]}{chr(ops[i])}" for i in range(a.n_ops))
    return Aln(a.score, (a.ref_start, a.ref_end), (a.query_start, a.query_end), cigar, a.ref_len, a.query_len)


def validate_profile_args(seq_len: int, gap_open: int, gap_extend: int) -> int:
    return lib().zo_validate_profile_args(C.c_uint64(seq_len), gap_open, gap_extend)


def striped_score(profiled: bytes, streamed: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True):
    """``StripedProfile::<T,N,S>::new(profiled,..).sw_score(streamed)`` -> (status, score)."""
    score = C.c_uint32(0)
    s = sc.c()
    rc = lib().zo_striped_score(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                                C.byref(s), bits, int(signed), lanes, C.byref(score))
    return rc, score.value


def striped_score_ends(profiled: bytes, streamed: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True):
    """``sw_score_ends(SeqSrc::Reference(streamed))`` -> (status, score, ref_end, query_end)."""
    score, re, qe = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    s = sc.c()
    rc = lib().zo_striped_score_ends(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed),
                                     C.c_uint64(len(streamed)), C.byref(s), bits, int(signed), lanes,
                                     C.byref(score), C.byref(re), C.byref(qe))
    return rc, score.value, re.value, qe.value


def _cap(profiled, streamed):
    return 2 * (len(profiled) + len(streamed)) + 8


def striped_align(profiled: bytes, streamed: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True,
                  streamed_is_query: bool = False):
    """``StripedProfile::sw_align(SeqSrc)``. ``streamed_is_query`` <=> ``SeqSrc::Query(streamed)``."""
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    s = sc.c()
    rc = lib().zo_striped_align(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                                C.byref(s), bits, int(signed), lanes, int(streamed_is_query), C.byref(a), ops, lens, cap)
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None)


def scalar_score(profiled: bytes, streamed: bytes, sc: Scoring):
    score = C.c_uint32(0)
    s = sc.c()
    rc = lib().zo_scalar_score(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                               C.byref(s), C.byref(score))
    return rc, score.value


def scalar_align(profiled: bytes, streamed: bytes, sc: Scoring, streamed_is_query: bool = False):
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    s = sc.c()
    rc = lib().zo_scalar_align(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                               C.byref(s), int(streamed_is_query), C.byref(a), ops, lens, cap)
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None)


def sw_score_from(profiled: bytes, streamed: bytes, sc: Scoring, lanes=(32, 16, 8), first_bits: int = 8):
    """``ProfileSets::sw_score_from_i{8,16,32}`` -> (status, score, tier)."""
    score, tier = C.c_uint32(0), C.c_int(0)
    s = sc.c()
    rc = lib().zo_sw_score_from(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                                C.byref(s), first_bits, lanes[0], lanes[1], lanes[2], C.byref(score), C.byref(tier))
    return rc, score.value, tier.value


def sw_align_from(profiled: bytes, streamed: bytes, sc: Scoring, lanes=(32, 16, 8), first_bits: int = 8,
                  streamed_is_query: bool = False):
    """``ProfileSets::sw_align_from_i{8,16,32}(SeqSrc)`` -> (status, Aln | None, tier)."""
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    tier = C.c_int(0)
    s = sc.c()
    rc = lib().zo_sw_align_from(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                                C.byref(s), first_bits, lanes[0], lanes[1], lanes[2], int(streamed_is_query),
                                C.byref(a), ops, lens, cap, C.byref(tier))
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None), tier.value


def striped_score_ranges(profiled: bytes, streamed: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True,
                         streamed_is_query: bool = False):
    """``StripedProfile::sw_score_ranges(SeqSrc)`` -> (status, score, ref_range, query_range)."""
    score = C.c_uint32(0)
    v = [C.c_uint64(0) for _ in range(4)]
    s = sc.c()
    rc = lib().zo_striped_score_ranges(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed),
                                       C.c_uint64(len(streamed)), C.byref(s), bits, int(signed), lanes,
                                       int(streamed_is_query), C.byref(score), *[C.byref(x) for x in v])
    return rc, score.value, (v[0].value, v[1].value), (v[2].value, v[3].value)


def sw_score_ranges_from(profiled: bytes, streamed: bytes, sc: Scoring, lanes=(32, 16, 8), first_bits: int = 8,
                         streamed_is_query: bool = False):
    """``ProfileSets::sw_score_ranges_from_i{8,16,32}(SeqSrc)`` -> (status, score, ref_range, query_range, tier)."""
    score, tier = C.c_uint32(0), C.c_int(0)
    v = [C.c_uint64(0) for _ in range(4)]
    s = sc.c()
    rc = lib().zo_sw_score_ranges_from(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed),
                                       C.c_uint64(len(streamed)), C.byref(s), first_bits, lanes[0], lanes[1], lanes[2],
                                       int(streamed_is_query), C.byref(score), *[C.byref(x) for x in v], C.byref(tier))
    return rc, score.value, (v[0].value, v[1].value), (v[2].value, v[3].value), tier.value


def striped_profile(profiled: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True) -> np.ndarray:
    """The striped profile as an ``[S, nv, N]`` int32 array (profile.rs:270-306)."""
    nv = C.c_int(0)
    cap = sc.S * (len(profiled) + lanes) * 1
    out = np.zeros(cap, dtype=np.int32)
    s = sc.c()
    rc = lib().zo_striped_profile_dump(_buf(profiled), C.c_uint64(len(profiled)), C.byref(s), bits, int(signed), lanes,
                                       out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint64(cap), C.byref(nv))
    if rc:
        raise ValueError(f"profile error {rc}")
    return out[: sc.S * nv.value * lanes].reshape(sc.S, nv.value, lanes)


def banded_align(profiled: bytes, streamed: bytes, sc: Scoring, band_width: int):
    """``sw_banded_align(streamed, &ScalarProfile::new(profiled,..), band_width)`` (banded.rs:40-133)."""
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    s = sc.c()
    rc = lib().zo_banded_align(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed), C.c_uint64(len(streamed)),
                               C.byref(s), C.c_uint64(band_width), C.byref(a), ops, lens, cap)
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None)


def striped_align_3pass(profiled: bytes, streamed: bytes, sc: Scoring, bits: int, lanes: int, signed: bool = True,
                        streamed_is_query: bool = False):
    """``StripedProfile::sw_align_3pass(SeqSrc, ..)`` (three_pass.rs:21-104) -> (status, Aln | None, path)."""
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    path = C.c_int(-1)
    s = sc.c()
    rc = lib().zo_striped_align_3pass(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed),
                                      C.c_uint64(len(streamed)), C.byref(s), bits, int(signed), lanes,
                                      int(streamed_is_query), C.byref(a), ops, lens, cap, C.byref(path))
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None), path.value


def sw_align_3pass_from(profiled: bytes, streamed: bytes, sc: Scoring, lanes=(32, 16, 8), first_bits: int = 8,
                        streamed_is_query: bool = False):
    """``ProfileSets::sw_align_from_i{8,16,32}_3pass(SeqSrc)`` -> (status, Aln | None, tier, path)."""
    cap = _cap(profiled, streamed)
    ops = (C.c_uint8 * cap)()
    lens = (C.c_uint32 * cap)()
    a = _Alignment()
    tier, path = C.c_int(0), C.c_int(-1)
    s = sc.c()
    rc = lib().zo_sw_align_3pass_from(_buf(profiled), C.c_uint64(len(profiled)), _buf(streamed),
                                      C.c_uint64(len(streamed)), C.byref(s), first_bits, lanes[0], lanes[1], lanes[2],
                                      int(streamed_is_query), C.byref(a), ops, lens, cap, C.byref(tier), C.byref(path))
    return rc, (_mk_aln(a, ops, lens) if rc == SOME else None), tier.value, path.value


def sneaky_snake(reference: bytes, query: bytes, threshold: float):
    """``sneaky_snake(reference, query, threshold)`` (sneaky_snake.rs:78-131) -> True / False / None."""
    f = lib().zo_sneaky_snake
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_float]
    rb, qb = _buf(reference), _buf(query)
    rc = f(C.cast(rb, C.c_void_p), len(reference), C.cast(qb, C.c_void_p), len(query), C.c_float(threshold))
    return {0: False, 1: True, 2: None}[rc]
