"""Host-side mirror of zoe's interface for the SW path, over the C-ABI CUDA library.

zoe names mirrored here (paths relative to the zoe repository):
  * ``ProfileError``   -- src/alignment/errors.rs:6-15 (raised by validate_profile_args,
                          src/alignment/profile.rs:32-44)
  * ``MaybeAligned``   -- src/alignment/types/output.rs:18-25 (Some / Overflowed / Unmapped)
  * ``Alignment``      -- src/alignment/types/output.rs:264-279
  * ``SeqSrc``         -- src/alignment/mod.rs:157-162
  * ``CudaProfiles``   -- the batched counterpart of ``SharedProfiles``
                          (src/alignment/profile_set.rs:552-560) with ``new_with_w128/256/512``
                          (profile_set.rs:434-483) and ``sw_score_from_i8`` / ``sw_align_from_i8``
                          semantics (profile_set.rs:71-78, 136-145), applied to a whole batch.

All arithmetic happens in the CUDA library; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import Enum
from typing import Generic, Iterable, Optional, Sequence, TypeVar

import numpy as np

from . import _lib
from .matrices import WeightMatrix

T = TypeVar("T")


class ProfileError(ValueError):
    """zoe's ``ProfileError`` (errors.rs:6-15)."""

    EmptySequence = "EmptySequence"
    GapOpenOutOfRange = "GapOpenOutOfRange"
    GapExtendOutOfRange = "GapExtendOutOfRange"
    BadGapWeights = "BadGapWeights"

    def __init__(self, kind: str, message: str = ""):
        super().__init__(f"{kind}: {message}" if message else kind)
        self.kind = kind


class ZoeCudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"zoe_cuda error {code}: {message}")
        self.code = code


_PROFILE_ERRORS = {
    _lib.E_EMPTY_SEQUENCE: ProfileError.EmptySequence,
    _lib.E_GAP_OPEN_RANGE: ProfileError.GapOpenOutOfRange,
    _lib.E_GAP_EXTEND_RANGE: ProfileError.GapExtendOutOfRange,
    _lib.E_BAD_GAP_WEIGHTS: ProfileError.BadGapWeights,
}


class Status(Enum):
    Some = 0
    Overflowed = 1
    Unmapped = 2


@dataclass(frozen=True)
class MaybeAligned(Generic[T]):
    """Tri-state alignment outcome (output.rs:18-25)."""

    status: Status
    value: Optional[T] = None

    @classmethod
    def some(cls, v: T) -> "MaybeAligned[T]":
        return cls(Status.Some, v)

    def is_some(self) -> bool:
        return self.status is Status.Some

    def unwrap(self) -> T:
        if self.status is not Status.Some:
            raise ValueError(f"called unwrap() on MaybeAligned::{self.status.name}")
        return self.value  # type: ignore[return-value]


MaybeAligned.Overflowed = MaybeAligned(Status.Overflowed)  # type: ignore[attr-defined]
MaybeAligned.Unmapped = MaybeAligned(Status.Unmapped)  # type: ignore[attr-defined]


@dataclass(frozen=True)
class SeqSrc(Generic[T]):
    """Which role the *non-profile* sequences play (alignment/mod.rs:157-162)."""

    kind: str
    seq: T

    @classmethod
    def Query(cls, seq: T) -> "SeqSrc[T]":
        return cls("Query", seq)

    @classmethod
    def Reference(cls, seq: T) -> "SeqSrc[T]":
        return cls("Reference", seq)


@dataclass(frozen=True)
class Alignment:
    """zoe's ``Alignment<u32>`` (output.rs:264-279); ``states`` rendered as a CIGAR string."""

    score: int
    ref_range: tuple
    query_range: tuple
    states: str
    ref_len: int
    query_len: int


@dataclass(frozen=True)
class ScoreAndRanges:
    """zoe's ``ScoreAndRanges<u32>`` (src/alignment/types/output.rs): score + 0-based half-open ranges."""

    score: int
    ref_range: tuple
    query_range: tuple


_CIGAR_OPS = {0: "M", 1: "I", 2: "D", 4: "S"}


def _pack(seqs: Sequence[bytes]):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(buf), offs


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class CudaProfiles:
    """A set of profiled sequences resident on one or more B200s.

    ``CudaProfiles.new_with_w256(targets, matrix, gap_open, gap_extend)`` corresponds to building one
    ``SharedProfiles::new_with_w256`` per target; ``sw_score_batch(seqs)[i][j]`` then equals
    ``profiles[j].sw_score_from_i8(seqs[i])`` and ``sw_align_batch(SeqSrc.Query(seqs))[i][j]`` equals
    ``profiles[j].sw_align_from_i8(SeqSrc::Query(seqs[i]))``.
    """

    def __init__(self, targets: Iterable[bytes], matrix: WeightMatrix, gap_open: int, gap_extend: int,
                 lanes=(32, 16, 8), devices: Optional[Sequence[int]] = None, n_devices: int = 1,
                 profiled_is_query: bool = False):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        targets = [bytes(t) for t in targets]
        # zoe validates the profile arguments before anything touches a device
        if not (-127 <= gap_open <= 0):
            raise ProfileError(ProfileError.GapOpenOutOfRange, str(gap_open))
        if not (-127 <= gap_extend <= 0):
            raise ProfileError(ProfileError.GapExtendOutOfRange, str(gap_extend))
        if gap_extend < gap_open:
            raise ProfileError(ProfileError.BadGapWeights, f"open {gap_open} extend {gap_extend}")
        if any(len(t) == 0 for t in targets) or not targets:
            raise ProfileError(ProfileError.EmptySequence)
        if devices is not None:
            ids = (C.c_int * len(devices))(*devices)
            rc = self._lib.zoe_cuda_create(C.byref(self._h), ids, len(devices))
        else:
            rc = self._lib.zoe_cuda_create(C.byref(self._h), None, n_devices)
        if rc:
            raise ZoeCudaError(rc, "zoe_cuda_create failed (no usable CUDA device?)")
        self.matrix = matrix
        self.targets = targets
        self.gap_open, self.gap_extend = gap_open, gap_extend
        self.profiled_is_query = bool(profiled_is_query)
        w = np.ascontiguousarray(matrix.weights, dtype=np.int8)
        lut = np.ascontiguousarray(matrix.mapping.index_map, dtype=np.uint8)
        self._check(self._lib.zoe_cuda_set_scoring(self._h, _p(w, C.c_int8), matrix.S, _p(lut, C.c_uint8), gap_open,
                                                    gap_extend, int(profiled_is_query)))
        self._check(self._lib.zoe_cuda_set_lanes(self._h, *lanes))
        buf, offs = _pack(targets)
        self._check(self._lib.zoe_cuda_set_profiled(self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), len(targets)))
        self.lanes = tuple(lanes)

    # lane presets: profile_set.rs:434-483
    @classmethod
    def new_with_w128(cls, targets, matrix, gap_open, gap_extend, **kw):
        return cls(targets, matrix, gap_open, gap_extend, lanes=(16, 8, 4), **kw)

    @classmethod
    def new_with_w256(cls, targets, matrix, gap_open, gap_extend, **kw):
        return cls(targets, matrix, gap_open, gap_extend, lanes=(32, 16, 8), **kw)

    @classmethod
    def new_with_w512(cls, targets, matrix, gap_open, gap_extend, **kw):
        return cls(targets, matrix, gap_open, gap_extend, lanes=(64, 32, 16), **kw)

    ALIGN_AUTO, ALIGN_FULL, ALIGN_WINDOW = 0, 1, 2

    def set_align_options(self, mode: int = 0, checkpoint_log2: int = 6, slack: int = 16):
        """Tuning of the align pipeline (never changes results): see ``zoe_cuda_set_align_options``."""
        self._check(self._lib.zoe_cuda_set_align_options(self._h, mode, checkpoint_log2, slack))

    def set_memory_budget(self, scratch_bytes: int = 0):
        """Bound the device scratch of one align / ranges / 3-pass call (0 = automatic); larger batches run in chunks."""
        self._check(self._lib.zoe_cuda_set_memory_budget(self._h, int(scratch_bytes)))

    def set_width_policy(self, first_bits: int = 8, last_bits: int = 32, unsigned: bool = False):
        """Which zoe integer types may report a result: ``(8, 32)`` = ``sw_*_from_i8`` (default), ``(16, 32)`` =
        ``..._from_i16``, ``(32, 32)`` = ``..._from_i32`` (profile_set.rs:71-179); ``first == last`` = a standalone
        ``StripedProfile<T, N, S>``; ``unsigned`` = zoe's u8/u16/u32 profiles over the biased matrix."""
        self._check(self._lib.zoe_cuda_set_width_policy(self._h, first_bits, last_bits, int(unsigned)))

    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self._lib.zoe_cuda_last_error(self._h).decode(errors="replace")
        if rc in _PROFILE_ERRORS:
            raise ProfileError(_PROFILE_ERRORS[rc], msg)
        raise ZoeCudaError(rc, msg)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.zoe_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_profiled(self) -> int:
        return len(self.targets)

    # ---- raw array API (what bench.py times) ----
    def sw_score_arrays(self, buf: np.ndarray, offs: np.ndarray):
        """Scores for a packed batch: returns (score u32, status u8, tier u8), each ``[n, n_profiled]``."""
        n = len(offs) - 1
        pairs = n * self.n_profiled
        score = np.zeros(max(pairs, 1), dtype=np.uint32)
        status = np.zeros(max(pairs, 1), dtype=np.uint8)
        tier = np.zeros(max(pairs, 1), dtype=np.uint8)
        self._check(self._lib.zoe_cuda_sw_score_batch(self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), n,
                                                      _p(score, C.c_uint32), _p(status, C.c_uint8), _p(tier, C.c_uint8)))
        shape = (n, self.n_profiled)
        return score[:pairs].reshape(shape), status[:pairs].reshape(shape), tier[:pairs].reshape(shape)

    def sw_score_into(self, buf: np.ndarray, offs: np.ndarray, score: np.ndarray, status: np.ndarray, tier: np.ndarray):
        """Like :meth:`sw_score_arrays` but writes into caller-owned (ideally pinned) output arrays."""
        self._check(self._lib.zoe_cuda_sw_score_batch(self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), len(offs) - 1,
                                                      _p(score, C.c_uint32), _p(status, C.c_uint8), _p(tier, C.c_uint8)))

    def stream_handle(self, dev_index: int = 0) -> int:
        return int(self._lib.zoe_cuda_stream(self._h, dev_index) or 0)

    def stage(self, buf: np.ndarray, offs: np.ndarray):
        self._check(self._lib.zoe_cuda_stage_streamed(self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), len(offs) - 1))
        self._staged_n = len(offs) - 1

    def run_score_staged(self):
        self._check(self._lib.zoe_cuda_run_score_staged(self._h))

    def run_align_staged(self):
        self._check(self._lib.zoe_cuda_run_align_staged(self._h))

    def fetch_scores(self):
        pairs = self._staged_n * self.n_profiled
        score = np.zeros(max(pairs, 1), dtype=np.uint32)
        status = np.zeros(max(pairs, 1), dtype=np.uint8)
        tier = np.zeros(max(pairs, 1), dtype=np.uint8)
        self._check(self._lib.zoe_cuda_fetch_scores(self._h, _p(score, C.c_uint32), _p(status, C.c_uint8),
                                                    _p(tier, C.c_uint8)))
        shape = (self._staged_n, self.n_profiled)
        return score[:pairs].reshape(shape), status[:pairs].reshape(shape), tier[:pairs].reshape(shape)

    def last_timing(self):
        t, k, n = C.c_float(0), C.c_float(0), C.c_uint32(0)
        self._lib.zoe_cuda_last_timing(self._h, C.byref(t), C.byref(k), C.byref(n))
        return {"total_ms": t.value, "dp_kernel_ms": k.value, "kernel_launches": n.value}

    def last_stats(self):
        s = _lib.Stats()
        self._lib.zoe_cuda_last_stats(self._h, C.byref(s))
        return s.as_dict()

    def dpx_peak(self, kind: int = 0):
        g, ms = C.c_double(0), C.c_float(0)
        self._check(self._lib.zoe_cuda_dpx_peak(self._h, kind, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def align_arrays(self, buf: np.ndarray, offs: np.ndarray, cigar_cap: Optional[int] = None, three_pass: bool = False):
        """Alignments for a packed batch; returns a dict of flat arrays (pair index = i*n_profiled+j).
        ``three_pass``: ``sw_align_from_i8_3pass`` (ranges + banded box alignment) instead of ``sw_align_from_i8``."""
        n = len(offs) - 1
        pairs = max(n * self.n_profiled, 1)
        out = {
            "score": np.zeros(pairs, dtype=np.uint32), "status": np.zeros(pairs, dtype=np.uint8),
            "tier": np.zeros(pairs, dtype=np.uint8), "ref_start": np.zeros(pairs, dtype=np.uint32),
            "ref_end": np.zeros(pairs, dtype=np.uint32), "query_start": np.zeros(pairs, dtype=np.uint32),
            "query_end": np.zeros(pairs, dtype=np.uint32), "cigar_off": np.zeros(pairs + 1, dtype=np.uint64),
            "hazard": np.zeros(pairs, dtype=np.uint8),
        }
        cap = int(cigar_cap) if cigar_cap else max(16 * pairs, 1024)
        for _ in range(2):
            out["cigar"] = np.zeros(cap, dtype=np.uint32)
            rc = self._align_call(buf, offs, out, three_pass)
            if rc == _lib.E_CIGAR_CAP:
                cap = int(out["cigar_off"][0]) + 16
                continue
            self._check(rc)
            break
        else:
            self._check(rc)
        return out

    def _align_call(self, buf: np.ndarray, offs: np.ndarray, out: dict, three_pass: bool) -> int:
        args = [self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), len(offs) - 1, _p(out["score"], C.c_uint32),
                _p(out["status"], C.c_uint8), _p(out["tier"], C.c_uint8), _p(out["ref_start"], C.c_uint32),
                _p(out["ref_end"], C.c_uint32), _p(out["query_start"], C.c_uint32), _p(out["query_end"], C.c_uint32),
                _p(out["cigar"], C.c_uint32), _p(out["cigar_off"], C.c_uint64), len(out["cigar"])]
        if three_pass:
            return self._lib.zoe_cuda_sw_align_3pass_batch(*args)
        return self._lib.zoe_cuda_sw_align_batch(*args, _p(out["hazard"], C.c_uint8))

    def align_into(self, buf: np.ndarray, offs: np.ndarray, out: dict, three_pass: bool = False):
        """Like :meth:`align_arrays` but into caller-owned (ideally pinned) arrays; ``out['cigar']`` is the capacity."""
        self._check(self._align_call(buf, offs, out, three_pass))

    def run_3pass_staged(self):
        self._check(self._lib.zoe_cuda_run_3pass_staged(self._h))

    def ranges_arrays(self, buf: np.ndarray, offs: np.ndarray):
        """``sw_score_ranges`` for a packed batch: dict of flat arrays (pair index = i*n_profiled+j)."""
        n = len(offs) - 1
        pairs = max(n * self.n_profiled, 1)
        out = {k: np.zeros(pairs, dtype=np.uint32) for k in ("score", "ref_start", "ref_end", "query_start", "query_end")}
        out.update({k: np.zeros(pairs, dtype=np.uint8) for k in ("status", "tier")})
        self._check(self._lib.zoe_cuda_sw_score_ranges_batch(
            self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), n, _p(out["score"], C.c_uint32),
            _p(out["status"], C.c_uint8), _p(out["tier"], C.c_uint8), _p(out["ref_start"], C.c_uint32),
            _p(out["ref_end"], C.c_uint32), _p(out["query_start"], C.c_uint32), _p(out["query_end"], C.c_uint32)))
        return out

    def ranges_into(self, buf: np.ndarray, offs: np.ndarray, out: dict):
        """Like :meth:`ranges_arrays` but into caller-owned (ideally pinned) arrays."""
        self._check(self._lib.zoe_cuda_sw_score_ranges_batch(
            self._h, _p(buf, C.c_uint8), _p(offs, C.c_uint64), len(offs) - 1, _p(out["score"], C.c_uint32),
            _p(out["status"], C.c_uint8), _p(out["tier"], C.c_uint8), _p(out["ref_start"], C.c_uint32),
            _p(out["ref_end"], C.c_uint32), _p(out["query_start"], C.c_uint32), _p(out["query_end"], C.c_uint32)))

    def run_ranges_staged(self):
        self._check(self._lib.zoe_cuda_run_ranges_staged(self._h))

    def sw_score_ranges_batch(self, src: "SeqSrc"):
        """``out[i][j] == profiles[j].sw_score_ranges_from_i8(SeqSrc::X(seqs[i]))`` as
        ``MaybeAligned[ScoreAndRanges]`` (profile_set.rs:313-322)."""
        want_pq = src.kind == "Reference"
        if want_pq != self.profiled_is_query:
            raise ValueError(
                f"this CudaProfiles was built with profiled_is_query={self.profiled_is_query}; "
                f"SeqSrc.{src.kind} needs the opposite orientation")
        seqs = [bytes(s) for s in src.seq]
        buf, offs = _pack(seqs)
        a = self.ranges_arrays(buf, offs)
        out = []
        for i in range(len(seqs)):
            row = []
            for j in range(self.n_profiled):
                k = i * self.n_profiled + j
                st = int(a["status"][k])
                if st != _lib.SOME:
                    row.append(self._maybe(st, None))
                else:
                    row.append(MaybeAligned.some(ScoreAndRanges(
                        int(a["score"][k]), (int(a["ref_start"][k]), int(a["ref_end"][k])),
                        (int(a["query_start"][k]), int(a["query_end"][k])))))
            out.append(row)
        return out

    def align_flag_bytes(self, n_seqs: int, seq_len: int) -> int:
        """Bytes of direction flags one align pass writes (8 lanes x 32 B per column per sequence pair for <= 152 rows)."""
        cols = sum(len(t) for t in self.targets)
        lanes_words = {True: 8 * 8}  # G=8, NW=8 (K=19) -- the short-read layout
        return int((n_seqs + 1) // 2 * cols * lanes_words[True] * 4) if seq_len <= 152 else 0

    # ---- zoe-shaped API ----
    def sw_score_batch(self, seqs: Sequence[bytes]):
        """``[[MaybeAligned[int]]]``: ``out[i][j] == profiles[j].sw_score_from_i8(seqs[i])``."""
        buf, offs = _pack([bytes(s) for s in seqs])
        score, status, _ = self.sw_score_arrays(buf, offs)
        return [[self._maybe(int(status[i, j]), int(score[i, j])) for j in range(self.n_profiled)]
                for i in range(len(seqs))]

    @staticmethod
    def _maybe(status: int, value):
        if status == _lib.SOME:
            return MaybeAligned.some(value)
        return MaybeAligned.Overflowed if status == _lib.OVERFLOWED else MaybeAligned.Unmapped  # type: ignore

    def sw_align_3pass_batch(self, src: SeqSrc):
        """``out[i][j] == profiles[j].sw_align_from_i8_3pass(SeqSrc::X(seqs[i]))`` (profile_set.rs:213-231)."""
        return self.sw_align_batch(src, three_pass=True)

    def sw_align_batch(self, src: SeqSrc, three_pass: bool = False):
        """``out[i][j] == profiles[j].sw_align_from_i8(SeqSrc::X(seqs[i]))`` for the SeqSrc given."""
        want_pq = src.kind == "Reference"  # streamed sequences are references <=> profiled are queries
        if want_pq != self.profiled_is_query:
            raise ValueError(
                f"this CudaProfiles was built with profiled_is_query={self.profiled_is_query}; "
                f"SeqSrc.{src.kind} needs the opposite orientation")
        seqs = [bytes(s) for s in src.seq]
        buf, offs = _pack(seqs)
        a = self.align_arrays(buf, offs, three_pass=three_pass)
        out = []
        for i in range(len(seqs)):
            row = []
            for j in range(self.n_profiled):
                k = i * self.n_profiled + j
                st = int(a["status"][k])
                if st != _lib.SOME:
                    row.append(self._maybe(st, None))
                    continue
                lo, hi = int(a["cigar_off"][k]), int(a["cigar_off"][k + 1])
                cig = "".join(f"{int(w) >> 4}{_CIGAR_OPS[int(w) & 15]}" for w in a["cigar"][lo:hi])
                streamed_len, prof_len = len(seqs[i]), len(self.targets[j])
                if self.profiled_is_query:
                    ref_len, query_len = streamed_len, prof_len
                else:
                    ref_len, query_len = prof_len, streamed_len
                row.append(MaybeAligned.some(Alignment(
                    int(a["score"][k]), (int(a["ref_start"][k]), int(a["ref_end"][k])),
                    (int(a["query_start"][k]), int(a["query_end"][k])), cig, ref_len, query_len)))
            out.append(row)
        return out
