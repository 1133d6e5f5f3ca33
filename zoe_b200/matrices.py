"""Alphabet maps and substitution matrices, mirroring zoe's data layer for the SW path.

Mirrors (names, argument meaning, error behaviour):
  * ``ByteIndexMap``  -- src/data/constants/mappings/byte_index.rs:231-358
  * ``DNA_PROFILE_MAP`` -- src/data/constants/mappings/dna.rs:177-178
  * ``AA_ALL_AMBIG_PROFILE_MAP_WITH_STOP`` -- src/data/constants/mappings/aa.rs:17-18
  * ``WeightMatrix`` -- src/data/matrices/mod.rs:230-235, 358-420
  * ``BLOSUM_62`` -- src/data/matrices/aa.rs:336-366 (zoe's own X / * values)

Only data and cheap host logic live here; nothing in this module computes alignments.
"""
from __future__ import annotations

import numpy as np


class ByteIndexMap:
    """256-entry byte -> index lookup (byte_index.rs:231-234, ``to_index`` :331-333)."""

    def __init__(self, byte_keys: bytes, catch_all: int | bytes, ignore_case: bool = True):
        if isinstance(catch_all, (bytes, bytearray)):
            catch_all = catch_all[0]
        keys = bytes(byte_keys)
        if ignore_case:
            keys = keys.upper()
            catch_all = bytes([catch_all]).upper()[0]
        if len(set(keys)) != len(keys):
            raise ValueError("byte_keys must not contain duplicates")
        if catch_all not in keys:
            raise ValueError("The catch_all must be present in the byte_keys.")
        self.byte_keys = keys
        table = np.full(256, keys.index(catch_all), dtype=np.uint8)
        for i, b in enumerate(keys):
            table[b] = i
            if ignore_case:
                table[bytes([b]).lower()[0]] = i
        self.index_map = table

    def __len__(self) -> int:
        return len(self.byte_keys)

    def add_synonym_ignore_case(self, new_key: bytes | int, previous_key: bytes | int) -> "ByteIndexMap":
        nk = new_key[0] if isinstance(new_key, (bytes, bytearray)) else new_key
        pk = previous_key[0] if isinstance(previous_key, (bytes, bytearray)) else previous_key
        out = ByteIndexMap.__new__(ByteIndexMap)
        out.byte_keys = self.byte_keys
        out.index_map = self.index_map.copy()
        val = out.index_map[pk]
        out.index_map[bytes([nk]).lower()[0]] = val
        out.index_map[bytes([nk]).upper()[0]] = val
        return out

    def to_index(self, b: int) -> int:
        return int(self.index_map[b])

    def to_byte(self, index: int) -> int:
        return self.byte_keys[index]

    def in_byte_keys(self, b: int) -> bool:
        return b == self.to_byte(self.to_index(b))


#: {0: A, 1: C, 2: G, 3: T, 4: N}; N is the catch-all, U is treated as T (dna.rs:177-178).
DNA_PROFILE_MAP = ByteIndexMap(b"ACGTN", b"N").add_synonym_ignore_case(b"U", b"T")

#: 20 residues, stop, BJZ, X as catch-all (aa.rs:17-18).
AA_ALL_AMBIG_PROFILE_MAP_WITH_STOP = ByteIndexMap(b"ACDEFGHIKLMNPQRSTVWY*BJZX", b"X")


class WeightMatrix:
    """Signed substitution matrix ``weights[ref_idx][query_idx]`` (matrices/mod.rs:230-235)."""

    def __init__(self, mapping: ByteIndexMap, weights: np.ndarray):
        w = np.asarray(weights)
        S = len(mapping)
        if w.shape != (S, S):
            raise ValueError(f"weights must be {S}x{S}")
        if w.min() < -128 or w.max() > 127:
            raise ValueError("weights must fit in i8")
        self.mapping = mapping
        self.weights = np.ascontiguousarray(w, dtype=np.int8)
        self.bias = 0

    @property
    def S(self) -> int:
        return len(self.mapping)

    @classmethod
    def new(cls, mapping: ByteIndexMap, matching: int, mismatch: int, ignoring: bytes | int | None = None) -> "WeightMatrix":
        """matrices/mod.rs:358-399: any pair involving the ignored residue scores 0."""
        S = len(mapping)
        skip = None
        if ignoring is not None:
            ig = ignoring[0] if isinstance(ignoring, (bytes, bytearray)) else ignoring
            if not mapping.in_byte_keys(ig):
                raise ValueError("An invalid byte was specified for the ignoring field.")
            skip = mapping.to_index(ig)
        w = np.full((S, S), mismatch, dtype=np.int64)
        np.fill_diagonal(w, matching)
        if skip is not None:
            w[skip, :] = 0
            w[:, skip] = 0
        return cls(mapping, w)

    @classmethod
    def new_dna_matrix(cls, matching: int, mismatch: int, ignoring: bytes | int | None = None) -> "WeightMatrix":
        return cls.new(DNA_PROFILE_MAP, matching, mismatch, ignoring)

    @classmethod
    def new_custom(cls, mapping: ByteIndexMap, weights) -> "WeightMatrix":
        return cls(mapping, np.asarray(weights))

    def get_weight(self, ref_residue: int, query_residue: int) -> int:
        """matrices/mod.rs:242-244."""
        return int(self.weights[self.mapping.to_index(ref_residue), self.mapping.to_index(query_residue)])

    def get_bias(self) -> int:
        """|min(0, min weight)| -- what ``to_biased_matrix`` (matrices/mod.rs:471-491) would use."""
        return int(-min(0, int(self.weights.min())))

    def is_symmetric(self) -> bool:
        return bool(np.array_equal(self.weights, self.weights.T))


def _parse_square(text: str) -> np.ndarray:
    rows = [[int(x) for x in line.split()] for line in text.strip().splitlines()]
    return np.array(rows, dtype=np.int64)


# Row/column order ACDEFGHIKLMNPQRSTVWY*BJZX (aa.rs:336-366). Values are zoe's, which differ
# from NCBI's table in the X and * rows.
_BLOSUM62_TEXT = """
 4  0 -2 -1 -2  0 -2 -1 -1 -1 -1 -2 -1 -1 -1  1  0  0 -3 -2 -4 -2 -1 -1  0
 0  9 -3 -4 -2 -3 -3 -1 -3 -1 -1 -3 -3 -3 -3 -1 -1 -1 -2 -2 -4 -3 -1 -3 -2
-2 -3  6  2 -3 -1 -1 -3 -1 -4 -3  1 -1  0 -2  0 -1 -3 -4 -3 -4  4 -3  1 -1
-1 -4  2  5 -3 -2  0 -3  1 -3 -2  0 -1  2  0  0 -1 -2 -3 -2 -4  1 -3  4 -1
-2 -2 -3 -3  6 -3 -1  0 -3  0  0 -3 -4 -3 -3 -2 -2 -1  1  3 -4 -3  0 -3 -1
 0 -3 -1 -2 -3  6 -2 -4 -2 -4 -3  0 -2 -2 -2  0 -2 -3 -2 -3 -4 -1 -4 -2 -1
-2 -3 -1  0 -1 -2  8 -3 -1 -3 -2  1 -2  0  0 -1 -2 -3 -2  2 -4  0 -3  0 -1
-1 -1 -3 -3  0 -4 -3  4 -3  2  1 -3 -3 -3 -3 -2 -1  3 -3 -1 -4 -3  3 -3 -1
-1 -3 -1  1 -3 -2 -1 -3  5 -2 -1  0 -1  1  2  0 -1 -2 -3 -2 -4  0 -3  1 -1
-1 -1 -4 -3  0 -4 -3  2 -2  4  2 -3 -3 -2 -2 -2 -1  1 -2 -1 -4 -4  3 -3 -1
-1 -1 -3 -2  0 -3 -2  1 -1  2  5 -2 -2  0 -1 -1 -1  1 -1 -1 -4 -3  2 -1 -1
-2 -3  1  0 -3  0  1 -3  0 -3 -2  6 -2  0  0  1  0 -3 -4 -2 -4  3 -3  0 -1
-1 -3 -1 -1 -4 -2 -2 -3 -1 -3 -2 -2  7 -1 -2 -1 -1 -2 -4 -3 -4 -2 -3 -1 -2
-1 -3  0  2 -3 -2  0 -3  1 -2  0  0 -1  5  1  0 -1 -2 -2 -1 -4  0 -2  3 -1
-1 -3 -2  0 -3 -2  0 -3  2 -2 -1  0 -2  1  5 -1 -1 -3 -3 -2 -4 -1 -2  0 -1
 1 -1  0  0 -2  0 -1 -2  0 -2 -1  1 -1  0 -1  4  1 -2 -3 -2 -4  0 -2  0  0
 0 -1 -1 -1 -2 -2 -2 -1 -1 -1 -1  0 -1 -1 -1  1  5  0 -2 -2 -4 -1 -1 -1  0
 0 -1 -3 -2 -1 -3 -3  3 -2  1  1 -3 -2 -2 -3 -2  0  4 -3 -1 -4 -3  1 -2 -1
-3 -2 -4 -3  1 -2 -2 -3 -3 -2 -1 -4 -4 -2 -3 -3 -2 -3 11  2 -4 -4 -2 -3 -2
-2 -2 -3 -2  3 -3  2 -1 -2 -1 -1 -2 -3 -1 -2 -2 -2 -1  2  7 -4 -3 -1 -2 -1
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1 -4 -4 -4 -4
-2 -3  4  1 -3 -1  0 -3  0 -4 -3  3 -2  0 -1  0 -1 -3 -4 -3 -4  4 -3  0 -1
-1 -1 -3 -3  0 -4 -3  3 -3  3  2 -3 -3 -2 -2 -2 -1  1 -2 -1 -4 -3  3 -3 -1
-1 -3  1  4 -3 -2  0 -3  1 -3 -1  0 -1  3  0  0 -1 -2 -3 -2 -4  0 -3  4 -1
 0 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2 -1 -1  0  0 -1 -2 -1 -4 -1 -1 -1 -1
"""

#: The BLOSUM62 matrix as zoe ships it (aa.rs:336-366).
BLOSUM_62 = WeightMatrix.new_custom(AA_ALL_AMBIG_PROFILE_MAP_WITH_STOP, _parse_square(_BLOSUM62_TEXT))
