"""CPU-only: host-side mirror of zoe's data layer and argument validation."""
import numpy as np
import pytest

from zoe_b200 import DNA_PROFILE_MAP, BLOSUM_62, WeightMatrix, ProfileError, CudaProfiles, MaybeAligned
from zoe_b200 import synth


def test_dna_profile_map():
    m = DNA_PROFILE_MAP
    assert [m.to_index(b) for b in b"ACGTN"] == [0, 1, 2, 3, 4]
    assert [m.to_index(b) for b in b"acgtn"] == [0, 1, 2, 3, 4]
    assert m.to_index(ord("U")) == 3 and m.to_index(ord("u")) == 3
    assert m.to_index(ord("R")) == 4 and m.to_index(0) == 4  # catch-all


def test_dna_matrix():
    w = WeightMatrix.new_dna_matrix(2, -5, b"N")
    assert w.get_weight(ord("A"), ord("A")) == 2 and w.get_weight(ord("A"), ord("C")) == -5
    assert w.get_weight(ord("N"), ord("A")) == 0 and w.get_weight(ord("N"), ord("N")) == 0
    assert w.get_bias() == 5 and w.is_symmetric()
    with pytest.raises(ValueError):
        WeightMatrix.new_dna_matrix(2, -5, b"X")


def test_blosum62_spot_values():
    b = BLOSUM_62
    assert b.S == 25 and b.is_symmetric()
    assert b.get_weight(ord("W"), ord("W")) == 11 and b.get_weight(ord("A"), ord("A")) == 4
    assert b.get_weight(ord("*"), ord("*")) == 1 and b.get_weight(ord("A"), ord("*")) == -4
    assert b.get_weight(ord("X"), ord("X")) == -1 and b.get_weight(ord("a"), ord("x")) == 0


def test_profile_errors_before_any_device_work():
    w = WeightMatrix.new_dna_matrix(2, -5, b"N")
    with pytest.raises(ProfileError) as e:
        CudaProfiles([b""], w, -10, -1)
    assert e.value.kind == ProfileError.EmptySequence
    with pytest.raises(ProfileError) as e:
        CudaProfiles([b"ACGT"], w, -128, -1)
    assert e.value.kind == ProfileError.GapOpenOutOfRange
    with pytest.raises(ProfileError) as e:
        CudaProfiles([b"ACGT"], w, -10, 1)
    assert e.value.kind == ProfileError.GapExtendOutOfRange
    with pytest.raises(ProfileError) as e:
        CudaProfiles([b"ACGT"], w, -1, -10)
    assert e.value.kind == ProfileError.BadGapWeights


def test_maybe_aligned():
    assert MaybeAligned.some(3).unwrap() == 3
    with pytest.raises(ValueError):
        MaybeAligned.Overflowed.unwrap()
    assert MaybeAligned.Unmapped != MaybeAligned.Overflowed


def test_synth_is_seeded_and_shaped():
    t1, r1 = synth.config2(n_reads=200)
    t2, r2 = synth.config2(n_reads=200)
    assert [len(t) for t in t1] == list(synth.FLU_SEGMENT_LENGTHS)
    assert r1.shape == (200, 150) and np.array_equal(r1, r2)
    assert set(np.unique(r1)) <= set(b"ACGT")
    _, q = synth.config5(n_queries=50)
    assert q.shape == (50, 300)
    _, reads = synth.config4(n_reads=5)
    assert all(800 < len(r) < 5600 for r in reads)
