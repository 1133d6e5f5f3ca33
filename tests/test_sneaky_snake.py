"""SneakySnake pre-alignment filter (src/alignment/sneaky_snake.rs:78-131): the oracle restatement against zoe's
doc-test and an independent edit-distance bound (CPU), and the CUDA kernel against the oracle (GPU)."""
import numpy as np
import pytest

from oracle import oracle as O


def _edit_distance(a: bytes, b: bytes) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _cases(rng, n):
    out = []
    for _ in range(n):
        L = int(rng.integers(1, 80))
        ref = rng.choice(list(b"ACGT"), L).astype(np.uint8)
        q = ref.copy()
        for _ in range(int(rng.integers(0, 8))):
            k = int(rng.integers(0, 3))
            pos = int(rng.integers(0, max(len(q), 1)))
            if k == 0 and len(q):
                q[pos] = rng.choice(list(b"ACGT"))
            elif k == 1 and len(q) > 1:
                q = np.delete(q, pos)
            else:
                q = np.insert(q, pos, rng.choice(list(b"ACGT")))
        if rng.random() < 0.15:
            q = rng.choice(list(b"ACGT"), int(rng.integers(1, 80))).astype(np.uint8)
        out.append((bytes(ref), bytes(q)))
    return out


def test_oracle_doc_example_and_argument_rules():
    # doc-test sneaky_snake.rs:54-62
    assert O.sneaky_snake(b"GGTGCAGAGCTC", b"GGTGAGAGTTGT", 0.25) is True
    # threshold outside 0..=1 -> None; length difference above the edit threshold -> None; threshold 1.0 -> Some(true)
    assert O.sneaky_snake(b"ACGT", b"ACGT", 1.5) is None and O.sneaky_snake(b"ACGT", b"ACGT", -0.1) is None
    assert O.sneaky_snake(b"ACGT", b"ACGT", float("nan")) is None
    assert O.sneaky_snake(b"AAAA", b"AAAAAAAA", 0.25) is None
    assert O.sneaky_snake(b"ACGT", b"TTTT", 1.0) is True
    assert O.sneaky_snake(b"ACGT", b"ACGT", 0.0) is True and O.sneaky_snake(b"ACGT", b"ACGA", 0.0) is False


def test_oracle_never_rejects_a_pair_within_the_threshold():
    # the documented guarantee: the approximated distance is <= the true edit distance, so any pair whose global edit
    # distance is within the threshold must pass (sneaky_snake.rs:30-33)
    rng = np.random.default_rng(3)
    seen = {True: 0, False: 0, None: 0}
    for ref, q in _cases(rng, 600):
        for thr in (0.05, 0.1, 0.25, 0.5):
            r = O.sneaky_snake(ref, q, thr)
            seen[r] += 1
            if r is False:
                assert _edit_distance(ref, q) > int(np.floor(np.float32(len(q)) * np.float32(thr))), (ref, q, thr)
    assert all(v > 0 for v in seen.values()), seen


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from zoe_b200 import SneakySnake
    rng = np.random.default_rng(4)
    cases = _cases(rng, 3000) + [(b"GGTGCAGAGCTC", b"GGTGAGAGTTGT"), (b"A", b"A"), (b"A", b"C"), (b"", b""), (b"ACGT", b"")]
    snake = SneakySnake()
    for thr in (0.0, 0.05, 0.1, 0.25, 0.5, 1.0, 1.5, float("nan")):
        got = snake.sneaky_snake_batch([c[0] for c in cases], [c[1] for c in cases], thr)
        want = [O.sneaky_snake(r, q, thr) for r, q in cases]
        assert got == want, [(i, cases[i], got[i], want[i]) for i in range(len(cases)) if got[i] != want[i]][:5]
    assert snake.sneaky_snake_batch([b"GGTGCAGAGCTC"], [b"GGTGAGAGTTGT"], 0.25) == [True]
    assert snake.sneaky_snake_batch([], [], 0.25) == []
    snake.close()


@pytest.mark.gpu
def test_gpu_read_sized_batch():
    # 150-nt reads against their true 150-nt windows (what a caller filters before sw_align)
    from zoe_b200 import SneakySnake
    rng = np.random.default_rng(8)
    refs, qs = [], []
    for _ in range(2000):
        ref = rng.choice(list(b"ACGT"), 150).astype(np.uint8)
        q = ref.copy()
        for pos in rng.integers(0, 150, int(rng.integers(0, 30))):
            q[pos] = rng.choice(list(b"ACGT"))
        refs.append(bytes(ref))
        qs.append(bytes(q))
    snake = SneakySnake()
    got = snake.sneaky_snake_batch(refs, qs, 0.1)
    assert got == [O.sneaky_snake(r, q, 0.1) for r, q in zip(refs, qs)]
    assert True in got and False in got
    snake.close()


@pytest.mark.gpu
def test_gpu_warp_and_thread_kernels_agree_on_long_and_ragged_pairs():
    # pairs beyond the warp kernel's staging area take zoe's loops one thread per pair; both must give the oracle's answer
    import os
    from zoe_b200 import SneakySnake
    rng = np.random.default_rng(12)
    refs, qs = [], []
    for L in list(rng.integers(1, 1400, 150)) + [1023, 1024, 1025]:
        ref = rng.choice(list(b"ACGT"), int(L)).astype(np.uint8)
        q = ref.copy()
        for pos in rng.integers(0, int(L), int(rng.integers(0, max(2, int(L) // 8)))):
            q[pos] = rng.choice(list(b"ACGT"))
        if rng.random() < 0.3 and L > 4:
            q = np.delete(q, int(rng.integers(0, len(q))))
        refs.append(bytes(ref))
        qs.append(bytes(q))
    short = [i for i in range(len(refs)) if max(len(refs[i]), len(qs[i])) <= 1024]
    snake = SneakySnake()
    for thr in (0.02, 0.1, 0.3):
        want = [O.sneaky_snake(r, q, thr) for r, q in zip(refs, qs)]
        assert snake.sneaky_snake_batch(refs, qs, thr) == want                       # thread kernel (a pair > 1024)
        sr, sq = [refs[i] for i in short], [qs[i] for i in short]
        assert snake.sneaky_snake_batch(sr, sq, thr) == [want[i] for i in short]    # warp kernel
        os.environ["ZOE_CUDA_SNAKE_THREAD"] = "1"
        try:
            assert snake.sneaky_snake_batch(sr, sq, thr) == [want[i] for i in short]
        finally:
            os.environ.pop("ZOE_CUDA_SNAKE_THREAD", None)
    snake.close()
