"""The oracle's restatement of numpy's legacy RandomState stream, checked against numpy itself
(the third-party dependency the reference's arithmetic lives in)."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("seed", [0, 1, 42, 20261019, 2**32 - 1])
def test_raw_and_double_stream(seed):
    r = O.Rng(seed)
    rs = np.random.RandomState(seed)
    ref = rs.random_sample(2000)
    got = np.array([r.double() for _ in range(2000)])
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("hi", [1, 2, 3, 255, 256, 1000003, 2**31, 2**32 - 1, 2**32, 2**32 + 1, 3_100_000_000,
                                5_000_000_000, 2**40 + 12345])
def test_bounded_matches_randint(hi):
    r = O.Rng(7)
    rs = np.random.RandomState(7)
    ref = [int(rs.randint(0, hi)) for _ in range(500)]
    got = [r.bounded(hi - 1) for _ in range(500)]
    assert ref == got


def test_interleaved_global_stream():
    """The exact call mix of the placement loop on the GLOBAL numpy stream (read_operations.py:101-103,65)."""
    np.random.seed(1234)
    r = O.Rng(1234)
    names = ["Barcode_0", "Barcode_1", "Barcode_2"]
    p = np.array([0.5, 0.3, 0.2])
    cdf = O.barcode_cdf(p)
    pool = np.arange(100, 5100)
    G = 3_100_000_000
    for _ in range(300):
        b = np.random.choice(names, p=p)
        u = r.double()
        assert names[int(np.searchsorted(cdf, u, side="right"))] == b
        L = np.random.choice(pool)
        assert pool[r.bounded(pool.size - 1)] == L
        s = np.random.randint(0, G - L)
        assert r.bounded(G - L - 1) == s
        assert np.random.rand() == r.double()


@pytest.mark.parametrize("sigma,minlen", [(1, 1), (2, 500), (1, 20000)])
def test_lognormal_pool_matches_numpy(sigma, minlen):
    np.random.seed(99)
    ref = np.random.lognormal(mean=np.log(15000) - 0.5, sigma=sigma, size=100001)  # odd size: cached gauss carries
    ref = np.round(ref).astype(int)
    ref = ref[ref > minlen]
    tail = np.random.rand()
    r = O.Rng(99)
    got = O.lognormal_pool(r, 15000, sigma, minlen, n=100001)
    assert np.array_equal(ref, got)
    assert tail == r.double()


def test_vector_choice_matches_numpy():
    np.random.seed(5)
    src = np.random.RandomState(1).randint(1, 10**6, size=77777)
    ref = np.random.choice(src, size=50000)
    r = O.Rng(5)
    got = O.choice(r, src, 50000)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("B,w", [(4, {"0": 10}), (4, None), (3, {"1": 2}), (24, {"0": 10, "1": 5}), (2, {"Barcode": 3})])
def test_barcode_probabilities(B, w):
    p = O.barcode_probabilities(B, w)
    assert abs(p.sum() - 1) < 1e-12
    if w == {"0": 10} and B == 4:  # SURVEY.md Appendix A.1 probe
        assert np.allclose(p, [10 / 13, 1 / 13, 1 / 13, 1 / 13])


def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the engine's native stream)."""
    from oracle import native_np as NN
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = NN.philox4x32_10(*[np.array([c], np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(o[0]) for o in out) == exp


def test_mulhi64_restatement():
    from oracle import native_np as NN
    r = np.random.RandomState(3)
    a = r.randint(0, 2**63, size=1000, dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    b = r.randint(1, 2**62, size=1000, dtype=np.int64).astype(np.uint64)
    got = NN.mulhi64(a, b)
    assert [int(x) for x in got] == [(int(x) * int(y)) >> 64 for x, y in zip(a, b)]
