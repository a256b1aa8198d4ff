"""Host-side preparation in the product library (seeding, COO->CSR/CSC, mean, bias initialisation) against the
golden vectors generated from the reference (tests/golden, tools/make_golden.py) -- bit-exact -- and against
the reference build itself where oracle/_ref exists.  Runs without a GPU."""
import os

import numpy as np
import pytest

from cmfrec_b200 import _lib
from refload import ArraysToFill, ptr, ref
from support import csr_csc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = [np.float32, np.float64]


def tag(dt):
    return "f32" if np.dtype(dt) == np.float32 else "f64"


@pytest.mark.parametrize("dtype", DTYPES)
def test_random_init_matches_golden(dtype):
    dt = np.dtype(dtype)
    L = _lib.load(dt)
    g = np.load(os.path.join(GOLD, "init_%s.npz" % tag(dt)))
    for name in ("small_normal", "small_uniform_req", "big_normal", "big_uniform"):
        sa, sb, normal, seed = [int(v) for v in g[name + "_args"]]
        A = np.zeros(sa, dt); B = np.zeros(max(sb, 1), dt)
        L.cmfb200_random_init(ptr(A), sa, ptr(B) if sb else None, sb, seed, bool(normal))
        assert np.array_equal(A[g[name + "_selA"]], g[name + "_A"]), name
        assert np.array_equal(B[:sb][:256], g[name + "_B"]), name
        assert np.array_equal(np.array([A.astype(np.float64).sum(), np.abs(A.astype(np.float64)).sum()]), g[name + "_sumA"])


@pytest.mark.parametrize("dtype", DTYPES)
def test_csr_mean_bias_match_golden(dtype):
    dt = np.dtype(dtype)
    L = _lib.load(dt)
    g = np.load(os.path.join(GOLD, "prep_%s.npz" % tag(dt)))
    m, n = int(g["m"]), int(g["n"])
    ia, ib, x = g["ixA"].copy(), g["ixB"].copy(), g["X"].copy()
    got = csr_csc(L, dt, ia, ib, x, m, n)
    for a, name in zip(got, ("csr_p", "csr_i", "csr_v", "csc_p", "csc_i", "csc_v")):
        assert np.array_equal(a, g[name]), name
    assert L.cmfb200_global_mean(ptr(x), x.size, 1) == g["mean_nt1"][0]
    # nthreads >= 8: the reference reduces partial sums in parallel; same value up to the last bit of a double sum
    assert abs(L.cmfb200_global_mean(ptr(x), x.size, 8) - g["mean_nt8"][0]) <= 4 * np.finfo(dt).eps * abs(g["mean_nt8"][0])
    xc = (x - g["mean_nt1"][0]).astype(dt)
    c = csr_csc(L, dt, ia, ib, xc, m, n)
    for scale in (0, 1):
        bA = np.zeros(m, dt); bB = np.zeros(n, dt)
        L.cmfb200_init_biases_twosided(m, n, *[ptr(t) for t in c], 0.05, 0.07, bool(scale), False, ptr(bA), ptr(bB), 2)
        assert np.array_equal(bA, g["biasA_scale%d" % scale])
        assert np.array_equal(bB, g["biasB_scale%d" % scale])


@pytest.mark.parametrize("dtype", DTYPES)
def test_edge_cases(dtype):
    """empty rows/columns, a single entry, duplicated (row, col) pairs kept as separate observations (Q18)"""
    dt = np.dtype(dtype)
    L = _lib.load(dt)
    ia = np.array([4, 4, 0, 4, 2], np.int32); ib = np.array([1, 1, 3, 0, 1], np.int32)
    x = np.array([1, 2, 3, 4, 5], dt)
    p, i, v, cp, ci, cv = csr_csc(L, dt, ia, ib, x, 6, 5)
    assert p.tolist() == [0, 1, 1, 2, 2, 5, 5]
    assert i.tolist() == [3, 1, 1, 1, 0] and v.tolist() == [3, 5, 1, 2, 4]
    assert cp.tolist() == [0, 1, 4, 4, 5, 5]
    assert ci.tolist() == [4, 4, 4, 2, 0] and cv.tolist() == [4, 1, 2, 5, 3]
    # nnz == 0
    p, i, v, cp, ci, cv = csr_csc(L, dt, ia[:0], ib[:0], x[:0], 3, 2)
    assert p.tolist() == [0, 0, 0, 0] and cp.tolist() == [0, 0, 0]


@pytest.mark.parametrize("dtype", DTYPES)
def test_against_reference_build(dtype):
    dt = np.dtype(dtype)
    R = ref(dt)
    if R is None:
        pytest.skip("oracle/_ref not built on this machine")
    L = _lib.load(dt)
    for sa, sb, normal, seed in [(1000, 500, True, 9), (300000, 70001, True, -5), (300001, 0, False, 2**31 - 1),
                                 (2**18 + 5, 3, False, 0), (7, 0, False, 3)]:
        A1 = np.full(sa, -7, dt); B1 = np.full(max(sb, 1), -7, dt); A2 = A1.copy(); B2 = B1.copy()
        R.random_parallel(ArraysToFill(ptr(A1), sa, ptr(B1) if sb else None, sb), seed, normal, 4)
        L.cmfb200_random_init(ptr(A2), sa, ptr(B2) if sb else None, sb, seed, normal)
        assert np.array_equal(A1, A2) and np.array_equal(B1, B2), (sa, sb, normal, seed)


@pytest.mark.parametrize("dtype", DTYPES)
def test_threaded_random_init_is_the_sequential_one(dtype):
    """cmfb200_random_init_threads (what the fits call): the generator jumped to every 2^15-draw chunk, chunks sampled
    independently for both ways of entering them and stitched in order -- must reproduce the sequential streams of
    reference random_parallel (src/helpers.c:930-1043) bit for bit, for the ziggurat (variable draws per sample) and
    the uniform sampler (odd-length quirk included), whatever the number of threads."""
    dt = np.dtype(dtype)
    L = _lib.load(dt)
    cases = [(1500007, 0, True, 1), (300000, 700001, True, -5), (300001, 0, False, 7), (2**18 + 5, 3, False, 0),
             (2**17 * 4, 2**17 * 4 + 1, True, 5), (1000003, 500001, False, 123), (1000, 500, True, 9)]
    for sa, sb, normal, seed in cases:
        A1 = np.full(sa, -7, dt); B1 = np.full(max(sb, 1), -7, dt)
        L.cmfb200_random_init(ptr(A1), sa, ptr(B1) if sb else None, sb, seed, normal)
        for nt in (2, 3, 8):
            A2 = np.full(sa, -7, dt); B2 = np.full(max(sb, 1), -7, dt)
            L.cmfb200_random_init_threads(ptr(A2), sa, ptr(B2) if sb else None, sb, seed, normal, nt)
            assert np.array_equal(A1, A2) and np.array_equal(B1, B2), (sa, sb, normal, seed, nt)


def test_threaded_random_init_in_several_rounds(monkeypatch):
    """The ziggurat fill guesses how many chunks of draws it needs; when the guess falls short it goes on from where it
    stopped (state and entry carried over).  CMFB200_RNG_TIGHT makes every round fall short."""
    dt = np.dtype(np.float32)
    L = _lib.load(dt)
    sa = 1500007
    A1 = np.zeros(sa, dt); A2 = np.zeros(sa, dt)
    L.cmfb200_random_init(ptr(A1), sa, None, 0, 3, True)
    monkeypatch.setenv("CMFB200_RNG_TIGHT", "1")
    L.cmfb200_random_init_threads(ptr(A2), sa, None, 0, 3, True, 4)
    assert np.array_equal(A1, A2)
