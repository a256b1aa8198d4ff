"""Fold-in of new rows and the precomputed matrices, through the reference-named entry points against the reference
build: factors_collective_explicit_multiple (src/collective.c:10865), factors_collective_implicit_multiple (:11176),
precompute_collective_explicit (:10209), precompute_collective_implicit (:10487).  The row solves are exact (Cholesky)
in both libraries: fp64 agrees to 1e-9 of the largest factor, fp32 to 2e-3 with the error against exact arithmetic no
worse than 3x the reference's own."""
import numpy as np
import pytest

from support import ptr, ref, synth_coo

pytestmark = pytest.mark.gpu


def _ref(dt):
    R = ref(dt)
    if R is None:
        pytest.fail("oracle/_ref is not built")
    return R


def call_explicit(lib, dt, m, n, k, ixA, ixB, X, B, biasB, glob_mean, *, user_bias, lam=0.7, lam_unique=None, scale_lam=False,
                  w_main=1.0, csr=None, scale_bias_const=False, scaling_biasA=0.0):
    A = np.full((m, k), 7.0, dt)
    biasA = np.full(m, 7.0, dt) if user_bias else None
    ia = None if ixA is None else np.ascontiguousarray(ixA, np.int32).copy()
    ib = None if ixB is None else np.ascontiguousarray(ixB, np.int32).copy()
    x = None if X is None else np.ascontiguousarray(X, dt).copy()
    lu = None if lam_unique is None else np.ascontiguousarray(lam_unique, dt)
    cp = ci = cv = None
    if csr is not None:
        cp, ci, cv = csr
    rc = lib.factors_collective_explicit_multiple(
        ptr(A), ptr(biasA), m, None, 0, 0, False, False, False,
        None, None, None, 0, None, None, None,
        None, 0, 0, None, None, glob_mean, ptr(biasB), None,
        ptr(x), ptr(ia), ptr(ib), 0 if x is None else x.size, ptr(cp), ptr(ci), ptr(cv), None, n, None, ptr(B),
        None, False, k, 0, 0, 0,
        lam, ptr(lu), 0.0, None, scale_lam, False, scale_bias_const, scaling_biasA,
        w_main, 1.0, 1.0, n, False,
        None, None, None, None, None, None, None, None, None, 4)
    return rc, A, biasA


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k", [8, 40, 64, 128])
@pytest.mark.parametrize("user_bias,item_bias", [(True, True), (False, False), (True, False)])
def test_explicit_foldin_matches_reference(gpu_libs, dtype, k, user_bias, item_bias):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n = 700, 500
    rng = np.random.default_rng(k + 3 * user_bias)
    ixA, ixB, X = synth_coo(m, n, 30000, dt, seed=k)
    keep = ixA % 50 != 7                      # some new rows have no entries at all
    ixA, ixB, X = ixA[keep], ixB[keep], X[keep]
    B = (rng.normal(size=(n, k)) / np.sqrt(k)).astype(dt)
    biasB = rng.normal(size=n).astype(dt) * 0.3 if item_bias else None
    for opts in (dict(), dict(scale_lam=True, lam=0.05), dict(lam_unique=[0.3, 0.1, 0.9, 0.2, 0.1, 0.1], w_main=2.0)):
        o = call_explicit(L, dt, m, n, k, ixA, ixB, X, B, biasB, 3.4, user_bias=user_bias, **opts)
        r = call_explicit(R, dt, m, n, k, ixA, ixB, X, B, biasB, 3.4, user_bias=user_bias, **opts)
        assert o[0] == 0 and r[0] == 0
        scale = max(np.abs(r[1]).max(), 1e-30)
        tol = 1e-9 if dt == np.float64 else 2e-3
        assert np.isfinite(o[1]).all()
        assert np.abs(o[1] - r[1]).max() <= tol * scale, np.abs(o[1] - r[1]).max() / scale
        if user_bias:
            assert np.abs(o[2] - r[2]).max() <= tol * max(np.abs(r[2]).max(), scale)
        empty = np.setdiff1d(np.arange(m), ixA)
        assert empty.size and not o[1][empty].any() and (not user_bias or not o[2][empty].any())
        if dt == np.float32:   # error against exact arithmetic no worse than 3x the reference's own
            e = call_explicit(gpu_libs[np.dtype(np.float64)], np.dtype(np.float64), m, n, k, ixA, ixB, X, B.astype(np.float64),
                              None if biasB is None else biasB.astype(np.float64), 3.4, user_bias=user_bias, **opts)
            err_o = np.abs(o[1] - e[1]).max(axis=1); err_r = np.abs(r[1] - e[1]).max(axis=1)
            assert np.quantile(err_o, 0.99) <= 3 * np.quantile(err_r, 0.99) + 1e-6 * scale


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_explicit_foldin_from_csr_and_bad_indices(gpu_libs, dtype):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n, k = 60, 90, 12
    rng = np.random.default_rng(4)
    ixA, ixB, X = synth_coo(m, n, 1500, dt, seed=2)
    order = np.argsort(ixA, kind="stable")
    ixA, ixB, X = ixA[order], ixB[order], X[order]
    B = rng.normal(size=(n, k)).astype(dt)
    biasB = rng.normal(size=n).astype(dt)
    cp = np.zeros(m + 1, np.uint64); np.add.at(cp, ixA + 1, 1); cp = np.cumsum(cp).astype(np.uint64)
    ci = ixB.astype(np.int32).copy(); cv = X.astype(dt).copy()
    o_coo = call_explicit(L, dt, m, n, k, ixA, ixB, X, B, biasB, 3.0, user_bias=True)
    o_csr = call_explicit(L, dt, m, n, k, None, None, None, B, biasB, 3.0, user_bias=True, csr=(cp, ci, cv))
    assert o_coo[0] == 0 and o_csr[0] == 0
    assert np.array_equal(o_coo[1], o_csr[1]) and np.array_equal(o_coo[2], o_csr[2])
    # an inadmissible column id turns that row into NaN, as in the reference (check_sparse_indices)
    ci_bad = ci.copy(); ci_bad[int(cp[5])] = n + 4
    o = call_explicit(L, dt, m, n, k, None, None, None, B, biasB, 3.0, user_bias=True, csr=(cp, ci_bad, cv))
    r = call_explicit(R, dt, m, n, k, None, None, None, B, biasB, 3.0, user_bias=True, csr=(cp, ci_bad, cv))
    assert o[0] == 0 and r[0] == 0
    assert np.isnan(o[1][5]).all() and np.isnan(o[2][5]) and np.array_equal(np.isnan(o[1]), np.isnan(r[1]))
    # refused combinations answer 2 and say so (no CPU fallback)
    U = np.ones((m, 3), dt)
    A = np.zeros((m, k), dt)
    rc = L.factors_collective_explicit_multiple(
        ptr(A), None, m, ptr(U), m, 3, False, False, False, None, None, None, 0, None, None, None, None, 0, 0, None, None, 0.0, None,
        None, ptr(cv), ptr(ixA.astype(np.int32)), ptr(ci), cv.size, None, None, None, None, n, None, ptr(B), None, False, k, 0, 0, 0,
        1.0, None, 0.0, None, False, False, False, 0.0, 1.0, 1.0, 1.0, n, False, None, None, None, None, None, None, None, None, None, 1)
    assert rc == 2


def call_implicit(lib, dt, m, n, k, ixA, ixB, X, B, *, lam=2.0, alpha=1.0, w_main=1.0, mult=1.0, log=False):
    A = np.full((m, k), 7.0, dt)
    ia = np.ascontiguousarray(ixA, np.int32).copy(); ib = np.ascontiguousarray(ixB, np.int32).copy()
    x = np.ascontiguousarray(X, dt).copy()
    kk = k
    BtB = (B.astype(np.float64).T @ B.astype(np.float64) + (lam / (w_main * mult)) * np.eye(kk)).astype(dt)   # what precompute leaves
    rc = lib.factors_collective_implicit_multiple(
        ptr(A), m, None, 0, 0, False, False, None, None, None, 0, None, None, None,
        ptr(x), ptr(ia), ptr(ib), x.size, None, None, None, ptr(B), n, None, None,
        k, 0, 0, 0, lam, 0.0, alpha, w_main, 1.0, mult, log, None, ptr(BtB), None, None, 4)
    return rc, A


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k", [8, 64, 100])
def test_implicit_foldin_matches_reference(gpu_libs, dtype, k):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n = 600, 800
    rng = np.random.default_rng(k)
    ixA, ixB, X = synth_coo(m, n, 25000, dt, seed=k + 1, kind="counts")
    B = (rng.normal(size=(n, k)) / np.sqrt(k)).astype(dt)
    # apply_log_transf is left out: the reference takes the logarithm of an uninitialised buffer there (src/collective.c:10806-10813)
    for opts in (dict(), dict(alpha=15.0), dict(w_main=2.0, mult=0.5, lam=1.0)):
        o = call_implicit(L, dt, m, n, k, ixA, ixB, X, B, **opts)
        r = call_implicit(R, dt, m, n, k, ixA, ixB, X, B, **opts)
        assert o[0] == 0 and r[0] == 0
        scale = max(np.abs(r[1]).max(), 1e-30)
        tol = 1e-9 if dt == np.float64 else 2e-3
        assert np.isfinite(o[1]).all() and np.abs(o[1] - r[1]).max() <= tol * scale, np.abs(o[1] - r[1]).max() / scale


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_precompute_entry_points_match_reference(gpu_libs, dtype):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    n, k, p = 400, 10, 6
    rng = np.random.default_rng(1)
    B = rng.normal(size=(n, k)).astype(dt); C = rng.normal(size=(p, k)).astype(dt); biasB = rng.normal(size=n).astype(dt)
    tol = 1e-10 if dt == np.float64 else 2e-4
    for user_bias in (True, False):
        for with_C in (False, True):
            outs = []
            for lib in (L, R):
                ub = 1 if user_bias else 0
                Bpb = np.zeros((n, k + 1), dt); BtB = np.zeros((k + ub, k + ub), dt); T = np.zeros((n, k + ub), dt)
                BeTBeChol = np.zeros((k + ub, k + ub), dt); TransCtCinvCt = np.zeros((p, k), dt); CtCw = np.zeros((k, k), dt)
                rc = lib.precompute_collective_explicit(
                    ptr(B), n, n, False, ptr(C) if with_C else None, p if with_C else 0, None, False, ptr(biasB), 2.5, False, None, False,
                    k, 0, 0, 0, user_bias, False, 0.8, None, False, False, False, 0.0, 1.0, 0.7, 1.0,
                    ptr(Bpb) if user_bias else None, ptr(BtB), ptr(T), None, ptr(BeTBeChol) if with_C else None, None,
                    ptr(TransCtCinvCt) if with_C else None, ptr(CtCw) if with_C else None, None)
                assert rc == 0
                outs.append((Bpb, np.triu(BtB), T, np.triu(BeTBeChol), TransCtCinvCt, np.triu(CtCw)))
            for a, b in zip(*outs):
                assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())
    outs = []
    for lib in (L, R):
        BtB = np.zeros((k, k), dt); BeTBe = np.zeros((k, k), dt); Chol = np.zeros((k, k), dt)
        rc = lib.precompute_collective_implicit(ptr(B), n, ptr(C), p, None, False, k, 0, 0, 0, 3.0, 1.0, 0.6, 1.0, False, True,
                                                ptr(BtB), ptr(BeTBe), ptr(Chol), None)
        assert rc == 0
        outs.append((np.triu(BtB), np.triu(BeTBe), np.triu(Chol)))
    for a, b in zip(*outs):
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())
