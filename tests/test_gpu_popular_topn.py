"""fit_most_popular and topN through the reference-named C entry points, against the reference build.
topN: the ranking is integer work -- indices must agree exactly wherever neighbouring scores differ by more
than 1e-6 relative (the reference's own test demands full agreement with an argsort, test_math/test_topN.py:78-81).
fit_most_popular: float32 results are reproduced bit for bit (same operations in the same order); float64 to
4 ulp (the reference's build may contract a*b+c into an FMA, this library's kernels do not)."""
import numpy as np
import pytest

from support import ptr, ref, synth_coo

pytestmark = pytest.mark.gpu


def _ref(dt):
    R = ref(dt)
    if R is None:
        pytest.fail("oracle/_ref is not built")
    return R


def call_topn(lib, dt, a, B, biasB, glob_mean, biasA, k, n_top, include=None, exclude=None):
    n = B.shape[0]
    out_ix = np.zeros(n_top, np.int32)
    out_sc = np.zeros(n_top, dt)
    inc = None if include is None else np.ascontiguousarray(include, np.int32).copy()
    exc = None if exclude is None else np.ascontiguousarray(exclude, np.int32).copy()
    rc = lib.topN(ptr(a), 0, ptr(B), 0, ptr(biasB), glob_mean, biasA, k, 0, ptr(inc), 0 if inc is None else inc.size,
                  ptr(exc), 0 if exc is None else exc.size, ptr(out_ix), ptr(out_sc), n_top, n, 4)
    return rc, out_ix, out_sc


def same_ranking(ix_a, sc_a, ix_b, sc_b, rtol):
    """identical indices except inside groups of (near-)tied scores"""
    if np.array_equal(ix_a, ix_b):
        return True
    scale = max(np.abs(sc_b).max(), 1e-30)
    for i in np.nonzero(ix_a != ix_b)[0]:
        if abs(sc_a[i] - sc_b[i]) > rtol * scale:
            return False
    return True


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,k", [(37, 4), (2000, 16), (50000, 64)])
def test_topn_matches_reference(gpu_libs, dtype, n, k):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    rng = np.random.default_rng(n + k)
    B = rng.normal(size=(n, k)).astype(dt)
    a = rng.normal(size=k).astype(dt)
    biasB = rng.normal(size=n).astype(dt)
    rtol = 1e-12 if dt == np.float64 else 1e-5
    cases = [dict(n_top=min(10, n)), dict(n_top=min(100, n // 2)), dict(n_top=n),
             dict(n_top=5, include=rng.choice(n, size=max(n // 3, 6), replace=False)),
             dict(n_top=7, exclude=rng.choice(n, size=n // 4, replace=False)),
             dict(n_top=3, exclude=np.sort(rng.choice(n, size=n // 2, replace=False)))]
    for use_bias in (True, False):
        for c in cases:
            rc1, ix1, sc1 = call_topn(L, dt, a, B, biasB if use_bias else None, 0.3, -0.2, k, **c)
            rc2, ix2, sc2 = call_topn(R, dt, a, B, biasB if use_bias else None, 0.3, -0.2, k, **c)
            assert rc1 == 0 and rc2 == 0
            assert same_ranking(ix1, sc1, ix2, sc2, rtol), (c.get("n_top"), ix1[:8], ix2[:8])
            assert np.allclose(sc1, sc2, rtol=0, atol=(1e-12 if dt == np.float64 else 2e-5) * max(1.0, np.abs(sc2).max()))
            # size-independent properties: scores decreasing, indices distinct and admissible
            assert np.all(np.diff(sc1) <= 0)
            assert len(set(ix1.tolist())) == ix1.size
            if "include" in c:
                assert set(ix1.tolist()) <= set(np.asarray(c["include"]).tolist())
            if "exclude" in c:
                assert not (set(ix1.tolist()) & set(np.asarray(c["exclude"]).tolist()))


def test_topn_rejects_bad_arguments(gpu_libs):
    dt = np.dtype(np.float64)
    L = gpu_libs[dt]
    B = np.ones((10, 3)); a = np.ones(3)
    assert call_topn(L, dt, a, B, None, 0.0, 0.0, 3, 0)[0] == 2                      # n_top == 0
    assert call_topn(L, dt, a, B, None, 0.0, 0.0, 3, 5, include=[1, 2], exclude=[3])[0] == 2
    assert call_topn(L, dt, a, B, None, 0.0, 0.0, 3, 5, exclude=[11])[0] == 2
    assert call_topn(L, dt, a, B, None, 0.0, 0.0, 3, 8, exclude=[1, 2, 3])[0] == 2   # not enough items left
    assert call_topn(L, dt, np.array([1.0, np.nan, 1.0]), B, None, 0.0, 0.0, 3, 2)[0] == 2


def call_most_popular(lib, dt, ixA, ixB, X, m, n, *, user_bias, implicit, lam=1.5, scale_lam=False, alpha=1.0,
                      adjust_weight=False, apply_log_transf=False, center=True):
    bA = np.zeros(m, dt) if user_bias else None
    bB = np.zeros(n, dt)
    g = np.zeros(1, dt)
    wm = np.ones(1, dt)
    ia = np.ascontiguousarray(ixA, np.int32).copy(); ib = np.ascontiguousarray(ixB, np.int32).copy()
    x = np.ascontiguousarray(X, dt).copy()
    rc = lib.fit_most_popular(ptr(bA), ptr(bB), ptr(g) if center else None, lam, lam * 0.7, scale_lam, False, alpha, m, n,
                              ptr(ia), ptr(ib), ptr(x), x.size, None, None, implicit, adjust_weight, apply_log_transf,
                              False, False, ptr(wm), 4)
    return rc, bA, bB, g[0], wm[0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("scale_lam", [False, True])
def test_most_popular_explicit(gpu_libs, dtype, scale_lam):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n = 900, 600
    ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=8)
    o = call_most_popular(L, dt, ixA, ixB, X, m, n, user_bias=True, implicit=False, scale_lam=scale_lam)
    r = call_most_popular(R, dt, ixA, ixB, X, m, n, user_bias=True, implicit=False, scale_lam=scale_lam)
    assert o[0] == 0 and r[0] == 0
    assert o[3] == r[3]
    if dt == np.float32:
        assert np.array_equal(o[1], r[1]) and np.array_equal(o[2], r[2])
    else:
        assert np.allclose(o[1], r[1], rtol=0, atol=4 * np.finfo(dt).eps * np.abs(r[1]).max())
        assert np.allclose(o[2], r[2], rtol=0, atol=4 * np.finfo(dt).eps * np.abs(r[2]).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("opts", [dict(), dict(alpha=40.0), dict(adjust_weight=True), dict(apply_log_transf=True)])
def test_most_popular_implicit(gpu_libs, dtype, opts):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n = 900, 600
    ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=9, kind="counts")
    o = call_most_popular(L, dt, ixA, ixB, X, m, n, user_bias=False, implicit=True, center=False, **opts)
    r = call_most_popular(R, dt, ixA, ixB, X, m, n, user_bias=False, implicit=True, center=False, **opts)
    assert o[0] == 0 and r[0] == 0
    assert o[4] == r[4]
    assert np.allclose(o[2], r[2], rtol=0, atol=4 * np.finfo(dt).eps * np.abs(r[2]).max())


def test_most_popular_refuses_what_it_does_not_cover(gpu_libs):
    dt = np.dtype(np.float64)
    L = gpu_libs[dt]
    ixA, ixB, X = synth_coo(50, 40, 300, dt, seed=1)
    assert call_most_popular(L, dt, ixA, ixB, X, 50, 40, user_bias=True, implicit=True)[0] == 2
