"""Serving on the GPU: the tensor-core product, predictions for (row, column) pairs and top-N for many users at once,
against the reference build (predict_multiple src/common.c:5066, topN src/common.c:5127, the collective-model wrappers
src/collective.c:11546-11614, 11797-11862).  Rankings are integer work: indices must agree exactly wherever
neighbouring scores differ by more than 1e-5 (fp32) / 1e-12 (fp64) relative -- the reference's own exact test is
test_math/test_topN.py:78-81."""
import ctypes as C

import numpy as np
import pytest

from support import ptr, ref
from test_gpu_popular_topn import call_topn, same_ranking

pytestmark = pytest.mark.gpu


def _ref(dt):
    R = ref(dt)
    if R is None:
        pytest.fail("oracle/_ref is not built")
    return R


@pytest.mark.parametrize("M,N,K,pad", [(300, 1000, 64, 0), (128, 256, 256, 0), (5, 17, 3, 0), (1000, 333, 130, 2), (129, 40, 65, 1),
                                       (4096, 512, 32, 0)])
def test_tensor_core_product_matches_fp64(gpu_libs, M, N, K, pad):
    """C = A B^T with the 3xTF32 split: every entry within 2e-6 of |a| |b| (plain TF32 would sit at 5e-4)"""
    L = gpu_libs[np.dtype(np.float32)]
    rng = np.random.default_rng(M + N + K)
    A = np.zeros((M, K + pad), np.float32); B = np.zeros((N, K + pad), np.float32)
    A[:, :K] = rng.normal(size=(M, K)); B[:, :K] = rng.normal(size=(N, K))
    A[:, K:] = 7.0; B[:, K:] = -3.0          # columns beyond K must not be read as part of the product
    Cc = np.zeros((M, N), np.float32)
    assert L.cmfb200_gemm_nt(ptr(A), K + pad, M, ptr(B), K + pad, N, K, ptr(Cc), 0, None) == 0
    exact = A[:, :K].astype(np.float64) @ B[:, :K].astype(np.float64).T
    bound = np.linalg.norm(A[:, :K].astype(np.float64), axis=1)[:, None] * np.linalg.norm(B[:, :K].astype(np.float64), axis=1)[None, :]
    assert np.isfinite(Cc).all()
    assert (np.abs(Cc - exact) <= 2e-6 * bound + 1e-30).all(), np.abs(Cc - exact).max()
    # linearity in A (size-independent property): (2A) B^T == 2 (A B^T) exactly in binary floating point
    C2 = np.zeros((M, N), np.float32)
    A2 = (2 * A).astype(np.float32)
    assert L.cmfb200_gemm_nt(ptr(A2), K + pad, M, ptr(B), K + pad, N, K, ptr(C2), 0, None) == 0
    assert np.array_equal(C2, 2 * Cc)


def test_tensor_core_product_is_fp32_only(gpu_libs):
    L = gpu_libs[np.dtype(np.float64)]
    A = np.ones((4, 4)); B = np.ones((4, 4)); Cc = np.zeros((4, 4))
    assert L.cmfb200_gemm_nt(ptr(A), 4, 4, ptr(B), 4, 4, 4, ptr(Cc), 0, None) == 3


def _factors(dt, m, n, k, seed, k_user=0, k_item=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(m, k_user + k)).astype(dt)
    B = rng.normal(size=(n, k_item + k)).astype(dt)
    return A, B, rng.normal(size=m).astype(dt), rng.normal(size=n).astype(dt), rng


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,k,k_user,k_item,npred", [(500, 300, 16, 0, 0, 4000), (70, 40, 5, 2, 3, 30), (3000, 2000, 64, 0, 0, 100)])
def test_predict_multiple_matches_reference(gpu_libs, dtype, m, n, k, k_user, k_item, npred):
    dt = np.dtype(dtype)
    Lb, R = gpu_libs[dt], _ref(dt)
    A, B, bA, bB, rng = _factors(dt, m, n, k, 5, k_user, k_item)
    row = rng.integers(0, m, size=npred).astype(np.int32)
    col = rng.integers(0, n, size=npred).astype(np.int32)
    row[::7] = m + 3          # unknown ids -> NaN (predict_multiple) / mean + known bias (predict_X_old_collective_explicit)
    col[3::11] = -1
    tol = (1e-12 if dt == np.float64 else 2e-5)
    for use_bias in (True, False):
        o = np.zeros(npred, dt); r = np.zeros(npred, dt)
        args = lambda out: (ptr(A), k_user, ptr(B), k_item, ptr(bA) if use_bias else None, ptr(bB) if use_bias else None, 0.25, k, 0,
                            m, n, ptr(row), ptr(col), npred, ptr(out), 4)
        assert Lb.predict_multiple(*args(o)) == 0
        R.predict_multiple(*args(r))
        assert np.array_equal(np.isnan(o), np.isnan(r))
        good = ~np.isnan(r)
        assert np.abs(o[good] - r[good]).max() <= tol * max(1.0, np.abs(r[good]).max())
    # the collective-model wrapper: no NaN left (col < 0 is not guarded by the reference either: keep those out)
    col2 = np.where(col < 0, 0, col).astype(np.int32)
    o = np.zeros(npred, dt); r = np.zeros(npred, dt)
    args = lambda out: (ptr(row), ptr(col2), ptr(out), npred, ptr(A), ptr(bA), ptr(B), ptr(bB), 0.25, k, k_user, k_item, 0, m, n, 4)
    assert Lb.predict_X_old_collective_explicit(*args(o)) == 0
    assert R.predict_X_old_collective_explicit(*args(r)) == 0
    assert np.isfinite(o).all() and np.abs(o - r).max() <= tol * max(1.0, np.abs(r).max())
    row2 = np.where(row >= m, 0, row).astype(np.int32)
    o = np.zeros(npred, dt); r = np.zeros(npred, dt)
    args = lambda out: (ptr(row2), ptr(col2), ptr(out), npred, ptr(A), ptr(B), k, k_user, k_item, 0, m, n, 4)
    assert Lb.predict_X_old_collective_implicit(*args(o)) == 0
    assert R.predict_X_old_collective_implicit(*args(r)) == 0
    assert np.abs(o - r).max() <= tol * max(1.0, np.abs(r).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_topn_old_wrappers_match_reference(gpu_libs, dtype):
    dt = np.dtype(dtype)
    Lb, R = gpu_libs[dt], _ref(dt)
    m, n, k = 50, 4000, 24
    A, B, bA, bB, rng = _factors(dt, m, n, k, 9)
    rtol = 1e-12 if dt == np.float64 else 1e-5
    exc = np.sort(rng.choice(n, size=300, replace=False)).astype(np.int32)
    for row_index in (0, 17, 49):
        res = []
        for lib in (Lb, R):
            ix = np.zeros(20, np.int32); sc = np.zeros(20, dt); e = exc.copy()
            rc = lib.topN_old_collective_explicit(None, 0.0, ptr(A), ptr(bA), row_index, ptr(B), ptr(bB), 0.5, k, 0, 0, 0, None, 0,
                                                  ptr(e), e.size, ptr(ix), ptr(sc), 20, n, n, False, 4)
            assert rc == 0
            res.append((ix, sc))
        assert same_ranking(res[0][0], res[0][1], res[1][0], res[1][1], rtol)
        res = []
        for lib in (Lb, R):
            ix = np.zeros(20, np.int32); sc = np.zeros(20, dt)
            rc = lib.topN_old_collective_implicit(None, ptr(A), row_index, ptr(B), k, 0, 0, 0, None, 0, None, 0, ptr(ix), ptr(sc), 20, n, 4)
            assert rc == 0
            res.append((ix, sc))
        assert same_ranking(res[0][0], res[0][1], res[1][0], res[1][1], rtol)


def serve_topn(lib, dt, A, B, bA, bB, glob_mean, k, users, seen, n_top, k_user=0, k_item=0, want_scores=True):
    m, n = A.shape[0], B.shape[0]
    rc = C.c_int(0)
    h = lib.cmfb200_serve_create(ptr(A), m, k_user, ptr(B), n, k_item, ptr(bA), ptr(bB), glob_mean, k, 0, C.byref(rc))
    assert rc.value == 0 and h
    users = np.ascontiguousarray(users, np.int32)
    ix = np.full((users.size, n_top), -1, np.int32)
    sc = np.zeros((users.size, n_top), dt) if want_scores else None
    sp = si = None
    if seen is not None:
        sp = np.zeros(users.size + 1, np.uint64)
        sp[1:] = np.cumsum([len(s) for s in seen])
        si = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in seen]) if sp[-1] else np.zeros(0, np.int32), np.int32)
    ms = C.c_float(0)
    rc2 = lib.cmfb200_serve_topn(h, ptr(users), users.size, ptr(sp), ptr(si), n_top, ptr(ix), ptr(sc), C.byref(ms))
    lib.cmfb200_serve_destroy(h)
    return rc2, ix, sc, ms.value


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,k,n_users,n_top", [(400, 3000, 32, 400, 10), (60, 517, 7, 25, 50), (3000, 20000, 64, 300, 100), (10, 40, 3, 10, 40)])
def test_batched_topn_equals_the_references_topn_per_user(gpu_libs, dtype, m, n, k, n_users, n_top):
    dt = np.dtype(dtype)
    Lb, R = gpu_libs[dt], _ref(dt)
    A, B, bA, bB, rng = _factors(dt, m, n, k, m + n)
    users = rng.choice(m, size=n_users, replace=n_users > m)
    max_seen = max(0, min(n - n_top, n // 3))
    seen = [np.sort(rng.choice(n, size=int(rng.integers(0, max_seen + 1)), replace=False)) for _ in users]
    rtol = 1e-12 if dt == np.float64 else 1e-5
    for use_bias in (True, False):
        rc, ix, sc, _ = serve_topn(Lb, dt, A, B, bA if use_bias else None, bB if use_bias else None, 0.3, k, users, seen, n_top)
        assert rc == 0
        for j, u in enumerate(users):
            rc2, rix, rsc = call_topn(R, dt, A[u], B, bB if use_bias else None, 0.3, float(bA[u]) if use_bias else 0.0, k, n_top,
                                      exclude=seen[j] if len(seen[j]) else None)
            assert rc2 == 0
            assert same_ranking(ix[j], sc[j], rix, rsc, rtol), (j, ix[j][:8], rix[:8])
            assert np.abs(sc[j] - rsc).max() <= (1e-12 if dt == np.float64 else 2e-5) * max(1.0, np.abs(rsc).max())
            assert np.all(np.diff(sc[j]) <= 0) and len(set(ix[j].tolist())) == n_top
            assert not (set(ix[j].tolist()) & set(seen[j].tolist()))


def test_batched_topn_breaks_ties_by_item_id_and_checks_arguments(gpu_libs):
    dt = np.dtype(np.float32)
    Lb = gpu_libs[dt]
    m, n, k = 6, 1000, 4
    A = np.ones((m, k), dt); B = np.zeros((n, k), dt)
    B[::3] = 1.0                                   # a third of the items tie at the top, the rest tie at zero
    rc, ix, sc, _ = serve_topn(Lb, dt, A, B, None, None, 0.0, k, np.arange(m), None, 400)
    assert rc == 0
    expect = np.concatenate([np.arange(0, n, 3), np.setdiff1d(np.arange(n), np.arange(0, n, 3))])[:400]
    assert all(np.array_equal(ix[j], expect) for j in range(m))
    assert serve_topn(Lb, dt, A, B, None, None, 0.0, k, [0, m], None, 5)[0] == 2          # unknown user
    assert serve_topn(Lb, dt, A, B, None, None, 0.0, k, [0], None, 0)[0] == 2             # n_top == 0
    assert serve_topn(Lb, dt, A, B, None, None, 0.0, k, [0], [np.arange(998)], 5)[0] == 2  # not enough items left


def test_batched_topn_at_serving_scale(gpu_libs):
    """LastFM-sized item set, 2048 users in one call: properties that do not need the oracle (sorted, distinct, unseen,
    and the best score equals the maximum of a NumPy product for a sample of users)"""
    dt = np.dtype(np.float32)
    Lb = gpu_libs[dt]
    m, n, k = 5000, 160112, 64
    A, B, bA, bB, rng = _factors(dt, m, n, k, 77)
    users = rng.choice(m, size=2048, replace=False)
    seen = [rng.choice(n, size=48, replace=False) for _ in users]
    rc, ix, sc, ms = serve_topn(Lb, dt, A, B, None, bB, 0.0, k, users, seen, 10)
    assert rc == 0 and ms > 0
    assert (np.diff(sc, axis=1) <= 0).all()
    for j in range(0, 2048, 97):
        full = A[users[j]].astype(np.float64) @ B.astype(np.float64).T + bB
        full[seen[j]] = -np.inf
        best = np.argsort(-full, kind="stable")[:10]
        assert same_ranking(ix[j], sc[j].astype(np.float64), best, full[best], 1e-5)
        assert not (set(ix[j].tolist()) & set(seen[j].tolist()))
