"""Shared helpers for the parity tests: synthetic inputs, and thin callers that drive the product library
and the reference build (oracle/_ref) through the SAME ctypes argument lists."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from refload import ref  # noqa: E402,F401
from cmfrec_b200.calls import AlsSession, csr_csc, fit_explicit, fit_implicit, ptr  # noqa: E402,F401


def synth_coo(m, n, nnz, dtype, seed=0, kind="ratings", dedup=True, zipf=True):
    """Random COO triplets, sorted by (row, col).  kind: 'ratings' (0.5..5.0) or 'counts' (>=1)."""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, m, size=nnz)
    if zipf:
        w = 1.0 / np.arange(1, n + 1) ** 0.8
        cols = rng.choice(n, size=nnz, p=w / w.sum())
        cols = rng.permutation(n)[cols]
    else:
        cols = rng.integers(0, n, size=nnz)
    if dedup:
        key = np.unique(rows.astype(np.int64) * n + cols)
        rows, cols = key // n, key % n
    else:
        o = np.lexsort((cols, rows))
        rows, cols = rows[o], cols[o]
    if kind == "ratings":
        vals = rng.integers(1, 11, size=rows.size) * 0.5
    else:
        vals = np.ceil(rng.lognormal(1.0, 1.5, size=rows.size))
    return rows.astype(np.int32), cols.astype(np.int32), vals.astype(dtype)


def ref_optimizeA(R, dtype, A, B, ptr_, idx, val, *, lam, lam_last, scale_lam, use_cg, max_cg_steps, nthreads=4):
    """reference optimizeA (src/common.c:2742) on sparse X, missing-as-unknown: updates A [m x k'] in place."""
    dt = np.dtype(dtype)
    m, kd = A.shape
    n = B.shape[0]
    assert B.shape[1] == kd
    buf = np.zeros(nthreads * (kd * kd + 8 * kd) + 16, dt)
    filled = C.c_bool(False)
    R.optimizeA(ptr(A), kd, ptr(B), kd, m, n, kd, ptr(ptr_), ptr(idx), ptr(val), None, 0, False, False, False,
                None, None, False, lam, lam_last, 0.0, 0.0, scale_lam, False, None, False, nthreads, False,
                use_cg, False, max_cg_steps, False, 0, None, None, None, 0.0, None, 1.0, False, None,
                C.byref(filled), ptr(buf), None)


def ref_optimizeA_implicit(R, dtype, A, B, ptr_, idx, val, *, lam, use_cg, max_cg_steps, nthreads=4):
    """reference optimizeA_implicit (src/common.c:3305): updates A [m x k] in place."""
    dt = np.dtype(dtype)
    m, k = A.shape
    n = B.shape[0]
    buf = np.zeros(k * k + nthreads * (k * k + 8 * k) + 16, dt)
    R.optimizeA_implicit(ptr(A), k, ptr(B), k, m, n, k, ptr(ptr_), ptr(idx), ptr(val), lam, 0.0, nthreads, False,
                         use_cg, False, max_cg_steps, False, 0, None, ptr(buf), None)


def rel_err(x, y):
    """max |x - y| / max |y| (the form the parity tolerances in SURVEY.md 8d are stated in)."""
    d = np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).max() if np.size(x) else 0.0
    s = np.abs(np.asarray(y, np.float64)).max() if np.size(y) else 1.0
    return d / max(s, 1e-300)


def frac_rows_within(x, y, tol):
    """fraction of rows whose max abs difference is <= tol * max|y|"""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    s = np.abs(y).max()
    d = np.abs(x - y).reshape(x.shape[0], -1).max(axis=1)
    return float((d <= tol * s).mean())


def rows_match(x, y, tol, outlier_frac=0.001, outlier_cap=0.25):
    """True when all but `outlier_frac` of the rows (at least one row is always allowed) agree to
    tol * max|y|.  Outliers exist because the CG exits on absolute ||r||^2 thresholds (1e-12 / 1e-8): a row
    sitting on a threshold can legitimately take one step more or fewer under a different summation order."""
    x = np.asarray(x, np.float64).reshape(np.shape(x)[0], -1)
    y = np.asarray(y, np.float64).reshape(np.shape(y)[0], -1)
    if not (np.isfinite(x).all() and np.isfinite(y).all()):
        return False                      # a NaN row must never pass as a match
    s = max(np.abs(y).max(), 1e-300)
    err = np.abs(x - y).max(axis=1)
    bad = int((~(err <= tol * s)).sum())
    if err.max() > outlier_cap * s:       # outliers (CG step-count flips) are bounded too
        return False
    return bad <= max(1, int(np.ceil(outlier_frac * x.shape[0])))


def implicit_objective(ixA, ixB, X, A, B, lam, alpha=1.0):
    """WRMF loss the implicit fit minimises, in float64: sum_all (p - a.b)^2 + sum_nz x (1 - a.b)^2-ish form
    written through the Gram trick: sum_all (a.b)^2 = <A^T A, B^T B>."""
    A = np.asarray(A, np.float64); B = np.asarray(B, np.float64); x = np.asarray(X, np.float64) * alpha
    pred = np.einsum("ij,ij->i", A[ixA], B[ixB])
    all_sq = np.sum((A.T @ A) * (B.T @ B))
    loss = all_sq + np.sum((x + 1.0) * (1.0 - pred) ** 2 - pred ** 2)
    return loss + lam * (np.sum(A * A) + np.sum(B * B))


def explicit_objective(ixA, ixB, X, out, lam):
    """squared error on the stored entries + L2 penalty, float64"""
    A = np.asarray(out["A"], np.float64); B = np.asarray(out["B"], np.float64)
    pred = np.einsum("ij,ij->i", A[ixA], B[ixB]) + float(out["glob_mean"])
    pred = pred + np.asarray(out["biasA"], np.float64)[ixA] + np.asarray(out["biasB"], np.float64)[ixB]
    err = np.asarray(X, np.float64) - pred
    return np.sum(err ** 2) + lam * (np.sum(A * A) + np.sum(B * B))


def cg_residual_trace(Gd, x, a0, lam_vec, steps, BtB=None):
    """||r||^2 before the first step and after every step of the reference's truncated CG on ONE row, in float64
    (explicit: src/common.c:1098-1188; implicit when `BtB` is given: src/common.c:1914-1986, residual as written
    at :1936-1942).  Gd [nnz x kd] are the gathered opposing rows, lam_vec the per-coordinate regulariser."""
    Gd = np.asarray(Gd, np.float64); x = np.asarray(x, np.float64); a = np.asarray(a0, np.float64).copy()
    lam_vec = np.asarray(lam_vec, np.float64)
    d = Gd @ a
    if BtB is None:
        r = Gd.T @ (x - d) - lam_vec * a
    else:
        r = -(BtB @ a) + Gd.T @ (-(d - 1.0) * x - d) - lam_vec * a
    out = [float(r @ r)]
    if out[0] <= 1e-12:
        return out
    p = r.copy()
    r_old = out[0]
    for _ in range(steps):
        dp = Gd @ p
        Ap = Gd.T @ (dp if BtB is None else dp * (x - 1.0) + dp) + lam_vec * p
        if BtB is not None:
            Ap += BtB @ p
        al = r_old / (p @ Ap)
        a += al * p
        r -= al * Ap
        r_new = float(r @ r)
        out.append(r_new)
        if r_new <= 1e-8:
            break
        p = r + (r_new / r_old) * p
        r_old = r_new
    return out


def near_cg_threshold(trace, factor=30.0):
    """True when some ||r||^2 of the trace lies within `factor` of one of the CG's absolute exit thresholds
    (1e-12 before the first step, 1e-8 after a step): a float32 run may then legitimately take one step more
    or fewer than the reference (the float32 ||r||^2 of such a row carries a relative error of 1e-2 and more,
    because r is a difference of sums that are 1e3-1e5 times larger)."""
    if 1e-12 / factor <= trace[0] <= 1e-12 * factor:
        return True
    return any(1e-8 / factor <= v <= 1e-8 * factor for v in trace[1:])
