"""Shared helpers for the parity tests: synthetic inputs, and thin callers that drive the product library
and the reference build (oracle/_ref) through the SAME ctypes argument lists."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from refload import ptr, ref  # noqa: E402,F401


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


def fit_explicit(lib, dtype, ixA, ixB, X, m, n, k, *, lam=0.05, user_bias=True, item_bias=True, center=True,
                 scale_lam=False, niter=3, use_cg=True, max_cg_steps=3, finalize_chol=False, seed=1, nthreads=4,
                 w_main=1.0, lam_unique=None, precompute=False, k_main=0, U=None, I=None, add_implicit_features=False,
                 w_user=1.0, w_item=1.0, w_implicit=1.0, center_side=True, copy_inputs=True, out=None):
    """Call fit_collective_explicit_als (reference src/cmfrec.h:1851) on `lib`; returns dict of outputs.
    copy_inputs=False passes the caller's index / value arrays as they are (the reference may overwrite X; this
    repo's library never writes to its inputs); `out` may hold preallocated A, B, biasA, biasB (e.g. pinned memory)."""
    dt = np.dtype(dtype)
    kk = k + k_main
    out = out or {}
    A = out["A"] if "A" in out else np.zeros((m, kk), dt)
    B = out["B"] if "B" in out else np.zeros((n, kk), dt)
    biasA = out["biasA"] if "biasA" in out else np.zeros(m, dt)
    biasB = out["biasB"] if "biasB" in out else np.zeros(n, dt)
    glob_mean = np.zeros(1, dt)
    sA = np.zeros(1, dt)
    sB = np.zeros(1, dt)
    lu = None if lam_unique is None else np.asarray(lam_unique, dt)
    ub = int(user_bias)
    has_bias = user_bias or item_bias
    Bpb = np.zeros((n, kk + 1), dt) if (precompute and has_bias) else None
    BtB = np.zeros((kk + ub, kk + ub), dt) if precompute else None
    TBt = np.zeros((n, kk + ub), dt) if precompute else None
    if copy_inputs:
        ixA = np.ascontiguousarray(ixA, np.int32).copy()
        ixB = np.ascontiguousarray(ixB, np.int32).copy()
        X = np.ascontiguousarray(X, dt).copy()
    else:
        assert ixA.dtype == np.int32 and ixB.dtype == np.int32 and X.dtype == dt
        assert ixA.flags.c_contiguous and ixB.flags.c_contiguous and X.flags.c_contiguous
    p = 0 if U is None else U.shape[1]
    q = 0 if I is None else I.shape[1]
    Uc = None if U is None else np.ascontiguousarray(U, dt).copy()
    Ic = None if I is None else np.ascontiguousarray(I, dt).copy()
    C = np.zeros((p, k), dt) if p else None
    D = np.zeros((q, k), dt) if q else None
    Ai = np.zeros((m, k), dt) if add_implicit_features else None
    Bi = np.zeros((n, k), dt) if add_implicit_features else None
    Ucm = np.zeros(p, dt) if (p and center_side) else None
    Icm = np.zeros(q, dt) if (q and center_side) else None
    collective = bool(p or q or add_implicit_features)
    BiTBi = np.zeros((kk, kk), dt) if (precompute and add_implicit_features) else None
    TCt = np.zeros((p, kk), dt) if (precompute and p) else None
    CtCw = np.zeros((kk, kk), dt) if (precompute and p) else None
    BeChol = np.zeros((kk + ub, kk + ub), dt) if (precompute and collective) else None
    rc = lib.fit_collective_explicit_als(
        ptr(biasA) if user_bias else None, ptr(biasB) if item_bias else None, ptr(A), ptr(B), ptr(C), ptr(D), ptr(Ai), ptr(Bi),
        add_implicit_features, True, seed, ptr(glob_mean), ptr(Ucm), ptr(Icm), m, n, k, ptr(ixA), ptr(ixB), ptr(X), X.size,
        None, None, user_bias, item_bias, center, lam, ptr(lu), 0.0, None, scale_lam, False, False, ptr(sA), ptr(sB),
        ptr(Uc), m if p else 0, p, ptr(Ic), n if q else 0, q, None, None, None, 0, None, None, None, 0, False, False, False,
        k_main, 0, 0, w_main, w_user, w_item, w_implicit, niter, nthreads, False, False, use_cg, max_cg_steps, False,
        finalize_chol, False, 100, False, False, precompute, True, ptr(Bpb), ptr(BtB), ptr(TBt), None, ptr(BeChol), ptr(BiTBi),
        ptr(TCt), ptr(CtCw), None)
    return dict(rc=rc, A=A, B=B, biasA=biasA, biasB=biasB, glob_mean=glob_mean[0], B_plus_bias=Bpb, BtB=BtB,
                TransBtBinvBt=TBt, C=C, D=D, Ai=Ai, Bi=Bi, U_colmeans=Ucm, I_colmeans=Icm, BeTBeChol=BeChol, BiTBi=BiTBi,
                TransCtCinvCt=TCt, CtCw=CtCw)


def fit_implicit(lib, dtype, ixA, ixB, X, m, n, k, *, lam=5.0, alpha=1.0, niter=3, use_cg=True, max_cg_steps=3,
                 finalize_chol=False, seed=1, nthreads=4, w_main=1.0, adjust_weight=False, apply_log_transf=False,
                 precompute=False, k_main=0, copy_inputs=True, out=None, U=None, I=None, w_user=1.0, w_item=1.0,
                 center_side=True, lam_unique=None):
    """Call fit_collective_implicit_als (reference src/cmfrec.h:1893) on `lib` (copy_inputs / out: see fit_explicit);
    U [m x p] / I [n x q]: dense side information."""
    dt = np.dtype(dtype)
    kk = k + k_main
    out = out or {}
    A = out["A"] if "A" in out else np.zeros((m, kk), dt)
    B = out["B"] if "B" in out else np.zeros((n, kk), dt)
    wmm = np.zeros(1, dt)
    BtB = np.zeros((kk, kk), dt) if precompute else None
    if copy_inputs:
        ixA = np.ascontiguousarray(ixA, np.int32).copy()
        ixB = np.ascontiguousarray(ixB, np.int32).copy()
        X = np.ascontiguousarray(X, dt).copy()
    p = 0 if U is None else U.shape[1]
    q = 0 if I is None else I.shape[1]
    Uc = None if U is None else np.ascontiguousarray(U, dt).copy()
    Ic = None if I is None else np.ascontiguousarray(I, dt).copy()
    Cm = np.zeros((p, k), dt) if p else None
    Dm = np.zeros((q, k), dt) if q else None
    Ucm = np.zeros(p, dt) if (p and center_side) else None
    Icm = np.zeros(q, dt) if (q and center_side) else None
    lu = None if lam_unique is None else np.asarray(lam_unique, dt)
    BeTBe = np.zeros((kk, kk), dt) if (precompute and p) else None
    BeChol = np.zeros((kk, kk), dt) if (precompute and p) else None
    rc = lib.fit_collective_implicit_als(
        ptr(A), ptr(B), ptr(Cm), ptr(Dm), True, seed, ptr(Ucm), ptr(Icm), m, n, k, ptr(ixA), ptr(ixB), ptr(X), X.size,
        lam, ptr(lu), 0.0, None, ptr(Uc), m if p else 0, p, ptr(Ic), n if q else 0, q, None, None, None, 0, None, None, None, 0,
        False, False, k_main, 0, 0, w_main, w_user, w_item, ptr(wmm), alpha, adjust_weight, apply_log_transf, niter, nthreads,
        False, False, use_cg, max_cg_steps, False, finalize_chol, False, 100, False, False, precompute,
        ptr(BtB), ptr(BeTBe), ptr(BeChol), None)
    return dict(rc=rc, A=A, B=B, w_main_multiplier=wmm[0], BtB=BtB, C=Cm, D=Dm, U_colmeans=Ucm, I_colmeans=Icm, BeTBe=BeTBe,
                BeTBeChol=BeChol)


def csr_csc(lib, dtype, ixA, ixB, X, m, n):
    """COO -> (csr_p, csr_i, csr_v, csc_p, csc_i, csc_v) through the product's host routine."""
    dt = np.dtype(dtype)
    nnz = X.size
    out = (np.zeros(m + 1, np.uint64), np.zeros(nnz, np.int32), np.zeros(nnz, dt),
           np.zeros(n + 1, np.uint64), np.zeros(nnz, np.int32), np.zeros(nnz, dt))
    lib.cmfb200_coo_to_csr_and_csc(ptr(ixA), ptr(ixB), ptr(X), m, n, nnz, *[ptr(t) for t in out])
    return out


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


class AlsSession:
    """Context manager around cmfb200_als_* (include/cmfrec_b200.h PART 2)."""

    def __init__(self, lib, dtype, csr, csc, m, n, k, *, implicit, user_bias=False, item_bias=False, lam_A=0.0,
                 lam_B=0.0, lam_biasA=None, lam_biasB=None, scale_lam=False, max_cg_steps=3):
        self.lib, self.dt = lib, np.dtype(dtype)
        self.m, self.n, self.k = m, n, k
        opt = lib.AlsOptions()
        opt.implicit = int(implicit)
        opt.m, opt.n, opt.k = m, n, k
        opt.user_bias, opt.item_bias = int(user_bias), int(item_bias)
        opt.lam_A, opt.lam_B = lam_A, lam_B
        opt.lam_biasA = lam_A if lam_biasA is None else lam_biasA
        opt.lam_biasB = lam_B if lam_biasB is None else lam_biasB
        opt.scale_lam = int(scale_lam)
        opt.max_cg_steps = max_cg_steps
        opt.rank, opt.world = 0, 1
        opt.nccl_id = None
        opt.stream = None
        self.h = C.c_void_p()
        rc = lib.cmfb200_als_create(C.byref(self.h), C.byref(opt), *[ptr(t) for t in csr], *[ptr(t) for t in csc])
        if rc:
            raise RuntimeError("cmfb200_als_create failed with code %d" % rc)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.lib.cmfb200_als_destroy(self.h)
        self.h = None

    def set_factors(self, A, biasA, B, biasB):
        rc = self.lib.cmfb200_als_set_factors(self.h, ptr(A), ptr(biasA), ptr(B), ptr(biasB))
        assert rc == 0, rc

    def get_factors(self, with_bias=False):
        A = np.zeros((self.m, self.k), self.dt)
        B = np.zeros((self.n, self.k), self.dt)
        bA = np.zeros(self.m, self.dt)
        bB = np.zeros(self.n, self.dt)
        rc = self.lib.cmfb200_als_get_factors(self.h, ptr(A), ptr(bA), ptr(B), ptr(bB))
        assert rc == 0, rc
        return (A, bA, B, bB) if with_bias else (A, B)

    def half_sweep(self, which, it, solver):
        rc = self.lib.cmfb200_als_half_sweep(self.h, which, it, solver)
        assert rc == 0, rc
        assert self.lib.cmfb200_als_sync(self.h) == 0


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
