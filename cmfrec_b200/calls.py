"""Host-side callers of the C ABI (include/cmfrec_b200.h) with numpy buffers: the reference-named fit entry points with
the reference's argument lists (src/cmfrec.h:1851, :1893) and the device-resident ALS state.  The SAME functions drive
the reference build in the parity tests (tests/support.py passes the oracle/_ref library instead of the product), which
is what makes those tests read like the reference's own; bench.py's end-to-end leg calls them on the product library.
Nothing here computes: every function forwards to the shared library it is given."""
import ctypes as C

import numpy as np

P = C.c_void_p


def ptr(a):
    return None if a is None else a.ctypes.data_as(P)


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
