"""TEST INFRASTRUCTURE -- python face of oracle/cmf_oracle.c (this repo's plain-C restatement of the reference's
ALS path).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}
P = C.c_void_p


def _lib(dtype):
    dt = np.dtype(dtype)
    tag = "f32" if dt == np.float32 else "f64"
    if tag in _libs:
        return _libs[tag]
    path = os.path.join(HERE, "_build", "libcmf_oracle_%s.so" % tag)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    real = C.c_float if tag == "f32" else C.c_double
    lib.oracle_random_init.argtypes = [P, C.c_size_t, P, C.c_size_t, C.c_int, C.c_bool]
    lib.oracle_random_init.restype = None
    lib.oracle_coo_to_csr.argtypes = [P, P, P, C.c_int, C.c_size_t, P, P, P]
    lib.oracle_coo_to_csr.restype = None
    lib.oracle_global_mean.argtypes = [P, C.c_size_t, C.c_int]
    lib.oracle_global_mean.restype = real
    lib.oracle_init_biases_twosided.argtypes = [C.c_int, C.c_int, P, P, P, P, P, P, real, real, C.c_bool, P, P]
    lib.oracle_init_biases_twosided.restype = None
    lib.oracle_optimizeA.argtypes = [P, C.c_int, P, C.c_int, C.c_int, C.c_int, P, P, P, real, real, C.c_bool, C.c_bool,
                                     C.c_int]
    lib.oracle_optimizeA.restype = None
    lib.oracle_optimizeA_implicit.argtypes = [P, C.c_size_t, P, C.c_size_t, C.c_int, C.c_int, C.c_int, P, P, P, real,
                                              C.c_bool, C.c_int]
    lib.oracle_optimizeA_implicit.restype = None
    lib.oracle_fit_explicit.argtypes = [P, P, P, P, C.c_int, P, C.c_int, C.c_int, C.c_int, P, P, P, C.c_size_t,
                                        C.c_bool, C.c_bool, C.c_bool, real, P, C.c_bool, real, C.c_int, C.c_int,
                                        C.c_bool, C.c_int, C.c_bool]
    lib.oracle_fit_explicit.restype = C.c_int
    lib.oracle_fit_implicit.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, C.c_size_t, real, real, P,
                                        real, C.c_bool, C.c_bool, C.c_int, C.c_bool, C.c_int, C.c_bool]
    lib.oracle_fit_implicit.restype = C.c_int
    _libs[tag] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


def random_init(dtype, sizeA, sizeB, seed, normal):
    dt = np.dtype(dtype)
    A = np.zeros(sizeA, dt)
    B = np.zeros(max(sizeB, 1), dt)
    _lib(dt).oracle_random_init(_p(A), sizeA, _p(B) if sizeB else None, sizeB, seed, normal)
    return A, B[:sizeB]


def coo_to_csr(dtype, row, col, val, m):
    dt = np.dtype(dtype)
    nnz = val.size
    p = np.zeros(m + 1, np.uint64)
    i = np.zeros(nnz, np.int32)
    v = np.zeros(nnz, dt)
    _lib(dt).oracle_coo_to_csr(_p(np.ascontiguousarray(row, np.int32)), _p(np.ascontiguousarray(col, np.int32)),
                               _p(np.ascontiguousarray(val, dt)), m, nnz, _p(p), _p(i), _p(v))
    return p, i, v


def global_mean(dtype, X, nthreads):
    return _lib(dtype).oracle_global_mean(_p(np.ascontiguousarray(X, dtype)), X.size, nthreads)


def init_biases_twosided(dtype, m, n, csr, csc, lam_user, lam_item, scale_lam):
    dt = np.dtype(dtype)
    bA = np.zeros(m, dt)
    bB = np.zeros(n, dt)
    _lib(dt).oracle_init_biases_twosided(m, n, *[_p(t) for t in csr], *[_p(t) for t in csc], lam_user, lam_item,
                                         scale_lam, _p(bA), _p(bB))
    return bA, bB


def optimizeA(dtype, A, B, ptr_, idx, val, *, lam, lam_last, scale_lam, use_cg, max_cg_steps):
    m, kd = A.shape
    _lib(dtype).oracle_optimizeA(_p(A), kd, _p(B), B.shape[1], m, kd, _p(ptr_), _p(idx), _p(val), lam, lam_last,
                                 scale_lam, use_cg, max_cg_steps)


def optimizeA_implicit(dtype, A, B, ptr_, idx, val, *, lam, use_cg, max_cg_steps):
    m, k = A.shape
    _lib(dtype).oracle_optimizeA_implicit(_p(A), k, _p(B), k, m, B.shape[0], k, _p(ptr_), _p(idx), _p(val), lam, use_cg,
                                          max_cg_steps)


def fit_explicit(dtype, ixA, ixB, X, m, n, k, *, lam=0.05, user_bias=True, item_bias=True, center=True, scale_lam=False,
                 niter=3, use_cg=True, max_cg_steps=3, finalize_chol=False, seed=1, nthreads=4, w_main=1.0,
                 lam_unique=None, k_main=0, **_ignored):
    dt = np.dtype(dtype)
    kk = k + k_main
    A = np.zeros((m, kk), dt)
    B = np.zeros((n, kk), dt)
    bA = np.zeros(m, dt)
    bB = np.zeros(n, dt)
    g = np.zeros(1, dt)
    lu = None if lam_unique is None else np.asarray(lam_unique, dt)
    rc = _lib(dt).oracle_fit_explicit(_p(bA), _p(bB), _p(A), _p(B), seed, _p(g), m, n, kk,
                                      _p(np.ascontiguousarray(ixA, np.int32)), _p(np.ascontiguousarray(ixB, np.int32)),
                                      _p(np.ascontiguousarray(X, dt)), X.size, user_bias, item_bias, center, lam, _p(lu),
                                      scale_lam, w_main, niter, nthreads, use_cg, max_cg_steps, finalize_chol)
    return dict(rc=rc, A=A, B=B, biasA=bA, biasB=bB, glob_mean=g[0])


def fit_implicit(dtype, ixA, ixB, X, m, n, k, *, lam=5.0, alpha=1.0, niter=3, use_cg=True, max_cg_steps=3,
                 finalize_chol=False, seed=1, w_main=1.0, adjust_weight=False, apply_log_transf=False, k_main=0,
                 **_ignored):
    dt = np.dtype(dtype)
    kk = k + k_main
    A = np.zeros((m, kk), dt)
    B = np.zeros((n, kk), dt)
    wm = np.zeros(1, dt)
    rc = _lib(dt).oracle_fit_implicit(_p(A), _p(B), seed, m, n, kk, _p(np.ascontiguousarray(ixA, np.int32)),
                                      _p(np.ascontiguousarray(ixB, np.int32)), _p(np.ascontiguousarray(X, dt)), X.size,
                                      lam, w_main, _p(wm), alpha, adjust_weight, apply_log_transf, niter, use_cg,
                                      max_cg_steps, finalize_chol)
    return dict(rc=rc, A=A, B=B, w_main_multiplier=wm[0])
