"""The oracle (oracle/cmf_oracle.c, this repo's plain-C restatement) pinned against the golden vectors produced
by the reference, and against the reference build where it exists.  Runs without a GPU.

Tolerances: integers and the seeded initial state bit-exact; one half-sweep fp64 1e-10 / fp32 2e-4 relative
(different summation order than OpenBLAS); whole fits fp64 1e-7 / fp32 5e-3 on >= 99 % of the rows."""
import os

import numpy as np
import pytest

from oracle import restatement as O
from support import fit_explicit, fit_implicit, ref, rel_err, rows_match, synth_coo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = [np.float32, np.float64]
SWEEP_TOL = {np.dtype(np.float64): 1e-10, np.dtype(np.float32): 2e-4}
FIT_TOL = {np.dtype(np.float64): 1e-7, np.dtype(np.float32): 5e-3}


def tag(dt):
    return "f32" if np.dtype(dt) == np.float32 else "f64"


@pytest.mark.parametrize("dtype", DTYPES)
def test_init_bit_exact(dtype):
    dt = np.dtype(dtype)
    g = np.load(os.path.join(GOLD, "init_%s.npz" % tag(dt)))
    for name in ("small_normal", "small_uniform_req", "big_normal", "big_uniform"):
        sa, sb, normal, seed = [int(v) for v in g[name + "_args"]]
        A, B = O.random_init(dt, sa, sb, seed, bool(normal))
        assert np.array_equal(A[g[name + "_selA"]], g[name + "_A"]), name
        assert np.array_equal(B[:256], g[name + "_B"]), name


@pytest.mark.parametrize("dtype", DTYPES)
def test_prep_bit_exact(dtype):
    dt = np.dtype(dtype)
    g = np.load(os.path.join(GOLD, "prep_%s.npz" % tag(dt)))
    m, n = int(g["m"]), int(g["n"])
    p, i, v = O.coo_to_csr(dt, g["ixA"], g["ixB"], g["X"], m)
    assert np.array_equal(p, g["csr_p"]) and np.array_equal(i, g["csr_i"]) and np.array_equal(v, g["csr_v"])
    cp, ci, cv = O.coo_to_csr(dt, g["ixB"], g["ixA"], g["X"], n)
    assert np.array_equal(cp, g["csc_p"]) and np.array_equal(ci, g["csc_i"]) and np.array_equal(cv, g["csc_v"])
    assert O.global_mean(dt, g["X"], 1) == g["mean_nt1"][0]
    xc = (g["X"] - g["mean_nt1"][0]).astype(dt)
    csr = O.coo_to_csr(dt, g["ixA"], g["ixB"], xc, m)
    csc = O.coo_to_csr(dt, g["ixB"], g["ixA"], xc, n)
    for scale in (0, 1):
        bA, bB = O.init_biases_twosided(dt, m, n, csr, csc, 0.05, 0.07, bool(scale))
        assert np.array_equal(bA, g["biasA_scale%d" % scale]) and np.array_equal(bB, g["biasB_scale%d" % scale])


@pytest.mark.parametrize("dtype", DTYPES)
def test_half_sweeps_match_golden(dtype):
    dt = np.dtype(dtype)
    g = np.load(os.path.join(GOLD, "sweep_%s.npz" % tag(dt)))
    gp = np.load(os.path.join(GOLD, "prep_%s.npz" % tag(dt)))
    m = int(gp["m"])
    xc = (gp["X"] - gp["mean_nt1"][0]).astype(dt)
    csr_c = O.coo_to_csr(dt, gp["ixA"], gp["ixB"], xc, m)
    csr = (gp["csr_p"], gp["csr_i"], gp["csr_v"])
    A0, B0 = g["A0"], g["B0"]
    tol = SWEEP_TOL[dt]
    for solver in ("cg", "chol"):
        for scale in (0, 1):
            A1 = A0.copy()
            O.optimizeA(dt, A1, B0.copy(), *csr_c, lam=0.7 if not scale else 0.05, lam_last=1.3 if not scale else 0.09,
                        scale_lam=bool(scale), use_cg=solver == "cg", max_cg_steps=3)
            assert rows_match(A1, g["explicit_%s_scale%d" % (solver, scale)], tol), (solver, scale)
        A1 = np.abs(A0).copy()
        O.optimizeA_implicit(dt, A1, np.abs(B0).copy(), *csr, lam=2.0, use_cg=solver == "cg", max_cg_steps=3)
        assert rows_match(A1, g["implicit_%s" % solver], tol), solver


@pytest.mark.parametrize("dtype", DTYPES)
def test_fits_match_golden(dtype):
    dt = np.dtype(dtype)
    g = np.load(os.path.join(GOLD, "fit_%s.npz" % tag(dt)))
    tol = FIT_TOL[dt]
    o = O.fit_explicit(dt, g["e_ixA"], g["e_ixB"], g["e_X"], 800, 450, 10, lam=0.8, niter=2, finalize_chol=True, nthreads=2)
    assert o["glob_mean"] == g["e_glob_mean"]
    for key in ("A", "B"):
        assert rows_match(o[key], g["e_" + key], tol, 0.01), (key, rel_err(o[key], g["e_" + key]))
    for key in ("biasA", "biasB"):
        assert rows_match(o[key][:, None], g["e_" + key][:, None], 10 * tol, 0.01), key
    m3, n3, nnz3, seed3 = [int(v) for v in g["i_args"]]
    ia, ib, x = synth_coo(m3, n3, nnz3, dt, seed=seed3, kind="counts")
    x = np.minimum(x, 20).astype(dt)
    o = O.fit_implicit(dt, ia, ib, x, m3, n3, 16, lam=4.0, niter=2, use_cg=False)
    assert rows_match(o["A"][g["i_selA"]], g["i_A"], tol, 0.01), rel_err(o["A"][g["i_selA"]], g["i_A"])
    assert rows_match(o["B"][::47], g["i_B"], tol, 0.01)


@pytest.mark.parametrize("dtype", DTYPES)
def test_oracle_against_reference_build(dtype):
    """more cases than the fixtures hold, where the reference build is available"""
    dt = np.dtype(dtype)
    R = ref(dt)
    if R is None:
        pytest.skip("oracle/_ref not built on this machine")
    tol = FIT_TOL[dt]
    m, n, k = 700, 400, 8
    ixA, ixB, X = synth_coo(m, n, 12000, dt, seed=2)
    for kw in (dict(), dict(use_cg=False), dict(scale_lam=True, lam=0.05), dict(user_bias=False), dict(item_bias=False),
               dict(user_bias=False, item_bias=False, center=False), dict(w_main=2.0, finalize_chol=True),
               dict(lam_unique=[0.3, 0.2, 0.07, 0.09, 1.0, 1.0])):
        args = dict(lam=1.0, niter=3, nthreads=2); args.update(kw)
        a = O.fit_explicit(dt, ixA, ixB, X, m, n, k, **args)
        b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, **args)
        assert a["glob_mean"] == b["glob_mean"]
        for key in ("A", "B"):
            assert rows_match(a[key], b[key], tol, 0.01), (kw, key, rel_err(a[key], b[key]))
    ixA, ixB, X = synth_coo(20000, 9000, 100000, dt, seed=3, kind="counts")
    X = np.minimum(X, 20).astype(dt)
    for kw in (dict(use_cg=False), dict(use_cg=False, alpha=3.0, adjust_weight=True, lam=0.01)):
        a = O.fit_implicit(dt, ixA, ixB, X, 20000, 9000, 8, niter=2, **kw)
        b = fit_implicit(R, dt, ixA, ixB, X, 20000, 9000, 8, niter=2, **kw)
        assert a["w_main_multiplier"] == b["w_main_multiplier"]
        assert rows_match(a["A"], b["A"], tol, 0.01), (kw, rel_err(a["A"], b["A"]))
