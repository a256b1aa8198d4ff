"""T2 parity: whole fits through the reference-named entry points (include/cmfrec_b200.h PART 1) against the
reference build with the same arguments and seed.  Tolerances: fp64 1e-7 relative after several iterations
(differences in summation order compound through the alternation), fp32 5e-3; the initial state (niter=0)
must be bit-identical."""
import numpy as np
import pytest

from support import fit_explicit, fit_implicit, ref, rel_err, rows_match, synth_coo

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-7, np.dtype(np.float32): 5e-3}


def _ref(dt):
    R = ref(dt)
    if R is None:
        pytest.fail("oracle/_ref is not built")
    return R


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_initial_state_bit_exact(gpu_libs, dtype):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n, k = 3000, 1500, 20
    ixA, ixB, X = synth_coo(m, n, 40000, dt, seed=3)
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=0, nthreads=4)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, niter=0, nthreads=4)
    assert a["rc"] == 0 and b["rc"] == 0
    for key in ("A", "B", "biasA", "biasB"):
        assert np.array_equal(a[key], b[key]), key
    assert a["glob_mean"] == b["glob_mean"]
    ixA, ixB, X = synth_coo(m, n, 40000, dt, seed=4, kind="counts")
    a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, niter=0)
    b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, niter=0)
    assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["B"], b["B"])


def _coo_with_long_rows(dt, seed):
    """Ratings with item rows of 9500 (several tiles of the block-per-row bias kernel, ragged last tile), 2000, 1024 and
    1023 entries (either side of its row-length threshold) and a user row of 2900 entries, on a sparse background."""
    rng = np.random.default_rng(seed)
    m, n = 12000, 3000
    pairs = set()
    for item, cnt in ((0, 9500), (1, 2000), (2, 1024), (3, 1023)):
        for u in rng.choice(m, size=cnt, replace=False):
            pairs.add((int(u), item))
    for i in 4 + rng.choice(n - 4, size=2900, replace=False):
        pairs.add((0, int(i)))
    bg_r = rng.integers(0, m, size=30000)
    bg_c = rng.integers(4, n, size=30000)
    pairs.update(zip(bg_r.tolist(), bg_c.tolist()))
    rc = np.array(sorted(pairs), dtype=np.int64)
    X = (rng.integers(1, 11, size=rc.shape[0]) * 0.5).astype(dt)
    return rc[:, 0].astype(np.int32), rc[:, 1].astype(np.int32), X, m, n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("biases", [(True, True), (True, False), (False, True)])
def test_initial_biases_bit_exact_with_long_rows(gpu_libs, dtype, biases):
    """The bias initialisation is a sequential chain per row (reference src/common.c:4643-4669, :4799-4825): rows long
    enough for the block-per-row path of bias_sweep_kernel must come out bit for bit like the short ones."""
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    ixA, ixB, X, m, n = _coo_with_long_rows(dt, 11)
    kw = dict(niter=0, nthreads=4, user_bias=biases[0], item_bias=biases[1], scale_lam=True, lam=0.05)
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, 8, **kw)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, 8, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    for key in ("biasA", "biasB"):
        assert np.array_equal(a[key], b[key]), key
    assert a["glob_mean"] == b["glob_mean"]


CASES_EXPLICIT = [
    dict(),                                                             # default: biases, centre, CG
    dict(finalize_chol=True),                                           # last iteration exact (CMF default)
    dict(use_cg=False),                                                 # Cholesky throughout
    dict(scale_lam=True, lam=0.05),                                     # benchmark hyper-parameters
    dict(user_bias=False, item_bias=False, center=False),
    dict(user_bias=True, item_bias=False),
    dict(user_bias=False, item_bias=True),
    dict(w_main=2.5, finalize_chol=True),
    dict(lam_unique=[0.3, 0.2, 0.07, 0.09, 1.0, 1.0], scale_lam=True),
    dict(k_main=3),
    dict(max_cg_steps=5, niter=2),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(CASES_EXPLICIT)))
def test_explicit_fit_matches_reference(gpu_libs, dtype, case):
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    kw = dict(lam=1.0, niter=3, nthreads=4)
    kw.update(CASES_EXPLICIT[case])
    m, n, k = 1200, 700, 16
    ixA, ixB, X = synth_coo(m, n, 30000, dt, seed=10 + case)
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    assert a["glob_mean"] == b["glob_mean"]
    tol = TOL[dt]
    assert rows_match(a["A"], b["A"], tol), rel_err(a["A"], b["A"])
    assert rows_match(a["B"], b["B"], tol), rel_err(a["B"], b["B"])
    s = max(np.abs(b["A"]).max(), 1e-30)
    if kw.get("user_bias", True):
        assert rows_match(a["biasA"][:, None], b["biasA"][:, None], tol * max(1.0, s / np.abs(b["biasA"]).max()))
    if kw.get("item_bias", True):
        assert rows_match(a["biasB"][:, None], b["biasB"][:, None], tol * max(1.0, s / np.abs(b["biasB"]).max()))


CASES_COLLECTIVE = [
    dict(side="UI", use_cg=False),                                 # config-3 style: Cholesky + dense U and I
    dict(side="UI", use_cg=False, w_user=1.7, w_item=0.6, scale_lam=True, lam=0.05),
    dict(side="UI"),                                               # CG with side information
    dict(side="U", finalize_chol=True),
    dict(side="I", user_bias=False, item_bias=False, center=False),
    dict(side="", add_implicit_features=True),                     # config-4 style: CG + implicit features
    dict(side="", add_implicit_features=True, use_cg=False, w_implicit=0.5),
    dict(side="", add_implicit_features=True, scale_lam=True, lam=0.05, w_implicit=0.5),
    dict(side="UI", add_implicit_features=True, w_main=2.0, finalize_chol=True),
    dict(side="U", center_side=False, lam_unique=[0.3, 0.2, 0.07, 0.09, 0.5, 0.6]),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(CASES_COLLECTIVE)))
def test_collective_fit_matches_reference(gpu_libs, dtype, case):
    """side information (dense U / I) and implicit features: C, D, Ai, Bi, A, B against the reference"""
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    kw = dict(lam=1.0, niter=3, nthreads=4)
    kw.update(CASES_COLLECTIVE[case])
    side = kw.pop("side")
    m, n, k = 900, 500, 12
    ixA, ixB, X = synth_coo(m, n, 25000, dt, seed=50 + case)
    rng = np.random.default_rng(case)
    if "U" in side:
        kw["U"] = rng.normal(size=(m, 7)).astype(dt) + 0.3
    if "I" in side:
        kw["I"] = rng.normal(size=(n, 5)).astype(dt) - 0.2
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    tol = TOL[dt] * (1 if dt == np.float32 else 10)
    for key in ("U_colmeans", "I_colmeans"):
        if b[key] is not None:
            assert np.array_equal(a[key], b[key]), key
    for key in ("A", "B", "C", "D", "Ai", "Bi"):
        if b[key] is not None:
            assert rows_match(a[key], b[key], tol, 0.01), (key, rel_err(a[key], b[key]))


CASES_IMPLICIT = [
    dict(),
    dict(alpha=40.0, lam=1.0),
    dict(use_cg=False),
    dict(finalize_chol=True),
    dict(apply_log_transf=True, alpha=2.0),
    dict(adjust_weight=True, lam=1e-3),
    dict(w_main=3.0),
    dict(k_main=2),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(CASES_IMPLICIT)))
def test_implicit_fit_matches_reference(gpu_libs, dtype, case):
    """The truncated CG on heavy-tailed counts amplifies summation-order noise: a row sitting on one of the
    absolute exit thresholds takes one step more or fewer and the difference spreads through the alternation
    (the CPU restatement in oracle/ differs from the reference by up to 1e-1 in fp32 / 2e-5 in fp64 on these
    inputs, and which single row is worst changes with the seed: tools/diag_implicit2.py).  So the GPU result
    is required (a) to be no further from the reference than 3x what that independent CPU restatement is
    (+1e-7 fp64 / 5e-3 fp32) at the median, the 90th and the 99th percentile of the per-row error -- the single
    worst row is decided by one step-count flip on one popular item and is only held to an absolute 0.25 --
    and (b) to reach the same objective value to 1e-3 relative."""
    from oracle import restatement as O
    from support import implicit_objective
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    kw = dict(niter=3, nthreads=4)
    kw.update(CASES_IMPLICIT[case])
    m, n, k = 20000, 9000, 16          # > 2^18 factor entries, so the uniform initialiser is taken (Q7)
    ixA, ixB, X = synth_coo(m, n, 150000, dt, seed=30 + case, kind="counts")
    a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    assert a["w_main_multiplier"] == b["w_main_multiplier"]
    o = O.fit_implicit(dt, ixA, ixB, X, m, n, k, **kw)
    for key in ("A", "B"):
        scale = np.abs(b[key]).max()
        row_err = lambda x: np.abs(x.astype(np.float64) - b[key]).max(axis=1) / scale
        ours, noise = np.quantile(row_err(a[key]), [0.5, 0.9, 0.99]), np.quantile(row_err(o[key]), [0.5, 0.9, 0.99])
        assert (ours <= 3 * noise + TOL[dt]).all(), (key, ours, noise)
        assert rel_err(a[key], b[key]) <= max(0.25 if dt == np.float32 else 1e-4, 3 * rel_err(o[key], b[key])), key
    Xt = np.log(X) if kw.get("apply_log_transf") else X
    lam = kw.get("lam", 5.0)
    fa = implicit_objective(ixA, ixB, Xt, a["A"], a["B"], lam, kw.get("alpha", 1.0))
    fb = implicit_objective(ixA, ixB, Xt, b["A"], b["B"], lam, kw.get("alpha", 1.0))
    assert abs(fa - fb) <= 1e-3 * abs(fb), (fa, fb)


CASES_IMPLICIT_SIDE = [
    dict(side="UI"),                                      # CG, both matrices
    dict(side="UI", use_cg=False),                        # Cholesky throughout
    dict(side="U", finalize_chol=True, w_user=2.5),
    dict(side="I", w_item=0.4, alpha=10.0, lam=1.0),
    dict(side="UI", w_main=2.0, lam_unique=[0, 0, 3.0, 4.0, 1.5, 2.5], center_side=False),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(CASES_IMPLICIT_SIDE)))
@pytest.mark.parametrize("niter", [1, 3])
def test_implicit_side_information_matches_reference(gpu_libs, dtype, case, niter):
    """Implicit feedback WITH dense side information: fit_collective_implicit_als with U / I against the
    reference (optimizeA_collective_implicit src/collective.c:5971, collective_block_cg_implicit :2905,
    collective_closed_form_block_implicit :1849).  niter = 1 is exactly one call of each per side from the
    bit-identical starting point (the per-half-sweep comparison); zero-mean-ish side information keeps the systems
    well conditioned, so float32 is held to 1e-3 of max|F| on 99 % of the rows."""
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    kw = dict(niter=niter, nthreads=4)
    kw.update(CASES_IMPLICIT_SIDE[case])
    side = kw.pop("side")
    m, n, k = 20000, 9000, 16          # > 2^18 factor entries: the uniform initialiser (Q7)
    ixA, ixB, X = synth_coo(m, n, 150000, dt, seed=60 + case, kind="counts")
    rng = np.random.default_rng(case)
    if "U" in side:
        kw["U"] = rng.normal(size=(m, 6)).astype(dt) + 0.3
    if "I" in side:
        kw["I"] = rng.normal(size=(n, 5)).astype(dt) - 0.2
    a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    for key in ("U_colmeans", "I_colmeans"):
        if b[key] is not None:
            assert np.array_equal(a[key], b[key]), key
    if dt == np.float64:
        for key in ("C", "D", "B", "A"):
            if b[key] is not None:
                assert rows_match(a[key], b[key], 1e-7, 0.01), (key, rel_err(a[key], b[key]))
        return
    # float32: the fit starts from all-positive uniform factors (ill conditioned, and the truncated CG amplifies it through
    # the alternation), so the GPU is held to the reference's OWN float32 accuracy: row errors against the float64
    # reference fit no larger than 3x those of the float32 reference fit, at the median / 90th / 99th percentile
    R64 = _ref(np.float64)
    kw64 = {key: (np.asarray(v, np.float64) if isinstance(v, np.ndarray) else v) for key, v in kw.items()}
    e = fit_implicit(R64, np.float64, ixA, ixB, X.astype(np.float64), m, n, k, **kw64)
    for key in ("C", "D", "B", "A"):
        if b[key] is None:
            continue
        scale = np.abs(e[key]).max()
        e_gpu = np.abs(a[key].astype(np.float64) - e[key]).max(axis=1) / scale
        e_ref = np.abs(b[key].astype(np.float64) - e[key]).max(axis=1) / scale
        qs = [0.5, 0.9, 0.99] if a[key].shape[0] > 100 else [0.5, 1.0]
        assert (np.quantile(e_gpu, qs) <= 3 * np.quantile(e_ref, qs) + 1e-4).all(), (key, np.quantile(e_gpu, qs), np.quantile(e_ref, qs))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_precomputed_outputs(gpu_libs, dtype):
    """precompute_for_predictions: B_plus_bias, BtB (upper triangle), TransBtBinvBt (src/collective.c:8935-9075)"""
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n, k = 900, 500, 12
    ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=77)
    kw = dict(lam=0.8, niter=2, precompute=True, finalize_chol=True)
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    tol = TOL[dt] * 10
    assert rel_err(a["B_plus_bias"], b["B_plus_bias"]) <= tol
    iu = np.triu_indices(k + 1)
    assert rel_err(a["BtB"][iu], b["BtB"][iu]) <= tol
    assert rel_err(a["TransBtBinvBt"], b["TransBtBinvBt"]) <= tol
    ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=78, kind="counts")
    a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, niter=2, precompute=True)
    b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, niter=2, precompute=True)
    assert rel_err(a["BtB"][np.triu_indices(k)], b["BtB"][np.triu_indices(k)]) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", ["UI", "U_implicit_features", "implicit_feedback_U"])
def test_precomputed_outputs_collective(gpu_libs, dtype, case):
    """precompute_for_predictions (the Python default) with side information / implicit features: BtB, BiTBi, CtCw,
    TransCtCinvCt, TransBtBinvBt, BeTBeChol (src/collective.c:8935-9255); implicit feedback with U: BtB, BeTBe, BeTBeChol
    (:10055-10105).  Upper triangles are what the reference defines."""
    dt = np.dtype(dtype)
    L, R = gpu_libs[dt], _ref(dt)
    m, n, k = 900, 500, 10
    rng = np.random.default_rng(3)
    U = rng.normal(size=(m, 6)).astype(dt); I = rng.normal(size=(n, 4)).astype(dt)
    tol = 1e-6 if dt == np.float64 else 5e-3
    iu = np.triu_indices(k + 1)
    ik = np.triu_indices(k)
    if case == "implicit_feedback_U":
        ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=81, kind="counts")
        kw = dict(niter=2, precompute=True, U=U, w_user=1.5)
        a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, **kw)
        b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, **kw)
        assert a["rc"] == 0 and b["rc"] == 0
        for key in ("BtB", "BeTBe", "BeTBeChol"):
            assert rel_err(a[key][ik], b[key][ik]) <= tol, key
        return
    ixA, ixB, X = synth_coo(m, n, 20000, dt, seed=80)
    kw = dict(lam=0.8, niter=2, precompute=True, finalize_chol=True, U=U, w_user=1.5)
    if case == "UI":
        kw["I"] = I
    else:
        kw.update(add_implicit_features=True, w_implicit=0.5)
    a = fit_explicit(L, dt, ixA, ixB, X, m, n, k, **kw)
    b = fit_explicit(R, dt, ixA, ixB, X, m, n, k, **kw)
    assert a["rc"] == 0 and b["rc"] == 0
    assert rel_err(a["BtB"][iu], b["BtB"][iu]) <= tol
    assert rel_err(a["CtCw"][ik], b["CtCw"][ik]) <= tol
    assert rel_err(a["BeTBeChol"][iu], b["BeTBeChol"][iu]) <= tol
    if case == "UI":
        assert rel_err(a["TransBtBinvBt"], b["TransBtBinvBt"]) <= tol
        assert rel_err(a["TransCtCinvCt"], b["TransCtCinvCt"]) <= tol
    else:
        assert rel_err(a["BiTBi"][ik], b["BiTBi"][ik]) <= tol


def test_interrupt_returns_code_3(gpu_libs):
    """SIGINT during the alternation: the fit stops between half-sweeps and returns 3 (reference src/collective.c:8890,
    src/helpers.c:1493); the previous handler is put back afterwards."""
    import os, signal, threading, time
    dt = np.dtype(np.float32)
    L = gpu_libs[dt]
    m, n, k = 20000, 8000, 32
    ixA, ixB, X = synth_coo(m, n, 600000, dt, seed=9)
    assert fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=1)["rc"] == 0       # warm the pools
    before = signal.getsignal(signal.SIGINT)
    threading.Timer(0.5, lambda: os.kill(os.getpid(), signal.SIGINT)).start()
    t0 = time.time()
    try:
        out = fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=1000000)
    except KeyboardInterrupt:                      # raised by Python's own handler if the signal landed outside the call
        pytest.fail("the signal was not handled inside the fit")
    assert out["rc"] == 3 and time.time() - t0 < 30
    assert np.isfinite(out["A"]).all()
    assert signal.getsignal(signal.SIGINT) is before


def test_verbose_prints_the_references_progress_lines(gpu_libs, capfd):
    dt = np.dtype(np.float64)
    L = gpu_libs[dt]
    from support import ptr
    m, n, k = 400, 300, 6
    ixA, ixB, X = synth_coo(m, n, 5000, dt, seed=2)
    A = np.zeros((m, k)); B = np.zeros((n, k)); bA = np.zeros(m); bB = np.zeros(n); g = np.zeros(1); s1 = np.zeros(1); s2 = np.zeros(1)
    rc = L.fit_collective_explicit_als(
        ptr(bA), ptr(bB), ptr(A), ptr(B), None, None, None, None, False, True, 1, ptr(g), None, None, m, n, k,
        ptr(ixA), ptr(ixB), ptr(X), X.size, None, None, True, True, True, 1.0, None, 0.0, None, False, False, False,
        ptr(s1), ptr(s2), None, 0, 0, None, 0, 0, None, None, None, 0, None, None, None, 0, False, False, False,
        0, 0, 0, 1.0, 1.0, 1.0, 1.0, 2, 1, True, False, True, 3, False, False, False, 100, False, False, False, True,
        None, None, None, None, None, None, None, None, None)
    assert rc == 0
    text = capfd.readouterr().out
    assert "Starting ALS optimization routine" in text and text.count("Updating B ... done") == 2
    assert "Completed ALS iteration  2" in text and "ALS procedure terminated successfully" in text


def test_unsupported_arguments_are_refused(gpu_libs):
    """No silent fallback: what the GPU path does not cover returns code 2."""
    dt = np.dtype(np.float64)
    L = gpu_libs[dt]
    m, n, k = 50, 40, 4
    ixA, ixB, X = synth_coo(m, n, 300, dt, seed=1)
    from support import ptr
    A = np.zeros((m, k)); B = np.zeros((n, k)); g = np.zeros(1)
    w = np.ones(X.size)
    U = np.zeros((m + 3, 2))
    assert fit_explicit(L, dt, ixA, ixB, X, m, n, k, U=np.full((m, 2), np.nan))["rc"] == 2     # missing values in U
    rc = L.fit_collective_explicit_als(
        None, None, ptr(A), ptr(B), None, None, None, None, False, True, 1, ptr(g), None, None, m, n, k,
        ptr(ixA), ptr(ixB), ptr(X), X.size, None, ptr(w), False, False, True, 1.0, None, 0.0, None, False, False, False,
        None, None, None, 0, 0, None, 0, 0, None, None, None, 0, None, None, None, 0, False, False, False,
        0, 0, 0, 1.0, 1.0, 1.0, 1.0, 2, 1, False, False, True, 3, False, False, False, 100, False, False, False, True,
        None, None, None, None, None, None, None, None, None)
    assert rc == 2


def test_pooled_buffers_do_not_leak_state_between_calls(gpu_libs):
    """Device buffers are recycled between fit calls through the device memory pool (als.h: DevBuf): a second and a
    third call -- the last one after cmfb200_trim_pool() handed the cache back -- must reproduce the first bit for bit,
    for both feedback models, and so must a call of a different shape in between."""
    dt = np.dtype(np.float32)
    L = gpu_libs[dt]
    m, n, k = 3000, 2000, 24
    ixA, ixB, X = synth_coo(m, n, 60000, dt, seed=5)
    first = fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=2)
    other = synth_coo(500, 700, 9000, dt, seed=6, kind="counts")
    assert fit_implicit(L, dt, *other, 500, 700, 8, niter=1)["rc"] == 0
    second = fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=2)
    L.cmfb200_trim_pool()
    third = fit_explicit(L, dt, ixA, ixB, X, m, n, k, niter=2, copy_inputs=False)
    for key in ("A", "B", "biasA", "biasB"):
        assert np.array_equal(first[key], second[key]), key
        assert np.array_equal(first[key], third[key]), key
