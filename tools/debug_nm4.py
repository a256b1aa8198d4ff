#!/usr/bin/env python
"""Developer aid: the failing order of tests/test_gpu_sweeps.py replayed outside pytest, with the reference library loaded
and called in between as the test does."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import AlsSession, csr_csc, synth_coo, ref, ref_optimizeA
from cmfrec_b200 import _lib

dt = np.dtype(np.float32); L = _lib.load(dt); L64 = _lib.load(np.float64)
R = ref(dt) if os.environ.get("WITH_REF", "1") == "1" else None
def run(k, scale_lam):
    m, n = 600, 380
    ixA, ixB, X = synth_coo(m, n, 8000, dt, seed=100 + k)
    X = (X - X.mean()).astype(dt)
    csr = csr_csc(L, dt, ixA, ixB, X, m, n)
    rng = np.random.default_rng(k)
    A0 = (rng.normal(size=(m, k)) * 0.1).astype(dt); B0 = (rng.normal(size=(n, k)) * 0.1).astype(dt)
    lam, lb = (0.05, 0.11) if scale_lam else (1.5, 2.5)
    with AlsSession(L, dt, csr[:3], csr[3:], m, n, k, implicit=False, user_bias=False, item_bias=False, lam_A=lam, lam_B=lam,
                    lam_biasA=lb, lam_biasB=lb, scale_lam=scale_lam) as s:
        s.set_factors(A0, None, B0, None)
        s.half_sweep(0, 1, 1)
        _, _, B1, bB1 = s.get_factors(with_bias=True)
        s.half_sweep(1, 1, 1)
        A1, _, _, _ = s.get_factors(with_bias=True)
    msg = ""
    if R is not None:
        Bsol = B0.copy()
        ref_optimizeA(R, dt, Bsol, A0.copy(), csr[3], csr[4], csr[5], lam=lam, lam_last=lam, scale_lam=scale_lam, use_cg=False, max_cg_steps=3)
        msg = "err vs ref %.2e" % (np.abs(B1 - Bsol).max() / np.abs(Bsol).max())
    print("k=%d scale_lam=%s finite=%s %s" % (k, scale_lam, np.isfinite(B1).all() and np.isfinite(A1).all(), msg), flush=True)

for sl in (False, True):
    for k in (3, 16, 40, 64, 128):
        run(k, sl)
