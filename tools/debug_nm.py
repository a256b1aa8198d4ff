#!/usr/bin/env python
"""Developer aid: one explicit half-sweep through the tensor-core sweep (CG and Cholesky) against the reference, with the
error broken down by column / row so that a wrong block of the normal matrix shows up."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import AlsSession, csr_csc, ref, ref_optimizeA, synth_coo
from cmfrec_b200 import _lib

dt = np.dtype(np.float32); L = _lib.load(dt); R = ref(dt)
for k in (16, 40, 64):
    for solver in ("cg", "chol"):
        m, n = 600, 380
        ixA, ixB, X = synth_coo(m, n, 8000, dt, seed=100 + k)
        X = (X - X.mean()).astype(dt)
        csr = csr_csc(L, dt, ixA, ixB, X, m, n)
        rng = np.random.default_rng(k)
        A0 = (rng.normal(size=(m, k)) * 0.1).astype(dt); B0 = (rng.normal(size=(n, k)) * 0.1).astype(dt)
        bA0 = (rng.normal(size=m) * 0.3).astype(dt); bB0 = (rng.normal(size=n) * 0.3).astype(dt)
        with AlsSession(L, dt, csr[:3], csr[3:], m, n, k, implicit=False, user_bias=True, item_bias=True, lam_A=1.5, lam_B=1.5,
                        lam_biasA=2.5, lam_biasB=2.5) as s:
            s.set_factors(A0, bA0, B0, bB0)
            s.half_sweep(0, 1, 0 if solver == "cg" else 1)
            _, _, B1, bB1 = s.get_factors(with_bias=True)
        A_b = np.concatenate([A0, np.ones((m, 1), dt)], 1); B_b = np.concatenate([B0, np.ones((n, 1), dt)], 1)
        Xcsc = (csr[5] - bA0[csr[4]]).astype(dt)
        ref_optimizeA(R, dt, B_b, A_b, csr[3], csr[4], Xcsc, lam=1.5, lam_last=2.5, scale_lam=False, use_cg=solver == "cg", max_cg_steps=3)
        got = np.concatenate([B1, bB1[:, None]], 1)
        err = np.abs(got - B_b) / np.abs(B_b).max()
        deg = np.diff(csr[3]).astype(int)
        print("k=%d %s max err %.2e | by column block [0:32) %.2e [32:k) %.2e bias %.2e | rows deg<=32 %.2e deg>32 %.2e" % (
            k, solver, err.max(), err[:, :32].max(), err[:, 32:k].max() if k > 32 else 0, err[:, k].max(),
            err[deg <= 32].max(), err[deg > 32].max() if (deg > 32).any() else 0))
