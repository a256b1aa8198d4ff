#!/usr/bin/env python
"""Developer aid: a k = 128 Cholesky half-sweep followed by k = 40 / 64 ones (the order in which the test-suite failed)."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import AlsSession, csr_csc, synth_coo
from cmfrec_b200 import _lib

dt = np.dtype(np.float32); L = _lib.load(dt)
def run(k, ub, ib, scale_lam, which=0):
    m, n = 600, 380
    ixA, ixB, X = synth_coo(m, n, 8000, dt, seed=100 + k)
    X = (X - X.mean()).astype(dt)
    csr = csr_csc(L, dt, ixA, ixB, X, m, n)
    rng = np.random.default_rng(k)
    A0 = (rng.normal(size=(m, k)) * 0.1).astype(dt); B0 = (rng.normal(size=(n, k)) * 0.1).astype(dt)
    bA0 = (rng.normal(size=m) * 0.3).astype(dt) if ub else None; bB0 = (rng.normal(size=n) * 0.3).astype(dt) if ib else None
    lam, lb = (0.05, 0.11) if scale_lam else (1.5, 2.5)
    with AlsSession(L, dt, csr[:3], csr[3:], m, n, k, implicit=False, user_bias=ub, item_bias=ib, lam_A=lam, lam_B=lam,
                    lam_biasA=lb, lam_biasB=lb, scale_lam=scale_lam) as s:
        s.set_factors(A0, bA0, B0, bB0)
        s.half_sweep(0, 1, 1)
        _, _, B1, bB1 = s.get_factors(with_bias=True)
        s.half_sweep(1, 1, 1)
        A1, _, _, _ = s.get_factors(with_bias=True)
    for name, F, ptr in (("B", B1, csr[3]), ("A", A1, csr[0])):
        bad = np.nonzero(~np.isfinite(F).all(axis=1))[0]
        deg = np.diff(ptr).astype(int)
        print("k=%d biases=%s scale_lam=%s %s: non-finite rows %d of %d; degrees %s; rows %s; cols of first %s" % (
            k, (ub, ib), scale_lam, name, bad.size, F.shape[0], sorted(set(deg[bad].tolist()))[:10], bad[:10].tolist(),
            np.nonzero(~np.isfinite(F[bad[0]]))[0][:6].tolist() if bad.size else []), flush=True)

for seq in ([(128, False), (3, True), (16, True), (40, True), (64, True)], [(40, True), (128, True), (40, True), (64, False)]):
    print("--- sequence", seq)
    for k, sl in seq:
        run(k, False, False, sl)
