#!/usr/bin/env python
"""Developer aid: Cholesky half-sweeps through the tensor-core sweep for several bias configurations / ranks; reports
which rows come back non-finite."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import AlsSession, csr_csc, synth_coo
from cmfrec_b200 import _lib

dt = np.dtype(np.float32); L = _lib.load(dt)
for k in (16, 24, 32, 33, 40, 64):
    for ub, ib in ((True, True), (False, False), (True, False), (False, True)):
        m, n = 600, 380
        ixA, ixB, X = synth_coo(m, n, 8000, dt, seed=100 + k)
        X = (X - X.mean()).astype(dt)
        csr = csr_csc(L, dt, ixA, ixB, X, m, n)
        rng = np.random.default_rng(k)
        A0 = (rng.normal(size=(m, k)) * 0.1).astype(dt); B0 = (rng.normal(size=(n, k)) * 0.1).astype(dt)
        bA0 = (rng.normal(size=m) * 0.3).astype(dt) if ub else None; bB0 = (rng.normal(size=n) * 0.3).astype(dt) if ib else None
        with AlsSession(L, dt, csr[:3], csr[3:], m, n, k, implicit=False, user_bias=ub, item_bias=ib, lam_A=1.5, lam_B=1.5,
                        lam_biasA=2.5, lam_biasB=2.5) as s:
            s.set_factors(A0, bA0, B0, bB0)
            assert L.cmfb200_debug_poison_smem(0x7fc00000) == 0
            s.half_sweep(0, 1, 1)
            _, _, B1, bB1 = s.get_factors(with_bias=True)
        bad = np.nonzero(~np.isfinite(B1).all(axis=1))[0]
        deg = np.diff(csr[3]).astype(int)
        print("k=%d biases=%s non-finite rows: %d of %d; their degrees %s; first bad cols %s" % (
            k, (ub, ib), bad.size, n, sorted(set(deg[bad].tolist()))[:12],
            np.nonzero(~np.isfinite(B1[bad[0]]))[0][:8].tolist() if bad.size else []))
