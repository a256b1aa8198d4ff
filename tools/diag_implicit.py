#!/usr/bin/env python
"""Developer diagnostic: how the GPU / restatement / reference implicit fits drift apart per iteration (fp32)."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cmfrec_b200 import _lib
import refload
from support import fit_implicit, synth_coo
from oracle import restatement as O

def q(a, b):
    e = np.abs(a.astype(np.float64) - b).max(axis=1) / np.abs(b).max()
    return "q50 %.1e q90 %.1e q99 %.1e q99.9 %.1e max %.1e" % tuple(np.quantile(e, [0.5, 0.9, 0.99, 0.999, 1.0]))

dt = np.dtype(np.float32)
L, R = _lib.load(dt), refload.ref(dt)
for case, kw in ((7, dict(k_main=2)), (0, dict()), (6, dict(w_main=3.0))):
    m, n, k = 20000, 9000, 16
    ixA, ixB, X = synth_coo(m, n, 150000, dt, seed=30 + case, kind="counts")
    for niter in (1, 2, 3):
        b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, niter=niter, nthreads=4, **kw)
        b1 = fit_implicit(R, dt, ixA, ixB, X, m, n, k, niter=niter, nthreads=1, **kw)
        o = O.fit_implicit(dt, ixA, ixB, X, m, n, k, niter=niter, nthreads=4, **kw)
        print("case %d niter %d  ref(1 thread) vs ref: A %s | B %s" % (case, niter, q(b1["A"], b["A"]), q(b1["B"], b["B"])))
        print("case %d niter %d  restatement vs ref:   A %s | B %s" % (case, niter, q(o["A"], b["A"]), q(o["B"], b["B"])))
        for res in ("0", "1"):
            os.environ["CMFB200_RESIDENT"] = res
            a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, niter=niter, nthreads=4, **kw)
            print("case %d niter %d  gpu resident=%s vs ref: A %s | B %s" % (case, niter, res, q(a["A"], b["A"]), q(a["B"], b["B"])))
