#!/usr/bin/env python
"""Developer diagnostic: spread over seeds of the fp32 implicit fit's distance to the reference (direct vs resident)."""
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
    return "q50 %.1e q90 %.1e max %.1e" % tuple(np.quantile(e, [0.5, 0.9, 1.0]))

dt = np.dtype(np.float32)
L, R = _lib.load(dt), refload.ref(dt)
m, n = 20000, 9000
for k, kw in ((16, dict(k_main=2)), (18, dict()), (24, dict()), (32, dict())):
    for seed in (37, 101, 102, 103):
        ixA, ixB, X = synth_coo(m, n, 150000, dt, seed=seed, kind="counts")
        b = fit_implicit(R, dt, ixA, ixB, X, m, n, k, niter=2, nthreads=4, **kw)
        o = O.fit_implicit(dt, ixA, ixB, X, m, n, k, niter=2, nthreads=4, **kw)
        line = "k=%d %s seed %d | restatement A %s B %s" % (k, kw, seed, q(o["A"], b["A"]), q(o["B"], b["B"]))
        for res in ("0", "1"):
            os.environ["CMFB200_RESIDENT"] = res
            a = fit_implicit(L, dt, ixA, ixB, X, m, n, k, niter=2, nthreads=4, **kw)
            line += " | gpu res=%s A %s B %s" % (res, q(a["A"], b["A"]), q(a["B"], b["B"]))
        print(line, flush=True)
