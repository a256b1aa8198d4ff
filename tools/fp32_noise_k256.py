#!/usr/bin/env python
"""How far apart two CPU float32 implementations of the same implicit CG half-sweep are at k = 256 (the reference
build vs this repo's C restatement vs exact float64 arithmetic), for all-positive uniform factors and for zero-mean
factors: the calibration behind the tolerances of tests/test_gpu_bench_shapes.py::test_large_rank_implicit_fp32.
CPU only; needs oracle/_ref."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import csr_csc, ref, ref_optimizeA_implicit, synth_coo
from oracle import restatement as O
from cmfrec_b200 import _lib

dt = np.dtype(np.float32); L = _lib.load(dt); R = ref(dt)
k, m, n = 256, 20000, 9000
ixA, ixB, X = synth_coo(m, n, 400000, dt, seed=k, kind="counts")
csr = csr_csc(L, dt, ixA, ixB, X, m, n)
for init in ("uniform", "normal"):
    rng = np.random.default_rng(k)
    draw = (lambda s: rng.random(s)) if init == "uniform" else (lambda s: rng.normal(size=s))
    A0 = (draw((m, k)) * 0.1).astype(dt); B0 = (draw((n, k)) * 0.1).astype(dt)
    Aref = A0.copy(); ref_optimizeA_implicit(R, dt, Aref, B0.copy(), *csr[:3], lam=5.0, use_cg=True, max_cg_steps=3, nthreads=8)
    Ao = A0.copy(); O.optimizeA_implicit(dt, Ao, B0.copy(), *csr[:3], lam=5.0, use_cg=True, max_cg_steps=3)
    T = A0.astype(np.float64); O.optimizeA_implicit(np.float64, T, B0.astype(np.float64), csr[0], csr[1], csr[2].astype(np.float64), lam=5.0, use_cg=True, max_cg_steps=3)
    s = np.abs(T).max()
    for name, e in (("reference - exact", np.abs(Aref - T).max(axis=1) / s), ("restatement - exact", np.abs(Ao - T).max(axis=1) / s),
                    ("restatement - reference", np.abs(Ao.astype(np.float64) - Aref).max(axis=1) / s)):
        print(init, name, "q50/q90/q99/max", np.quantile(e, [0.5, 0.9, 0.99, 1.0]))
