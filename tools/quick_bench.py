#!/usr/bin/env python
"""Developer timing loop (not the graded bench): device time per ALS iteration on a synthetic shape."""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from cmfrec_b200 import _lib, synth
from support import AlsSession, csr_csc

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="ml10m")
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--implicit", type=int, default=0)
ap.add_argument("--solver", default="cg")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
dt = np.dtype(np.float32 if a.dtype == "f32" else np.float64)
L = _lib.load(dt)
t = time.time()
cache = "/tmp/cmfb200_qb_%s_%s_%g.npz" % (a.shape, a.dtype, a.scale)
if os.path.exists(cache):
    z = np.load(cache); ixA, ixB, X, m, n = z["a"], z["b"], z["x"], int(z["m"]), int(z["n"])
else:
    ixA, ixB, X, m, n = synth.make(a.shape, dt, a.scale)
    np.savez(cache, a=ixA, b=ixB, x=X, m=m, n=n)
print("generated %s: m=%d n=%d nnz=%d in %.1fs" % (a.shape, m, n, X.size, time.time() - t), flush=True)
if not a.implicit:
    X = (X - X.mean()).astype(dt)
t = time.time()
csr = csr_csc(L, dt, ixA, ixB, X, m, n)
print("csr+csc %.2fs" % (time.time() - t), flush=True)
rng = np.random.default_rng(0)
A0 = (rng.normal(size=(m, a.k)) * 0.01).astype(dt) if not a.implicit else (rng.random((m, a.k)) * 0.01).astype(dt)
B0 = np.zeros((n, a.k), dt)
bA = np.zeros(m, dt); bB = np.zeros(n, dt)
kw = dict(implicit=bool(a.implicit), lam_A=5.0 if a.implicit else 0.05, lam_B=5.0 if a.implicit else 0.05)
if not a.implicit:
    kw.update(user_bias=True, item_bias=True, scale_lam=True)
t = time.time()
with AlsSession(L, dt, csr[:3], csr[3:], m, n, a.k, **kw) as s:
    s.set_factors(A0, bA, B0, bB)
    print("setup+upload %.2fs" % (time.time() - t), flush=True)
    ms = C.c_float(0)
    use_cg = int(a.solver == "cg")
    rc = L.cmfb200_als_timed_iterate(s.h, 0, 2, 1000, use_cg, 0, C.byref(ms))
    assert rc == 0, rc
    print("warmup 2 iters: %.3f ms/iter" % (ms.value / 2), flush=True)
    rc = L.cmfb200_als_timed_iterate(s.h, 2, a.iters, 1000, use_cg, 0, C.byref(ms))
    assert rc == 0, rc
    per = ms.value / a.iters
    k1 = a.k + (0 if a.implicit else 1)
    w = dt.itemsize
    nnz = X.size
    bytes_iter = 2 * nnz * (k1 * w + 4 + w) + (m + n) * (8 + 2 * k1 * w) + ((m + n) * a.k * w if a.implicit else 0)
    print("RESULT shape=%s k=%d %s implicit=%d solver=%s: %.3f ms/iter  %.2f Mrows/s  alg %.2f GB/iter -> %.0f GB/s"
          % (a.shape, a.k, a.dtype, a.implicit, a.solver, per, (m + n) / per / 1e3, bytes_iter / 1e9, bytes_iter / per / 1e6), flush=True)
    A1, B1 = s.get_factors()
    print("finite:", np.isfinite(A1).all(), np.isfinite(B1).all(), "absmax", np.abs(A1).max(), np.abs(B1).max())
