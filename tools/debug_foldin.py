import sys, numpy as np
sys.path.insert(0, "tests")
from support import ptr, ref, synth_coo
from cmfrec_b200 import _lib
from test_gpu_foldin import call_explicit
dt = np.dtype(np.float64)
L = _lib.load(dt); R = ref(dt)
m, n, k = 50, 40, 8
rng = np.random.default_rng(0)
ixA, ixB, X = synth_coo(m, n, 600, dt, seed=1)
B = rng.normal(size=(n, k)).astype(dt)
for ub in (False, True):
    o = call_explicit(L, dt, m, n, k, ixA, ixB, X, B, None, 3.4, user_bias=ub)
    r = call_explicit(R, dt, m, n, k, ixA, ixB, X, B, None, 3.4, user_bias=ub)
    # numpy closed form for row 0
    for lam_try in (0.7,):
        sel = ixA == 0
        G = B[ixB[sel]]; x = X[sel] - 3.4
        if ub:
            G = np.hstack([G, np.ones((G.shape[0], 1))])
        M = G.T @ G + lam_try * np.eye(G.shape[1])
        a = np.linalg.solve(M, G.T @ x)
        print("ub", ub, "numpy", a[:4], "ours", o[1][0][:4], "ref", r[1][0][:4])
