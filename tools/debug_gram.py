#!/usr/bin/env python
"""Developer diagnostic for the tensor-core Gram: structured inputs whose products show the operand layout."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cmfrec_b200 import _lib
from support import ptr
dt = np.dtype(np.float32)
L = _lib.load(dt)
def gram(G):
    out = np.full((G.shape[1], G.shape[1]), -7.0, dt); ms = C.c_float(0)
    rc = L.cmfb200_gram(ptr(G), G.shape[0], G.shape[1], ptr(out), 0, C.byref(ms)); assert rc == 0, rc
    return out
np.set_printoptions(linewidth=250, precision=3, suppress=True)
for kk in (64, 128):
    G = np.ones((8, kk), dt); g = gram(G); print("ones 8 x", kk, "unique", np.unique(g)[:10], "nnz", np.count_nonzero(g))
    G = np.zeros((8, kk), dt); G[np.arange(8), np.arange(8)] = 1; g = gram(G); print("eye: nonzero at", np.argwhere(g != 0)[:20].tolist(), g[g != 0][:20])
    G = np.zeros((1, kk), dt); G[0, :] = np.arange(1, kk + 1); g = gram(G); w = G.T @ G
    print("arange row: got[0,:8]", g[0, :8], "want", w[0, :8]); print("got[:8,0]", g[:8, 0], "got[1,:8]", g[1, :8], "err", np.abs(g - w).max())
    G = np.zeros((16, kk), dt); G[9, :] = np.arange(1, kk + 1); g = gram(G); print("row 9 only: err", np.abs(g - w).max(), g[0, :4])
