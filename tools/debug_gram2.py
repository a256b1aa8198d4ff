import ctypes as C, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cmfrec_b200 import _lib
from support import ptr
dt = np.dtype(np.float32)
L = _lib.load(dt)
np.set_printoptions(linewidth=250, precision=3, suppress=True)
kk = int(sys.argv[1])
G = np.ones((8, kk), dt)
out = np.full((kk, kk), -7.0, dt); ms = C.c_float(0)
rc = L.cmfb200_gram(ptr(G), 8, kk, ptr(out), 0, C.byref(ms))
print("dbg", os.environ.get("CMFB200_GRAM_DBG"), "kk", kk, "rc", rc, "unique", np.unique(out)[:8], "nnz", np.count_nonzero(out))
G = np.zeros((8, kk), dt); G[0, :] = np.arange(1, kk + 1)
rc = L.cmfb200_gram(ptr(G), 8, kk, ptr(out), 0, C.byref(ms))
print(" arange: out[0,:6]", out[0, :6], "out[:6,0]", out[:6, 0], "out[1,:6]", out[1,:6], "out[40,40:44]", out[40,40:44])
