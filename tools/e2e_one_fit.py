import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import bench
from cmfrec_b200 import _lib
from support import fit_explicit
w = bench.WORKLOADS["ml10m_explicit_cg_k64_f32"]
a, b, x, m, n, dt = bench.load_data(w)
L = _lib.load(dt)
fit_explicit(L, dt, a, b, x, m, n, 64, lam=0.05, scale_lam=True, niter=2, nthreads=os.cpu_count())
