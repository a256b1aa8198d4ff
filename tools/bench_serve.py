#!/usr/bin/env python
"""Batched top-N serving (cmfb200_serve_topn) at LastFM-360K item count next to the reference's topN, one user per call on
the host cores.  Prints one JSON line.   usage: bench_serve.py [n_users] [k]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cmfrec_b200 import _lib
from support import ptr, ref
from test_gpu_serve import serve_topn
from test_gpu_popular_topn import call_topn

n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dt = np.dtype(np.float32)
L = _lib.load(dt); R = ref(dt)
m, n, n_top = 358858, 160112, 10
rng = np.random.default_rng(5)
A = rng.normal(size=(m, k)).astype(dt); B = rng.normal(size=(n, k)).astype(dt)
users = rng.choice(m, size=n_users, replace=False)
seen = [rng.choice(n, size=48, replace=False) for _ in users]
serve_topn(L, dt, A, B, None, None, 0.0, k, users[:64], seen[:64], n_top)          # warm-up (module load, pool)
t0 = time.perf_counter()
rc, ix, sc, ms = serve_topn(L, dt, A, B, None, None, 0.0, k, users, seen, n_top)
t_wall = time.perf_counter() - t0
assert rc == 0
sample = min(64, n_users)
t0 = time.perf_counter()
for j in range(sample):
    rc2, rix, rsc = call_topn(R, dt, A[users[j]], B, None, 0.0, 0.0, k, n_top, exclude=seen[j])
    assert rc2 == 0 and (rix == ix[j]).mean() >= 0.9
t_ref = (time.perf_counter() - t0) / sample
print(json.dumps(dict(what="top-%d of %d items for %d users, k=%d, fp32, 48 seen items each" % (n_top, n, n_users, k),
                      users_per_s_device=n_users / (ms * 1e-3), device_ms=ms,
                      users_per_s_whole_call=n_users / t_wall, whole_call_s=t_wall,
                      whole_call_includes="upload of A (%d MB) and B (%d MB), scores, selection, download" % (A.nbytes >> 20, B.nbytes >> 20),
                      reference_users_per_s=1.0 / t_ref, reference="topN (src/common.c:5127) per user, nthreads=%d, %d users timed" % (4, sample),
                      cores=os.cpu_count())))
