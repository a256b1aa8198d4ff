#!/usr/bin/env python
"""Developer aid: the reference's Python package on top of this library, step by step, with faulthandler on."""
import faulthandler, os, sys
faulthandler.enable()
import numpy as np
from scipy.sparse import coo_matrix
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "integration", "_dropin")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import synth_coo
import cmfrec
from cmfrec import CMF, CMF_implicit, MostPopular
print(cmfrec.__file__, flush=True)
m, n = 3000, 1800
a, b, x = synth_coo(m, n, 90000, np.float64, seed=5)
X = coo_matrix((x, (a, b)), shape=(m, n))
for step in ("cmf_noprecompute", "cmf_default", "topn", "predict", "warm", "cmf_U", "implicit", "popular"):
    print("step", step, flush=True)
    if step == "cmf_noprecompute":
        CMF(k=20, nthreads=4, precompute_for_predictions=False).fit(X)
    elif step == "cmf_default":
        mod = CMF(k=20, nthreads=4).fit(X)
    elif step == "topn":
        print(mod.A_.dtype, mod.B_.dtype, float(mod.glob_mean_), mod.A_[1, :3], mod.user_bias_[:3], mod.item_bias_[:3], flush=True)
        print(mod.topN(user=7, n=10), flush=True)
        print(mod.topN(user=7, n=10, output_score=True), flush=True)
    elif step == "predict":
        print(mod.predict(user=[1, 2, 3], item=[4, 5, 6]), flush=True)
        manual = (mod.A_[[1, 2, 3]] * mod.B_[[4, 5, 6]]).sum(axis=1) + mod.glob_mean_ + mod.user_bias_[[1, 2, 3]] + mod.item_bias_[[4, 5, 6]]
        print("manual", manual, flush=True)
    elif step == "warm":
        print(mod.factors_warm(X_col=b[:5], X_val=x[:5]).shape)
    elif step == "cmf_U":
        CMF(k=12, nthreads=4, niter=3).fit(X, U=np.random.default_rng(0).normal(size=(m, 4)))
    elif step == "implicit":
        CMF_implicit(k=20, nthreads=4).fit(X)
    elif step == "popular":
        MostPopular(user_bias=True).fit(X)
print("ALL_OK")
