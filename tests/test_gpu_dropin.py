"""The reference's UNMODIFIED Python package (cmfrec/__init__.py + its Cython shim) running on top of this library
(integration/build_dropin.py; installed by __graft_entry__.build() into integration/_dropin where the reference
sources exist).  CMF / CMF_implicit / MostPopular are fitted with their PYTHON DEFAULTS (precompute_for_predictions,
finalize_chol, ...) and compared with the reference's C entry points (oracle/_ref) driven with the same settings."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from support import fit_explicit, fit_implicit, ref, rel_err, rows_match, synth_coo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "integration", "_dropin")

SCRIPT = r'''
import json, sys
import numpy as np
from scipy.sparse import coo_matrix
sys.path.insert(0, %(dropin)r)
import cmfrec
from cmfrec import CMF, CMF_implicit, MostPopular
assert cmfrec.__file__.startswith(%(dropin)r), cmfrec.__file__
z = np.load(%(data)r)
m, n = int(z["m"]), int(z["n"])
X = coo_matrix((z["x"], (z["a"], z["b"])), shape=(m, n))
Xc = coo_matrix((z["c"], (z["a"], z["b"])), shape=(m, n))
out = {}
mod = CMF(k=20, nthreads=4, verbose=False).fit(X)              # everything else: Python defaults
np.save(%(out)r + "_cmf_A.npy", mod.A_); np.save(%(out)r + "_cmf_B.npy", mod.B_)
np.save(%(out)r + "_cmf_bA.npy", mod.user_bias_); np.save(%(out)r + "_cmf_bB.npy", mod.item_bias_)
out["cmf_glob_mean"] = float(mod.glob_mean_)
ix, sc = mod.topN(user=7, n=10, output_score=True)
out["cmf_topn"] = [int(v) for v in ix]
out["cmf_pred"] = [float(v) for v in mod.predict(user=[1, 2, 3], item=[4, 5, 6])]
out["cmf_warm_shape"] = list(mod.factors_warm(X_col=z["b"][:5], X_val=z["x"][:5]).shape)
U = np.random.default_rng(0).normal(size=(m, 4)).astype(np.float32)
mu = CMF(k=12, nthreads=4, niter=3).fit(X, U=U)                 # side information with the defaults (precompute on)
np.save(%(out)r + "_cmfu_A.npy", mu.A_); np.save(%(out)r + "_cmfu_C.npy", mu.C_)
mi = CMF_implicit(k=20, nthreads=4).fit(Xc)
np.save(%(out)r + "_imp_A.npy", mi.A_); np.save(%(out)r + "_imp_B.npy", mi.B_)
out["imp_topn"] = [int(v) for v in mi.topN(user=3, n=10)]
mp = MostPopular(user_bias=True).fit(X)
np.save(%(out)r + "_mp_bB.npy", mp.item_bias_)
json.dump(out, open(%(out)r + ".json", "w"))
print("DROPIN_OK")
'''


def test_python_package_on_top_of_the_gpu_library(tmp_path):
    if not os.path.isdir(os.path.join(DROPIN, "cmfrec")):
        pytest.fail("integration/_dropin is not built (python -c 'import __graft_entry__ as g; g.build()' where "
                    "/root/reference exists)")
    dt = np.dtype(np.float32)          # use_float=True is the Python default
    R = ref(dt)
    if R is None:
        pytest.fail("oracle/_ref is not built")
    m, n = 3000, 1800
    a, b, x = synth_coo(m, n, 90000, dt, seed=5)
    _, _, c = synth_coo(m, n, 90000, dt, seed=5, kind="counts")
    c = c[: x.size]
    data = str(tmp_path / "data.npz")
    np.savez(data, a=a, b=b, x=x, c=c, m=m, n=n)
    out = str(tmp_path / "out")
    proc = subprocess.run([sys.executable, "-c", SCRIPT % dict(dropin=DROPIN, data=data, out=out)], cwd=ROOT,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0 and "DROPIN_OK" in proc.stdout, proc.stdout[-3000:]
    res = json.load(open(out + ".json"))
    # ---- CMF with the Python defaults: k given, lambda_=10, method=als, use_cg, finalize_chol, biases, centring, niter=10
    want = fit_explicit(R, dt, a, b, x, m, n, 20, lam=10.0, niter=10, use_cg=True, max_cg_steps=3, finalize_chol=True,
                        nthreads=4, precompute=True)
    assert want["rc"] == 0
    A = np.load(out + "_cmf_A.npy"); B = np.load(out + "_cmf_B.npy")
    assert A.dtype == np.float32
    assert res["cmf_glob_mean"] == float(want["glob_mean"])
    assert rows_match(A, want["A"], 5e-3, 0.001), rel_err(A, want["A"])
    assert rows_match(B, want["B"], 5e-3, 0.001), rel_err(B, want["B"])
    assert rel_err(np.load(out + "_cmf_bA.npy"), want["biasA"]) < 5e-3
    assert rel_err(np.load(out + "_cmf_bB.npy"), want["biasB"]) < 5e-3
    # topN of the fitted model: the reference's scores of the returned items are the 10 best up to float32 noise
    # (seen items are not excluded by default)
    scores = want["B"].astype(np.float64) @ want["A"][7] + want["biasB"]
    tenth = np.sort(scores)[-10]
    assert len(set(res["cmf_topn"])) == 10 and (scores[res["cmf_topn"]] >= tenth - 1e-2).all()
    pred = np.einsum("ij,ij->i", want["A"][[1, 2, 3]], want["B"][[4, 5, 6]]) + want["glob_mean"] + want["biasA"][[1, 2, 3]] + want["biasB"][[4, 5, 6]]
    assert np.allclose(res["cmf_pred"], pred, rtol=0, atol=2e-2)
    assert res["cmf_warm_shape"] == [20]
    # ---- CMF with side information and the default precompute_for_predictions=True
    U = np.random.default_rng(0).normal(size=(m, 4)).astype(np.float32)
    wantu = fit_explicit(R, dt, a, b, x, m, n, 12, lam=10.0, niter=3, use_cg=True, finalize_chol=True, nthreads=4, U=U)
    Au = np.load(out + "_cmfu_A.npy")
    assert rows_match(Au, wantu["A"], 5e-3, 0.001), rel_err(Au, wantu["A"])
    assert rel_err(np.load(out + "_cmfu_C.npy"), wantu["C"]) < 5e-3
    # ---- CMF_implicit defaults: lambda_=1, alpha=1, use_cg, finalize_chol=False, niter=10
    wanti = fit_implicit(R, dt, a, b, c, m, n, 20, lam=1.0, alpha=1.0, niter=10, use_cg=True, finalize_chol=False, nthreads=4)
    Ai = np.load(out + "_imp_A.npy")
    # float32 implicit fits: bulk of the rows (the truncated CG's step-count flips spread through the alternation, see
    # test_gpu_fit.py::test_implicit_fit_matches_reference)
    row_err = np.abs(Ai.astype(np.float64) - wanti["A"]).max(axis=1) / np.abs(wanti["A"]).max()
    assert np.quantile(row_err, 0.5) < 1e-2 and np.quantile(row_err, 0.9) < 5e-2, np.quantile(row_err, [0.5, 0.9, 0.99])
    assert len(res["imp_topn"]) == 10
