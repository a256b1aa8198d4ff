"""Gram matrix G^T G (the cblas_tsyrk of the implicit half-sweep, reference src/common.c:3328) on the GPU:
tcgen05 tensor cores with the 3xTF32 split (fp32 library, padded row widths 64 / 128 / 256) and the FMA kernel
(everything else), against float64 arithmetic on the host."""
import ctypes as C

import numpy as np
import pytest

from support import ptr

pytestmark = pytest.mark.gpu


def _gram(L, dt, G, repeats=0):
    rows, kk = G.shape
    out = np.zeros((kk, kk), dt)
    ms = C.c_float(0)
    rc = L.cmfb200_gram(ptr(G), rows, kk, ptr(out), repeats, C.byref(ms))
    assert rc == 0, rc
    return out, ms.value


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rows,kk", [(1, 64), (7, 40), (127, 64), (128, 64), (129, 64), (5000, 64), (100003, 64), (777, 128),
                                     (40000, 100), (300, 256), (20011, 256), (3000, 200), (1000, 16), (1000, 65)])
def test_gram_matches_float64(gpu_libs, dtype, rows, kk):
    """fp32: every entry within 2e-5 of max|gram| (the 3xTF32 split keeps fp32 accuracy -- plain TF32 operands would
    sit at 5e-4; what is left is the tensor core's truncating fp32 accumulation, kept short by rotating accumulators);
    fp64: 1e-13.  The output is exactly symmetric and run-to-run identical (fixed-order reduction of the slices)."""
    dt = np.dtype(dtype)
    L = gpu_libs[dt]
    rng = np.random.default_rng(rows + kk)
    G = (rng.normal(size=(rows, kk)) * rng.lognormal(0, 1, size=(rows, 1))).astype(dt)
    got, _ = _gram(L, dt, G)
    want = G.astype(np.float64).T @ G.astype(np.float64)
    tol = 2e-5 if dt == np.float32 else 1e-13
    assert np.abs(got - want).max() <= tol * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    assert np.array_equal(got, got.T)
    again, _ = _gram(L, dt, G)
    assert np.array_equal(got, again)


def test_gram_tensor_core_path_is_faster_than_fma(gpu_libs, monkeypatch):
    """LastFM-shaped factor (358858 x 64 fp32): the tensor-core path must at least halve the FMA kernel's time
    (measured: see profiles/README.md); also pins that both paths agree."""
    dt = np.dtype(np.float32)
    L = gpu_libs[dt]
    rng = np.random.default_rng(5)
    G = rng.random((358858, 64)).astype(dt)
    got, ms_tc = _gram(L, dt, G, repeats=20)
    want = G.astype(np.float64).T @ G.astype(np.float64)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    print("gram 358858x64 fp32: %.3f ms per launch" % ms_tc)
    assert ms_tc < 0.25, ms_tc
