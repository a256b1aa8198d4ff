"""The C-ABI libraries load, export every symbol include/cmfrec_b200.h declares, and refuse to compute
without a CUDA device (no CPU fallback).  Runs without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cmfrec_b200 import _abi, _lib
from support import fit_explicit, fit_implicit, synth_coo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cmfrec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", text, flags=re.M)
    return sorted({n for n in names if n.startswith("cmfb200_") or n in _abi.REFERENCE_ENTRY_POINTS})


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_library_exports_every_declared_symbol(dtype):
    lib = _lib.load(dtype)
    syms = declared_symbols()
    assert set(_abi.PRODUCT_ENTRY_POINTS) <= set(syms), set(_abi.PRODUCT_ENTRY_POINTS) - set(syms)
    for name in syms:
        assert hasattr(lib, name), name
    assert lib.cmfb200_real_name().decode() == ("f32" if np.dtype(dtype) == np.float32 else "f64")
    assert lib.get_has_openmp() in (True, False)


def test_header_is_plain_c():
    """the header must compile as C99 (plain pointers and sizes, no C++/torch types)"""
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "cmfrec_b200.h"\nint main(void){return 0;}\n')
        for flag in ([], ["-DUSE_FLOAT"]):
            subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                                   "-I", os.path.join(ROOT, "include")] + flag + [src])


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load(np.float64)
    if lib.cmfb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    dt = np.dtype(np.float64)
    ixA, ixB, X = synth_coo(60, 40, 400, dt, seed=1)
    out = fit_explicit(lib, dt, ixA, ixB, X, 60, 40, 4, niter=1)
    assert out["rc"] == 1
    out = fit_implicit(lib, dt, ixA, ixB, X, 60, 40, 4, niter=1)
    assert out["rc"] == 1
    h = C.c_void_p()
    opt = lib.AlsOptions()
    opt.m, opt.n, opt.k, opt.world = 60, 40, 4, 1
    assert lib.cmfb200_als_create(C.byref(h), C.byref(opt), None, None, None, None, None, None) == 1


def test_missing_library_raises(monkeypatch):
    monkeypatch.setattr(_lib, "_LIBS", {})
    monkeypatch.setattr(_lib, "_HERE", "/nonexistent")
    with pytest.raises(_lib.CudaLibraryMissing):
        _lib.load(np.float32)
