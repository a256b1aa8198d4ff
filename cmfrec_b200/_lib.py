"""Loader for the in-tree CUDA libraries (cmfrec_b200/lib/libcmfrec_b200_{f32,f64}.so).

There is no fallback of any kind: if the library has not been built this raises, and if it is loaded on a
machine without a CUDA device every compute entry point returns an error code that the wrappers turn into
an exception.
"""
import ctypes
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


class CudaLibraryMissing(ImportError):
    pass


def lib_path(dtype):
    tag = "f32" if np.dtype(dtype) == np.float32 else "f64"
    return os.path.join(_HERE, "lib", "libcmfrec_b200_%s.so" % tag)


def load(dtype):
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("dtype must be float32 or float64")
    if dtype in _LIBS:
        return _LIBS[dtype]
    path = lib_path(dtype)
    if not os.path.exists(path):
        raise CudaLibraryMissing(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cmfrec_b200/csrc`). cmfrec_b200 has no CPU fallback." % path)
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    _abi.bind_product(lib, dtype)
    _LIBS[dtype] = lib
    return lib


def ptr(arr):
    """Raw pointer of a C-contiguous numpy array, or NULL for None."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.c_void_p)
