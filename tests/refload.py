"""Test infrastructure: ctypes access to the checkers under oracle/.

* `ref(dtype)`   -> oracle/_ref/libcmfrec_ref_{f32,f64}.so, the unmodified reference compiled from
                    /root/reference/src by oracle/Makefile (None when it has not been built).
* `oracle(dtype)` -> oracle/_build/libcmf_oracle_{f32,f64}.so, this repo's plain-C restatement.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cmfrec_b200 import _abi  # noqa: E402

P = C.c_void_p
c_int, c_bool, c_size_t = C.c_int, C.c_bool, C.c_size_t
_cache = {}


def _tag(dtype):
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


class ArraysToFill(C.Structure):
    _fields_ = [("A", P), ("sizeA", c_size_t), ("B", P), ("sizeB", c_size_t)]


def ref(dtype):
    key = ("ref", _tag(dtype))
    if key in _cache:
        return _cache[key]
    path = os.path.join(ROOT, "oracle", "_ref", "libcmfrec_ref_%s.so" % _tag(dtype))
    if not os.path.exists(path):
        _cache[key] = None
        return None
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    real = _abi.real_ctype(dtype)
    _abi.bind_reference_names(lib, dtype)
    # internal (non-static) functions of the reference used for per-function parity
    lib.random_parallel.argtypes = [ArraysToFill, c_int, c_bool, c_int]          # src/helpers.c:930
    lib.random_parallel.restype = c_int
    lib.coo_to_csr_and_csc.argtypes = [P, P, P, P, c_int, c_int, c_size_t, P, P, P, P, P, P, P, P, c_int]  # helpers.c:1375
    lib.coo_to_csr_and_csc.restype = None
    lib.calc_mean_and_center.argtypes = [P, P, P, c_size_t, P, P, c_int, c_int, P, P, P, P, P, P, P, c_bool, c_bool,
                                         c_bool, c_int, P, P, P, c_bool]          # src/common.c:3423
    lib.calc_mean_and_center.restype = c_int
    lib.initialize_biases_twosided.argtypes = [P, P, P, P, c_int, c_int, c_bool, c_bool, C.c_double, P, P, P, P, P, P,
                                               P, P, P, P, real, real, c_bool, P, P, P, P, c_int]  # common.c:4410
    lib.initialize_biases_twosided.restype = c_int
    # optimizeA, src/common.c:2742
    lib.optimizeA.argtypes = [P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P, c_int, c_bool, c_bool, c_bool, P, P,
                              c_bool, real, real, real, real, c_bool, c_bool, P, c_bool, c_int, c_bool, c_bool, c_bool,
                              c_int, c_bool, c_int, P, P, P, real, P, real, c_bool, P, P, P, P]
    lib.optimizeA.restype = None
    # optimizeA_implicit, src/common.c:3305
    lib.optimizeA_implicit.argtypes = [P, c_size_t, P, c_size_t, c_int, c_int, c_int, P, P, P, real, real, c_int,
                                       c_bool, c_bool, c_bool, c_int, c_bool, c_int, P, P, P]
    lib.optimizeA_implicit.restype = None
    _cache[key] = lib
    return lib


def oracle(dtype):
    key = ("oracle", _tag(dtype))
    if key in _cache:
        return _cache[key]
    path = os.path.join(ROOT, "oracle", "_build", "libcmf_oracle_%s.so" % _tag(dtype))
    if not os.path.exists(path):
        _cache[key] = None
        return None
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    _cache[key] = lib
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(P)
