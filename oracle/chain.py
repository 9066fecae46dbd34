"""ORACLE (test / baseline infrastructure) — ctypes wrapper of oracle/fused_chain.c."""
import ctypes
import os

import numpy as np

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle_chain.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle_chain.so missing: run `make -C oracle`")
        L = ctypes.CDLL(path)
        L.chain_f64.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.chain_f64.restype = None
        L.sum_affine_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double]
        L.sum_affine_f32.restype = ctypes.c_double
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def chain_f64(A, B, C, D, global_start=0, make_A=False):
    for x in (A, B, C, D):
        assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().chain_f64(A.ctypes.data, B.ctypes.data, C.ctypes.data, D.ctypes.data, A.size, global_start, 1 if make_A else 0)


def sum_affine_f32(X, mul, add):
    assert X.dtype == np.float32 and X.flags.c_contiguous
    return lib().sum_affine_f32(X.ctypes.data, X.size, mul, add)


def num_threads():
    return int(lib().oracle_num_threads())
