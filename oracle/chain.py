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
        L.laplace7_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.laplace7_f32.restype = None
        L.bcast_add_axis0_sum_f32.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int64]
        L.bcast_add_axis0_sum_f32.restype = None
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_set_num_threads.restype = None
        _LIB = L
    return _LIB


def chain_f64(A, B, C, D, global_start=0, make_A=False):
    for x in (A, B, C, D):
        assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().chain_f64(A.ctypes.data, B.ctypes.data, C.ctypes.data, D.ctypes.data, A.size, global_start, 1 if make_A else 0)


def sum_affine_f32(X, mul, add):
    assert X.dtype == np.float32 and X.flags.c_contiguous
    return lib().sum_affine_f32(X.ctypes.data, X.size, mul, add)


def laplace7_f32(U, V):
    assert U.dtype == np.float32 and V.dtype == np.float32 and U.flags.c_contiguous and V.flags.c_contiguous and U.shape == V.shape
    lib().laplace7_f32(U.ctypes.data, V.ctypes.data, U.shape[0])


def bcast_add_axis0_sum_f32(M, v, red):
    assert M.dtype == np.float32 and v.dtype == np.float32 and red.dtype == np.float32 and M.flags.c_contiguous
    lib().bcast_add_axis0_sum_f32(M.ctypes.data, v.ctypes.data, red.ctypes.data, M.shape[0], M.shape[1])


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def calibrate_threads(A, B, C, D):
    """Pick the faster of {all logical CPUs, half of them (one per physical core with SMT)} for the
    chain loop — the CPU baseline should be the reference path at its best on this host."""
    import os
    import time

    cands = sorted({os.cpu_count() or 1, max(1, (os.cpu_count() or 2) // 2)}, reverse=True)
    best = None
    for t in cands:
        set_num_threads(t)
        chain_f64(A, B, C, D)
        t0 = time.perf_counter()
        chain_f64(A, B, C, D)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, t)
    set_num_threads(best[1])
    return best[1]
