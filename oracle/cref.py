"""ctypes loader of oracle/libmepp_ref.so (the C restatement used as a faster checker and as the
CPU baseline).  Test infrastructure only -- see oracle/__init__.py."""
import ctypes as C
import os
import subprocess

from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmepp_ref.so")
_lib = None


def build():
    src = os.path.join(_HERE, "mepp_ref.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libmepp_ref.so"], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.mepp_ref_scan.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                       C.c_void_p]
        _lib.mepp_ref_blur.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    return _lib


def scan(codes, Wf32, Wr32, thr, threads=1):
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    N, L = codes.shape
    W0 = np.ascontiguousarray(Wf32, dtype=np.float32)
    W1 = None if Wr32 is None else np.ascontiguousarray(Wr32, dtype=np.float32)
    X0 = np.empty((N, L), dtype=np.float32)
    f = lib().mepp_ref_scan

    def run(lo, hi):
        f(codes[lo:hi].ctypes.data, hi - lo, L, W0.ctypes.data, None if W1 is None else W1.ctypes.data,
          W0.shape[1], C.c_float(np.float32(-thr)), X0[lo:hi].ctypes.data)
    _threaded(run, N, threads)
    return X0


def blur(X0, margin, threads=1):
    X0 = np.ascontiguousarray(X0, dtype=np.float32)
    X = np.empty_like(X0)
    f = lib().mepp_ref_blur
    L = X0.shape[1]
    _threaded(lambda lo, hi: f(X0[lo:hi].ctypes.data, hi - lo, L, int(margin), X[lo:hi].ctypes.data),
              X0.shape[0], threads)
    return X


def _threaded(fn, n, threads):
    """Row-range threading from Python (ctypes drops the GIL); this image's gcc has no libgomp."""
    threads = max(1, min(int(threads), n))
    if threads == 1:
        fn(0, n)
        return
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: fn(int(bounds[k]), int(bounds[k + 1])), range(threads)))
