"""ctypes loader for oracle/fma_helper.c (single-rounding fma over numpy arrays).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_fma.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "fma_helper.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        for name, ct in (("fma64_vec", ctypes.c_double), ("fma32_vec", ctypes.c_float)):
            fn = getattr(_lib, name)
            p = ctypes.POINTER(ct)
            fn.argtypes = [p, p, p, p, ctypes.c_size_t]
            fn.restype = None
    return _lib


def _fma(a, b, c, dtype, fname, ct):
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype), np.asarray(b, dtype), np.asarray(c, dtype))
    a, b, c = (np.ascontiguousarray(v) for v in (a, b, c))
    out = np.empty(a.shape, dtype)
    p = ctypes.POINTER(ct)
    getattr(_load(), fname)(a.ctypes.data_as(p), b.ctypes.data_as(p), c.ctypes.data_as(p),
                            out.ctypes.data_as(p), out.size)
    return out


def fma64(a, b, c):
    return _fma(a, b, c, np.float64, "fma64_vec", ctypes.c_double)


def fma32(a, b, c):
    return _fma(a, b, c, np.float32, "fma32_vec", ctypes.c_float)
