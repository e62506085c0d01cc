"""ctypes wrapper around oracle/c/librxg_oracle.so (fp64 C twin) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_HERE, "librxg_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "rxg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # -march=native is avoided: the .so built here travels to a different host CPU
        subprocess.check_call(
            ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-std=gnu11", "-shared",
             "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.rxo_lgssm_smooth_f64.restype = ctypes.c_int
        _lib.rxo_lgssm_smooth_f64.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long,
            dp, dp, dp, dp, dp, dp, ctypes.POINTER(ctypes.c_float), dp, dp, dp, ctypes.c_int]
        _lib.rxo_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, ct=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def smooth(y, A, B, P, Q, m0, S0, nthreads=0):
    """y[T, m, batch] fp32 -> dict(mean[T,d,batch], cov[T,d,d,batch], neg_log_evidence[batch])."""
    y = np.ascontiguousarray(y, dtype=np.float32)
    T, m, batch = y.shape
    d = A.shape[0]
    f = lambda M: np.ascontiguousarray(M, dtype=np.float64)
    A, B, P, Q, m0, S0 = map(f, (A, B, P, Q, m0, S0))
    mean = np.empty((T, d, batch)); cov = np.empty((T, d, d, batch)); nle = np.empty(batch)
    rc = lib().rxo_lgssm_smooth_f64(d, m, T, batch, _p(A), _p(B), _p(P), _p(Q), _p(m0), _p(S0),
                                    _p(y, ctypes.c_float), _p(mean), _p(cov), _p(nle), nthreads)
    if rc != 0:
        raise RuntimeError(f"rxo_lgssm_smooth_f64 failed / non-SPD events: {rc}")
    return dict(mean=mean, cov=cov, neg_log_evidence=nle)


def num_threads():
    return lib().rxo_num_threads()
