"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/scan_oracle.c (the plain-C checker)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzigma_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "scan_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def set_threads(n):
    """OpenMP threads of the C loops (returns the resulting maximum)."""
    return int(lib().zigma_oracle_set_threads(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def scan_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=True):
    """numpy fp32 in / out.  u, delta, z (Bt,E,L); A (E,N); B, C (Bt,N,L) or (Bt,G,N,L)."""
    u, delta, A, B, C, D, z, delta_bias = map(_f, (u, delta, A, B, C, D, z, delta_bias))
    if B.ndim == 3:
        B, C = B[:, None], C[:, None]
    Bt, E, L = u.shape
    N, G = A.shape[1], B.shape[1]
    out = np.empty_like(u)
    last = np.empty((Bt, E, N), dtype=np.float32)
    rc = lib().zigma_oracle_scan_fwd(_p(u), _p(delta), _p(A), _p(np.ascontiguousarray(B)),
                                     _p(np.ascontiguousarray(C)), _p(D), _p(z), _p(delta_bias),
                                     int(delta_softplus), _p(out), _p(last), Bt, E, L, N, G)
    assert rc == 0
    return out, last


def conv1d_fwd(x, w, bias=None, silu=True):
    x, w, bias = map(_f, (x, w, bias))
    Bt, E, L = x.shape
    out = np.empty_like(x)
    rc = lib().zigma_oracle_conv1d_fwd(_p(x), _p(w), _p(bias), int(silu), _p(out), Bt, E, L, w.shape[1])
    assert rc == 0
    return out
