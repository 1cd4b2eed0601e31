"""ctypes loader of oracle/liboracle.so (plain-C op oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "bvg_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def activation1d(x, up, dn, alpha, beta, logscale=True):
    x = _f(x); y = np.empty_like(x); B, Cc, T = x.shape
    _load().oracle_activation1d(_p(x), _p(y), B, Cc, T, _p(_f(up)), _p(_f(dn)), _p(_f(alpha)), _p(_f(beta)),
                                int(logscale))
    return y


def conv1d(x, w, bias, dilation=1, resid=None):
    x = _f(x); w = _f(w); B, Cin, T = x.shape; Cout, _, k = w.shape
    y = np.empty((B, Cout, T), np.float32)
    r = _f(resid) if resid is not None else None
    _load().oracle_conv1d(_p(x), _p(y), _p(r), B, Cin, Cout, T, _p(w), _p(_f(bias)), k, int(dilation))
    return y


def conv_transpose1d(x, w, bias, u):
    x = _f(x); w = _f(w); B, Cin, T = x.shape; _, Cout, k = w.shape
    y = np.empty((B, Cout, T * u), np.float32)
    _load().oracle_conv_transpose1d(_p(x), _p(y), B, Cin, Cout, T, _p(w), _p(_f(bias)), k, int(u))
    return y
