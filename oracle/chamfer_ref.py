"""ORACLE (test infrastructure): ctypes front for oracle/chamfer_ref.c (CPU, pthreads)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libchamfer_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "chamfer_ref.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        _lib.sc_oracle_chamfer_forward.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, ip, ip]
        _lib.sc_oracle_chamfer_forward.restype = ctypes.c_int
        _lib.sc_oracle_chamfer_backward.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    fp, fp, ip, ip, fp, fp]
        _lib.sc_oracle_chamfer_backward.restype = ctypes.c_int
        _lib.sc_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.sc_oracle_get_threads.restype = ctypes.c_int
    return _lib


def set_threads(t):
    _load().sc_oracle_set_threads(int(t))


def get_threads():
    return int(_load().sc_oracle_get_threads())


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def chamfer_forward(xyz1, xyz2):
    """xyz1 [B,N,3], xyz2 [B,M,3] float32 numpy -> dist1 [B,N], dist2 [B,M] (squared), idx1, idx2 int32."""
    lib = _load()
    a = np.ascontiguousarray(xyz1, dtype=np.float32)
    b = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    d1, d2 = np.zeros((B, N), np.float32), np.zeros((B, M), np.float32)
    i1, i2 = np.zeros((B, N), np.int32), np.zeros((B, M), np.int32)
    lib.sc_oracle_chamfer_forward(_f(a), _f(b), B, N, M, _f(d1), _f(d2), _i(i1), _i(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    lib = _load()
    a = np.ascontiguousarray(xyz1, dtype=np.float32)
    b = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    g1, g2 = np.zeros_like(a), np.zeros_like(b)
    lib.sc_oracle_chamfer_backward(_f(a), _f(b), B, N, M,
                                   _f(np.ascontiguousarray(graddist1, np.float32)),
                                   _f(np.ascontiguousarray(graddist2, np.float32)),
                                   _i(np.ascontiguousarray(idx1, np.int32)), _i(np.ascontiguousarray(idx2, np.int32)),
                                   _f(g1), _f(g2))
    return g1, g2


def chamfer_forward_numpy(xyz1, xyz2):
    """Slow independent cross-check of the C file for tiny inputs (float32 emulation of the fma chain
    through float64: exact for the products, so only used with inputs whose squares are exact)."""
    a = np.asarray(xyz1, np.float32)
    b = np.asarray(xyz2, np.float32)
    diff = b[:, None, :, :] - a[:, :, None, :]
    d = (diff.astype(np.float64) ** 2).sum(-1)
    return d.min(-1).astype(np.float32), d.argmin(-1).astype(np.int32)
