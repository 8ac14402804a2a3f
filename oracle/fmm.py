"""ORACLE (test infrastructure only) -- ctypes front-end of oracle/fmm_oracle.c.

See the header of fmm_oracle.c for what is restated and the parity status.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .rbf import RBF_NAMES

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None
kClassic = -1


def build(force=False):
    src = os.path.join(_HERE, "fmm_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = ctypes.CDLL(_LIB)
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        lib.orc_direct.restype = ci
        lib.orc_direct.argtypes = [ci, ci, ci, vp, vp, ci, vp, i64, vp, i64, vp, ci, vp]
        lib.orc_fmm.restype = ci
        lib.orc_fmm.argtypes = [ci, ci, ci, vp, vp, ci, vp, vp, vp, i64, vp, i64, vp, ci, ci, ci, ci, vp]
        lib.orc_tree_height.restype = ci
        lib.orc_tree_height.argtypes = [ci, i64]
        lib.orc_num_threads.restype = ci
        _lib = lib
    return _lib


def _params(name, params):
    params = [float(p) for p in params]
    if name in ("bh3", "th3", "bh2", "th2"):
        params = (params + [1.0, 0.0][len(params):])[:2] if len(params) < 2 else params
    assert len(params) == 2
    return np.asarray(params, dtype=np.float64)


def _prep(name, params, dim, aniso):
    p = _params(name, params)
    a = np.eye(dim) if aniso is None else np.ascontiguousarray(aniso, dtype=np.float64)
    return RBF_NAMES.index(name), p, a


def num_threads():
    return load().orc_num_threads()


def tree_height(dim, n):
    return load().orc_tree_height(dim, n)


def direct(name, params, dim, kind, src, trg, w, aniso=None, part=0, symmetric=False):
    """src/fmm/full_direct.hpp (+ self interaction when symmetric)."""
    lib = load()
    rid, p, a = _prep(name, params, dim, aniso)
    src = np.ascontiguousarray(src, dtype=np.float64).reshape(-1, dim)
    trg = src if symmetric else np.ascontiguousarray(trg, dtype=np.float64).reshape(-1, dim)
    w = np.ascontiguousarray(w, dtype=np.float64)
    kn = dim if kind in (2, 3) else 1
    out = np.zeros(len(trg) * kn)
    rc = lib.orc_direct(rid, part, dim, p.ctypes.data, a.ctypes.data, kind, src.ctypes.data, len(src),
                        trg.ctypes.data, len(trg), w.ctypes.data, int(symmetric), out.ctypes.data)
    assert rc == 0
    return out


def fmm(name, params, dim, kind, bbox_min, bbox_max, src, trg, w, order=6, d=kClassic, tree_height=0,
        aniso=None, part=0, symmetric=False):
    """The reference's evaluate() on the FMM branch with a fixed (order, d)."""
    lib = load()
    rid, p, a = _prep(name, params, dim, aniso)
    src = np.ascontiguousarray(src, dtype=np.float64).reshape(-1, dim)
    trg = src if symmetric else np.ascontiguousarray(trg, dtype=np.float64).reshape(-1, dim)
    w = np.ascontiguousarray(w, dtype=np.float64)
    bmin = np.ascontiguousarray(bbox_min, dtype=np.float64)
    bmax = np.ascontiguousarray(bbox_max, dtype=np.float64)
    kn = dim if kind in (2, 3) else 1
    out = np.zeros(len(trg) * kn)
    rc = lib.orc_fmm(rid, part, dim, p.ctypes.data, a.ctypes.data, kind, bmin.ctypes.data, bmax.ctypes.data,
                     src.ctypes.data, len(src), trg.ctypes.data, len(trg), w.ctypes.data, int(symmetric),
                     order, d, tree_height, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle FMM: degenerate configuration")
    return out
