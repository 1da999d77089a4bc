"""ctypes wrapper of the plain-C oracle port (``kalman_c.c``) - TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kalman_c.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libkalman_c.so")
LOG_2PI = float(np.log(2 * np.pi))


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        # -march=x86-64-v3 rather than native: the .so built here must also run on the GPU box's host CPU
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", SO, "-lm"])
    return SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.kalman_c_batch.restype = ctypes.c_int
        _lib.kalman_c_batch.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7 + [
            ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.kalman_c_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(_load().kalman_c_max_threads())


def logp_grad_batch(y, a0, P0, T, Z, H, C, ll_const=LOG_2PI, want_grads=True, nthreads=0):
    """Standard filter, static matrices.  y[n,p]; a0[B,m]; P0,T,C[B,m,m]; Z[p,m], H[p,p] shared.
    Returns (ll[B], grads dict of [B,...] arrays incl. 'C' = d/d(RQR^T), n_bad)."""
    lib = _load()
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)  # noqa: E731
    y, a0, P0, T, Z, H, C = map(f8, (y, a0, P0, T, Z, H, C))
    B, m = a0.shape
    n, p = y.shape
    mm = m * m
    gsz = m + mm + mm + p * m + p * p + mm + m + p
    ll = np.empty(B)
    g = np.empty((B, gsz)) if want_grads else None
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    bad = lib.kalman_c_batch(B, n, m, p, vp(y), vp(a0), vp(P0), vp(T), vp(Z), vp(H), vp(C), ll_const, int(want_grads),
                             vp(ll), vp(g), int(nthreads))
    if bad < 0:
        raise ValueError("dims exceed the C port's static limits")
    grads = None
    if want_grads:
        o = 0
        grads = {}
        for k, shp in (("a0", (m,)), ("P0", (m, m)), ("T", (m, m)), ("Z", (p, m)), ("H", (p, p)), ("C", (m, m)),
                       ("c", (m,)), ("d", (p,))):
            sz = int(np.prod(shp))
            grads[k] = g[:, o:o + sz].reshape((B,) + shp)
            o += sz
    return ll, grads, bad
