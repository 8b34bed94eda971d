"""ctypes binding + build recipe for oracle/fdtd_oracle.c -- TEST INFRASTRUCTURE ONLY.

The C oracle restates python-src/main.py:12-76 and python-src/fdtd.py:30-34 (see
the header of fdtd_oracle.c).  It is the fast checker for the GPU parity tests
and the "port" CPU baseline of bench.py; it is never imported by the product.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import numpy_oracle as npo

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fdtd_oracle.c")
_LIBS = {False: os.path.join(_HERE, "libfdtd_oracle.so"), True: os.path.join(_HERE, "libfdtd_oracle_omp.so")}
_loaded = {}


def build(force: bool = False) -> None:
    """gcc -O2 -ffp-contract=off (no FMA contraction, no fast-math); serial and OpenMP flavours."""
    for omp, out in _LIBS.items():
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(_SRC):
            continue
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c11",
               "-Wall", "-Wno-unknown-pragmas", _SRC, "-o", out]
        if omp:
            cmd.insert(1, "-fopenmp")
        subprocess.run(cmd, check=True)


def _lib(omp: bool):
    if omp not in _loaded:
        if not os.path.exists(_LIBS[omp]):
            build()
        lib = ctypes.CDLL(_LIBS[omp])
        assert lib.fdtd_oracle_abi_version() == 1
        _loaded[omp] = lib
    return _loaded[omp]


_SFX = {np.dtype(np.float32): ("f32", ctypes.c_float), np.dtype(np.float64): ("f64", ctypes.c_double)}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(Ez, Hx, Hy):
    R, C = Ez.shape
    assert Hx.shape == (R, C - 1) and Hy.shape == (R - 1, C), "reference shapes (main.py:79-85)"
    for a in (Ez, Hx, Hy):
        assert a.dtype == Ez.dtype and a.flags.c_contiguous
    return R, C


def update_h(Ez, Hx, Hy, ch, omp=False):
    """In-place H half-step; ``ch`` is the (R, C) map dt/(mu*dx)."""
    R, C = _check(Ez, Hx, Hy)
    sfx, _ = _SFX[Ez.dtype]
    ch = np.ascontiguousarray(ch, dtype=Ez.dtype)
    getattr(_lib(omp), f"fdtd_oracle_update_h_{sfx}")(_ptr(Ez), _ptr(Hx), _ptr(Hy), _ptr(ch), R, C)
    return Hx, Hy


def update_e(Ez, Hx, Hy, ce, coef, omp=False):
    """In-place Ez step (interior + Mur + corners); ``ce`` is the (R, C) map dt/(eps*dx)."""
    R, C = _check(Ez, Hx, Hy)
    sfx, ct = _SFX[Ez.dtype]
    ce = np.ascontiguousarray(ce, dtype=Ez.dtype)
    prev = np.empty_like(Ez)
    fn = getattr(_lib(omp), f"fdtd_oracle_update_e_{sfx}")
    fn.argtypes = [ctypes.c_void_p] * 4 + [ct, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    fn(_ptr(Ez), _ptr(Hx), _ptr(Hy), _ptr(ce), ct(float(coef)), R, C, _ptr(prev))
    return Ez


def run(Ez, Hx, Hy, ce, ch, coef, nsteps, amp=None, src_cells=None, probes=None, omp=False):
    """``nsteps`` leapfrog steps in place.  amp: float64[nsteps] source values (or None);
    src_cells: list of (row, col) sharing ``amp``; probes: list of (row, col).
    Returns the probe trace (nsteps, n_probe) in the run dtype, or None."""
    R, C = _check(Ez, Hx, Hy)
    sfx, ct = _SFX[Ez.dtype]
    ce = np.ascontiguousarray(ce, dtype=Ez.dtype)
    ch = np.ascontiguousarray(ch, dtype=Ez.dtype)
    amp_p, src_p, n_src = None, None, 0
    if amp is not None:
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        assert amp.shape[0] >= nsteps
        src = np.ascontiguousarray(np.asarray(src_cells, dtype=np.int32).reshape(-1, 2))
        amp_p, src_p, n_src = _ptr(amp), _ptr(src), src.shape[0]
    trace, pr_p, n_pr, tr_p = None, None, 0, None
    if probes is not None and len(probes):
        pr = np.ascontiguousarray(np.asarray(probes, dtype=np.int32).reshape(-1, 2))
        trace = np.zeros((nsteps, pr.shape[0]), dtype=Ez.dtype)
        pr_p, n_pr, tr_p = _ptr(pr), pr.shape[0], _ptr(trace)
    fn = getattr(_lib(omp), f"fdtd_oracle_run_{sfx}")
    fn.argtypes = [ctypes.c_void_p] * 5 + [ct] + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int,
                                                                       ctypes.c_void_p, ctypes.c_int,
                                                                       ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    rc = fn(_ptr(Ez), _ptr(Hx), _ptr(Hy), _ptr(ce), _ptr(ch), ct(float(coef)), R, C, nsteps, amp_p, n_src,
            src_p, n_pr, pr_p, tr_p)
    if rc != 0:
        raise MemoryError("fdtd_oracle_run: scratch allocation failed")
    return trace


def coefficients(eps, mu, dt, dx, dtype):
    """(ce, ch, mur_coef) in ``dtype`` exactly as the reference forms them per step."""
    eps = np.asarray(eps, dtype=dtype)
    mu = np.asarray(mu, dtype=dtype)
    return npo.e_coeff(eps, dt, dx), npo.h_coeff(mu, dt, dx), npo.mur_coef(mu, eps, dt, dx)
