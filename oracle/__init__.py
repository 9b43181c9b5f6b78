"""CPU oracle package (TEST INFRASTRUCTURE — never imported by the product).

ctypes front-end of ``oracle/_build/liboracle.so`` (built from ``oracle/*.c`` by
``oracle/Makefile``; the C files cite the reference file:line they restate) plus
the numpy/scipy restatements of the Jacobi/Hermite/Laguerre paths.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_c_i64 = ctypes.c_int64
_c_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, a few seconds)."""
    # make tracks the dependencies itself (incl. the library's csrc/fgq_math.h for the twin mode)
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        for sfx in ("", "_twin", "_jl"):
            f = getattr(L, "oracle_gausslegendre" + sfx)
            f.argtypes = [_c_i64, _c_dp, _c_dp]
            f.restype = ctypes.c_int
            f = getattr(L, "oracle_legendre_asy_half" + sfx)
            f.argtypes = [_c_i64, _c_i64, _c_i64, _c_dp, _c_dp]
            f.restype = ctypes.c_int
            f = getattr(L, "oracle_legendre_theta" + sfx)
            f.argtypes = [_c_i64, _c_i64, _c_i64, _c_dp]
            f.restype = ctypes.c_int
            f = getattr(L, "oracle_gausslegendre_batch" + sfx)
            f.argtypes = [_c_i64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_c_i64), _c_dp, _c_dp]
            f.restype = ctypes.c_int
        L.oracle_gl_rec_newton.argtypes = [ctypes.c_double, _c_i64, ctypes.c_double, _c_dp, _c_dp]
        L.oracle_gl_rec_newton.restype = None
        L.oracle_gl_rec_newton2.argtypes = [ctypes.c_double, _c_i64, ctypes.c_double, ctypes.c_int, _c_dp, _c_dp]
        L.oracle_gl_rec_newton2.restype = None
        L.oracle_eval_laguerre_rec.argtypes = [_c_i64, ctypes.c_double, ctypes.c_double, _c_dp, _c_dp]
        L.oracle_eval_laguerre_rec.restype = None
        L.oracle_innerjacobi_rec.argtypes = [_c_i64, _c_i64, _c_dp, ctypes.c_double, ctypes.c_double, _c_dp, _c_dp]
        L.oracle_innerjacobi_rec.restype = None
        for nm in ("oracle_twin_sincos_q1", "oracle_twin_sincos_medium"):
            f = getattr(L, nm)
            f.argtypes = [_c_i64, _c_dp, _c_dp, _c_dp]
            f.restype = None
        L.oracle_twin_tan_q1.argtypes = [_c_i64, _c_dp, _c_dp]
        L.oracle_twin_tan_q1.restype = None
        L.oracle_jl_sincostan.argtypes = [_c_i64, _c_dp, _c_dp, _c_dp, _c_dp]
        L.oracle_jl_sincostan.restype = None
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_set_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_dp)


def set_threads(n: int) -> int:
    """1 = single thread (what the reference does); 0 = all cores."""
    return lib().oracle_set_threads(int(n))


def _sfx(twin) -> str:
    """Elementary-function mode of the Legendre oracle: False / "libm" = glibc, True / "twin" = the
    library's own fgq_math.h on the host, "julia" = the restated Julia Base sin/cos/tan (julia_libm.h)."""
    if twin is True or twin == "twin":
        return "_twin"
    if twin == "julia":
        return "_jl"
    if twin is False or twin == "libm":
        return ""
    raise ValueError(f"unknown oracle libm mode {twin!r}")


def gausslegendre(n: int, twin=False):
    """Oracle of ``gausslegendre(n)`` (reference src/gausslegendre.jl:22-68)."""
    if n < 0:
        raise ValueError("DomainError: Input n must be a non-negative integer")
    x = np.empty(n, dtype=np.float64)
    w = np.empty(n, dtype=np.float64)
    f = getattr(lib(), "oracle_gausslegendre" + _sfx(twin))
    rc = f(n, _dp(x), _dp(w))
    if rc:
        raise RuntimeError(f"oracle_gausslegendre rc={rc}")
    return x, w


def legendre_asy_half(n: int, k0: int, k1: int, twin=False):
    """cos(theta_k), w_k for half-indices k0..k1 (1-based, inclusive) of the asy path."""
    cnt = max(k1 - k0 + 1, 0)
    x = np.empty(cnt, dtype=np.float64)
    w = np.empty(cnt, dtype=np.float64)
    f = getattr(lib(), "oracle_legendre_asy_half" + _sfx(twin))
    rc = f(n, k0, k1, _dp(x), _dp(w))
    if rc:
        raise MemoryError("oracle_legendre_asy_half")
    return x, w


def legendre_theta(n: int, k0: int, k1: int, twin=False):
    cnt = max(k1 - k0 + 1, 0)
    t = np.empty(cnt, dtype=np.float64)
    f = getattr(lib(), "oracle_legendre_theta" + _sfx(twin))
    if f(n, k0, k1, _dp(t)):
        raise MemoryError("oracle_legendre_theta")
    return t


def gausslegendre_batch(n_i: np.ndarray, twin=False):
    """CSR batch of small rules: returns (offsets, x, w)."""
    n_i = np.ascontiguousarray(n_i, dtype=np.int32)
    offsets = np.zeros(len(n_i) + 1, dtype=np.int64)
    np.cumsum(n_i, out=offsets[1:])
    x = np.empty(int(offsets[-1]), dtype=np.float64)
    w = np.empty(int(offsets[-1]), dtype=np.float64)
    f = getattr(lib(), "oracle_gausslegendre_batch" + _sfx(twin))
    rc = f(len(n_i), n_i.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
           offsets.ctypes.data_as(ctypes.POINTER(_c_i64)), _dp(x), _dp(w))
    if rc:
        raise RuntimeError(f"oracle_gausslegendre_batch rc={rc}")
    return offsets, x, w


def gl_rec_newton(x0: float, n: int, alpha: float, computeweight: bool = True):
    """gl_rec_newton(x0, n, alpha; computeweight) of src/gausslaguerre.jl:547-579 -> (xk, wk)."""
    xk = ctypes.c_double(0.0)
    wk = ctypes.c_double(0.0)
    lib().oracle_gl_rec_newton2(float(x0), int(n), float(alpha), int(bool(computeweight)), ctypes.byref(xk), ctypes.byref(wk))
    return xk.value, wk.value


def innerjacobi_rec(n: int, x: np.ndarray, a: float, b: float):
    """innerjacobi_rec! of src/gaussjacobi.jl:129-155 (C, muladd -> fma) -> (P, PP)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    P = np.empty_like(x)
    PP = np.empty_like(x)
    lib().oracle_innerjacobi_rec(int(n), len(x), _dp(x), float(a), float(b), _dp(P), _dp(PP))
    return P, PP


def twin_sincos(t: np.ndarray, medium: bool = False):
    """(sin t, cos t) from the library's fgq_math.h compiled for the host (first-quadrant or medium-range)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    s = np.empty_like(t)
    c = np.empty_like(t)
    f = lib().oracle_twin_sincos_medium if medium else lib().oracle_twin_sincos_q1
    f(len(t), _dp(t), _dp(s), _dp(c))
    return s, c


def twin_tan(t: np.ndarray):
    """tan t on (0, 3pi/4) from the library's fgq_math.h compiled for the host."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    tn = np.empty_like(t)
    lib().oracle_twin_tan_q1(len(t), _dp(t), _dp(tn))
    return tn


def julia_sincostan(t: np.ndarray):
    """(sin t, cos t, tan t) from julia_libm.h, the restatement of Julia Base's fdlibm-derived functions."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    s = np.empty_like(t)
    c = np.empty_like(t)
    tn = np.empty_like(t)
    lib().oracle_jl_sincostan(len(t), _dp(t), _dp(s), _dp(c), _dp(tn))
    return s, c, tn


def ulp_diff(got: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """|got - ref| / spacing(|ref|) (SURVEY.md §8d parity metric); 0 where both are equal."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.abs(got - ref)
    sp = np.spacing(np.abs(ref))
    out = np.where(d == 0, 0.0, d / sp)
    return out
