"""ctypes binding of libfgq_cuda.so (the C ABI declared in include/fgq.h).

There is no CPU fallback: if the shared library is missing, cannot be loaded, or no CUDA
device is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# FGQ_LIB_PATH: developer aid for timing an alternative build of the same library (tools/time_tail.py)
LIB_PATH = os.environ.get("FGQ_LIB_PATH") or os.path.join(_HERE, "libfgq_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

FGQ_OK, FGQ_EDOMAIN, FGQ_ENOTIMPL, FGQ_EINVAL, FGQ_ECUDA = 0, 1, 2, 3, 100

_i64 = ctypes.c_int64
_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)

class FgqShard(ctypes.Structure):
    """``fgq_shard`` of include/fgq.h."""
    _fields_ = [("device", ctypes.c_int32), ("owner", ctypes.c_int32), ("lo", ctypes.c_int64), ("hi", ctypes.c_int64),
                ("x", ctypes.c_void_p), ("w", ctypes.c_void_p)]


_shp = ctypes.POINTER(FgqShard)
_vpp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol include/fgq.h declares
SIGNATURES = {
    "fgq_version": (ctypes.c_int, []),
    "fgq_last_error": (ctypes.c_char_p, []),
    "fgq_warnings": (ctypes.c_int, []),
    "fgq_device_count": (ctypes.c_int, []),
    "fgq_launch_count": (_i64, []),
    "fgq_host_alloc": (ctypes.c_int, [_i64, _vp]),
    "fgq_host_free": (ctypes.c_int, [_vp]),
    "fgq_gausslegendre_host": (ctypes.c_int, [_i64, _vp, _vp]),
    "fgq_gausslegendre_host_slice": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp]),
    "fgq_gausslegendre_host_pair": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "fgq_gausslegendre_device": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp, _vp]),
    "fgq_gausslegendre_device_pair": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "fgq_gausslegendre_batch_device": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fgq_gausslegendre_batch_host": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "fgq_gaussjacobi_host": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _vp, _vp]),
    "fgq_gaussjacobi_workspace_bytes": (_i64, [_i64]),
    "fgq_gaussjacobi_device": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _i64, _i64, _vp, _vp, _vp, _vp]),
    "fgq_gaussjacobi_stats": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "fgq_gaussradau_host": (ctypes.c_int, [_i64, _vp, _vp]),
    "fgq_gausslobatto_host": (ctypes.c_int, [_i64, _vp, _vp]),
    "fgq_gaussradau_device": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "fgq_gausslobatto_device": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "fgq_gaussradau_jacobi_host": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _vp, _vp]),
    "fgq_gaussradau_jacobi_device": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp]),
    "fgq_approx_besselroots": (ctypes.c_int, [ctypes.c_double, _i64, _vp]),
    "fgq_approx_besselroots_device": (ctypes.c_int, [ctypes.c_double, _i64, _vp, _vp]),
    "fgq_besselj": (ctypes.c_double, [ctypes.c_double, ctypes.c_double]),
    "fgq_gausschebyshev_host": (ctypes.c_int, [ctypes.c_int, _i64, _vp, _vp]),
    "fgq_gausschebyshev_device": (ctypes.c_int, [ctypes.c_int, _i64, _i64, _i64, _vp, _vp, _vp]),
    "fgq_gausshermite_host": (ctypes.c_int, [_i64, ctypes.c_int, _vp, _vp]),
    "fgq_gausshermite_workspace_bytes": (_i64, [_i64]),
    "fgq_gausshermite_device": (ctypes.c_int, [_i64, ctypes.c_int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "fgq_gausshermite_sweeps": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int)]),
    "fgq_gausslaguerre_host": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_int, _vp, _vp, _i64p]),
    "fgq_gausslaguerre_workspace_bytes": (_i64, [_i64]),
    "fgq_gausslaguerre_device": (ctypes.c_int, [_i64, ctypes.c_double, _vp, _vp, _vp, _vp]),
    "fgq_gausslaguerre_stats": (ctypes.c_int, [_vp, _i64p, _i64p, _i64p, _i64p]),
    "fgq_gausshermite_reduced_bound": (_i64, [_i64]),
    "fgq_gausshermite_reduced_device": (ctypes.c_int, [_i64, ctypes.c_int, _vp, _vp, _vp, _vp, _i64p, _i64p]),
    "fgq_gausshermite_reduced_host": (ctypes.c_int, [_i64, ctypes.c_int, _vp, _vp, _i64p, _i64p]),
    "fgq_gaussjacobi_batch_workspace_bytes": (_i64, [_i64]),
    "fgq_gaussjacobi_batch_device": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fgq_gaussjacobi_batch_host": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fgq_gausshermite_batch_workspace_bytes": (_i64, [_i64]),
    "fgq_gausshermite_batch_device": (ctypes.c_int, [_i64, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "fgq_gausshermite_batch_host": (ctypes.c_int, [_i64, _vp, _vp, ctypes.c_int, _vp, _vp]),
    "fgq_gausslaguerre_batch_workspace_bytes": (_i64, [_i64]),
    "fgq_gausslaguerre_batch_device": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fgq_gausslaguerre_batch_host": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fgq_rule_moments_device": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "fgq_gausslegendre_moments_fused_device": (ctypes.c_int, [_i64, _i64, _i64, ctypes.c_int, _vp, _vp]),
    "fgq_gausslegendre_integrate_device": (ctypes.c_int, [_i64, _i64, _i64, ctypes.c_char_p, _vp, _vp]),
    "fgq_rule_integrate_device": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_char_p, _vp, _vp]),
    "fgq_measure_dfma_peak": (ctypes.c_int, [_dp]),
    "fgq_measure_store_peak": (ctypes.c_int, [_i64, _dp]),
    "fgq_measure_host_store_peak": (ctypes.c_int, [_vp, _i64, ctypes.c_int, _dp]),
    "fgq_debug_legendre_theta_device": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp]),
    "fgq_debug_fastdiv_check": (ctypes.c_int, [_i64, _i64, _i64, _vp, _vp]),
    "fgq_debug_codec_stats": (ctypes.c_int, [_i64p, _i64p]),
    "fgq_debug_codec_roundtrip_host": (ctypes.c_int, [_vp, _i64, _vp, _i64p]),
    "fgq_debug_jacobi_skip_angles": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _dp, _dp]),
    "fgq_debug_jacobi_norm_tiles": (ctypes.c_int, [_i64, ctypes.c_double, ctypes.c_double, _vp]),
    "fgq_debug_codec_unpack_host": (ctypes.c_int, [_vp, _i64, _vp, _vp, _i64, ctypes.c_int, ctypes.c_int]),
    "fgq_debug_drain_stats": (ctypes.c_int, [_i64p, _i64p, _i64p, ctypes.POINTER(ctypes.c_int)]),
    # multi-GPU, one process
    "fgq_init": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "fgq_finalize": (ctypes.c_int, []),
    "fgq_num_devices": (ctypes.c_int, []),
    "fgq_device_ordinal": (ctypes.c_int, [ctypes.c_int]),
    "fgq_shard_count": (ctypes.c_int, [_i64]),
    "fgq_gausslegendre_device_sharded": (ctypes.c_int, [_i64, _shp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), _dp]),
    "fgq_free_shards": (ctypes.c_int, [_shp, ctypes.c_int]),
    "fgq_synchronize": (ctypes.c_int, []),
    "fgq_timer_start": (ctypes.c_int, []),
    "fgq_timer_stop": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "fgq_allgather": (ctypes.c_int, [_shp, ctypes.c_int, _i64, _vpp, _vpp]),
    "fgq_gausslegendre_mirror_device": (ctypes.c_int, [_i64, _vp, _vp, _vp]),
    "fgq_gausslegendre_device_replicated": (ctypes.c_int, [_i64, _vpp, _vpp]),
    "fgq_free_replicas": (ctypes.c_int, [_vpp, _vpp]),
    "fgq_gausslegendre_host_multi": (ctypes.c_int, [_i64, _vp, _vp]),
    "fgq_debug_ordered_planner": (_i64, [_i64, _i64]),
    "fgq_debug_laguerre_batch_buckets": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, ctypes.c_int]),
    "fgq_debug_half_index_range": (ctypes.c_int, [_i64, ctypes.c_int, ctypes.c_int, _i64p, _i64p]),
}

_lib = None


class FGQError(RuntimeError):
    """CUDA / internal failure inside libfgq_cuda.so."""


class DomainError(ValueError):
    """Mirror of Julia's DomainError at the reference's argument checks."""


class NotImplementedOnDevice(NotImplementedError):
    """Parameter region the reference serves off the hot path (eigen-solves, non-Float64)."""


def build(verbose: bool = False) -> str:
    """Compile libfgq_cuda.so for sm_100a with nvcc (in-tree; cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise FGQError("building libfgq_cuda.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FGQError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " — this package has no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc == FGQ_OK:
        return
    msg = lib().fgq_last_error().decode("utf-8", "replace")
    if rc == FGQ_EDOMAIN:
        raise DomainError(msg)
    if rc == FGQ_ENOTIMPL:
        raise NotImplementedOnDevice(msg)
    if rc == FGQ_EINVAL:
        raise ValueError(msg)
    raise FGQError(f"libfgq_cuda rc={rc}: {msg}")
