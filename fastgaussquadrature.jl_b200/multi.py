"""Single-process multi-GPU face of the C ABI (``fgq_init`` ... ``fgq_allgather``, include/fgq.h):
what a Julia ``ccall`` client uses to drive 1/2/4/8 B200s from one process, mirrored here over ctypes so
that the tests and ``bench.py --driver capi`` exercise exactly those symbols (no torchrun, no
``torch.distributed``; torch is only used to LOOK at the device buffers).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from ._lib import DomainError, FgqShard, check, lib


def init(device_ids: Optional[Sequence[int]] = None, ndev: int = 0) -> int:
    """``fgq_init``: select the devices (default: all visible); returns their number."""
    L = lib()
    if device_ids is None:
        check(L.fgq_init(None, int(ndev)))
    else:
        ids = (ctypes.c_int * len(device_ids))(*[int(d) for d in device_ids])
        check(L.fgq_init(ids, len(device_ids)))
    return int(L.fgq_num_devices())


def finalize() -> None:
    check(lib().fgq_finalize())


def num_devices() -> int:
    return int(lib().fgq_num_devices())


def synchronize() -> None:
    check(lib().fgq_synchronize())


def timer_start() -> None:
    check(lib().fgq_timer_start())


def timer_stop():
    """(slowest device's milliseconds, per-device milliseconds) since ``timer_start``."""
    nd = max(num_devices(), 1)
    per = (ctypes.c_float * nd)()
    worst = ctypes.c_float(0.0)
    check(lib().fgq_timer_stop(ctypes.byref(worst), per))
    return float(worst.value), [float(v) for v in per]


class _DevView:
    """Zero-copy view of a device buffer for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, count: int, device: int):
    import torch
    if count == 0:
        return torch.empty(0, dtype=torch.float64, device=f"cuda:{device}")
    with torch.cuda.device(device):
        return torch.as_tensor(_DevView(ptr, count), device=f"cuda:{device}")


class ShardedRule:
    """Handle of ``fgq_gausslegendre_device_sharded``: the library-owned shards of an n-point rule."""

    def __init__(self, n: int):
        self.n = int(n)
        self.capacity = max(int(lib().fgq_shard_count(self.n)), 1)
        self.shards = (FgqShard * self.capacity)()  # zero-initialised: allocated by the first fill
        self.count = 0
        self.sums = None

    def fill(self, sums: bool = True) -> "ShardedRule":
        cnt = ctypes.c_int(0)
        s = (ctypes.c_double * 2)() if sums else None
        check(lib().fgq_gausslegendre_device_sharded(self.n, self.shards, self.capacity, ctypes.byref(cnt), s))
        self.count = int(cnt.value)
        self.sums = (float(s[0]), float(s[1])) if sums else None
        return self

    def segments(self) -> List[tuple]:
        """[(device, lo, hi, x, w)] with x, w zero-copy torch views of the device buffers."""
        out = []
        for i in range(self.count):
            s = self.shards[i]
            c = int(s.hi - s.lo)
            out.append((int(s.device), int(s.lo), int(s.hi), _as_tensor(s.x or 0, c, s.device), _as_tensor(s.w or 0, c, s.device)))
        return out

    def to_host(self):
        """The whole rule as numpy arrays (tests): every segment copied back into place."""
        x = np.full(self.n, np.nan)
        w = np.full(self.n, np.nan)
        for _, lo, hi, xs, ws in self.segments():
            x[lo:hi] = xs.cpu().numpy()
            w[lo:hi] = ws.cpu().numpy()
        return x, w

    def free(self) -> None:
        if self.count:
            check(lib().fgq_free_shards(self.shards, self.count))
            self.count = 0

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


def gausslegendre_sharded(n, out: Optional[ShardedRule] = None, sums: bool = True) -> ShardedRule:
    """Device-resident ``gausslegendre(n)`` over all initialised devices of THIS process (reference
    src/gausslegendre.jl:22-69).  ``out`` reuses the buffers of a previous call; ``sums=False`` only enqueues."""
    n = int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    if out is None:
        out = ShardedRule(n)
    elif out.n != n:
        raise ValueError("out belongs to a different n")
    return out.fill(sums)


class Replicas:
    """Per-device full copies of a rule (``fgq_allgather`` / ``fgq_gausslegendre_device_replicated``)."""

    def __init__(self, n: int):
        self.n = int(n)
        nd = max(num_devices(), 1)
        self.x = (ctypes.c_void_p * nd)()
        self.w = (ctypes.c_void_p * nd)()
        self.ndev = nd

    def tensors(self, d: int):
        dev = int(lib().fgq_device_ordinal(d))   # replicas are listed in fgq_init order
        return _as_tensor(self.x[d] or 0, self.n, dev), _as_tensor(self.w[d] or 0, self.n, dev)

    def free(self) -> None:
        if any(self.x):
            check(lib().fgq_free_replicas(self.x, self.w))

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


def allgather(rule: ShardedRule, out: Optional[Replicas] = None) -> Replicas:
    """Replicate the sharded rule on every device (transport: FGQ_ALLGATHER=push|ce|nccl)."""
    out = out or Replicas(rule.n)
    check(lib().fgq_allgather(rule.shards, rule.count, rule.n, out.x, out.w))
    return out


def gausslegendre_replicated(n, out: Optional[Replicas] = None) -> Replicas:
    """The replicated rule by generation on every device (no transfer)."""
    out = out or Replicas(int(n))
    check(lib().fgq_gausslegendre_device_replicated(int(n), out.x, out.w))
    return out


def gausslegendre_host_multi(n, out=None, pinned: bool = False):
    """Host-returning ``gausslegendre(n)`` drained through every initialised device at once."""
    from .api import _host_pair, _ptr, _check_out
    n = int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    if out is None:
        x, w = _host_pair(n, pinned)
    else:
        x, w = out
        _check_out(x, w, n)
    if n:
        check(lib().fgq_gausslegendre_host_multi(n, _ptr(x), _ptr(w)))
    return x, w
