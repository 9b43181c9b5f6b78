"""Reference-facing functions.  Same names, argument meaning and error behaviour as the Julia
package; each returns ``(x, w)`` as freshly allocated float64 arrays, nodes ascending.

Host-returning calls (`gausslegendre(n)` ...) hand back numpy arrays, exactly like the
reference hands back ``Vector{Float64}``.  The ``*_device`` calls are the device-resident
variant: torch CUDA tensors that stay in HBM, sharded by rank (one process per GPU).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import DomainError, check, lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _as_int(n) -> int:
    if isinstance(n, bool) or int(n) != n:
        raise TypeError(f"n must be an integer, got {n!r}")  # Julia: MethodError for non-Integer n
    return int(n)


class _PinnedBlock:
    """Owner of one fgq_host_alloc block (released when the last array viewing it is collected)."""

    def __init__(self, nbytes: int):
        p = ctypes.c_void_p()
        check(lib().fgq_host_alloc(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes

    def __del__(self):
        if getattr(self, "ptr", None):
            try:
                lib().fgq_host_free(self.ptr)
            except Exception:
                pass
            self.ptr = None


def empty_pinned(n: int) -> np.ndarray:
    """A length-n float64 array in page-locked memory from the library (``fgq_host_alloc``): a destination the
    ``*_host`` calls fill by DMA, without staging and without first-touch page faults.  Reuse it across calls."""
    n = int(n)
    if n < 0:
        raise ValueError("n must be non-negative")
    if n == 0:
        return np.empty(0, dtype=np.float64)
    block = _PinnedBlock(8 * n)
    buf = (ctypes.c_double * n).from_address(block.ptr)
    buf._fgq_owner = block            # the ctypes array is the ndarray's base: keeps the block alive
    return np.frombuffer(buf, dtype=np.float64, count=n)


def _host_pair(n: int, pinned: bool):
    """Two length-n float64 host arrays; pinned (page-locked) ones come from fgq_host_alloc."""
    if pinned:
        return empty_pinned(n), empty_pinned(n)
    return np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)


def _check_out(x, w, cnt: int) -> None:
    """Caller-supplied result buffers: the C side writes `cnt` CONTIGUOUS doubles through the base pointer."""
    for a in (x, w):
        if not isinstance(a, np.ndarray) or a.shape != (cnt,) or a.dtype != np.float64:
            raise ValueError("out must be two float64 arrays of the requested length")
        if not a.flags.c_contiguous or not a.flags.writeable:
            raise ValueError("out arrays must be C-contiguous and writeable (a strided view would be overrun)")


def _check_offsets(n_i: np.ndarray, offsets: np.ndarray) -> None:
    """CSR layout of a batch: len(n_i) + 1 non-decreasing offsets, rule r inside [offsets[r], offsets[r+1]), so
    that offsets[-1] (the size of x and w) bounds everything the C side writes."""
    if offsets.ndim != 1 or len(offsets) != len(n_i) + 1:
        raise ValueError(f"offsets must have len(n_i) + 1 = {len(n_i) + 1} entries, got {offsets.shape}")
    if len(n_i) == 0:
        return
    if offsets[0] < 0 or np.any(np.diff(offsets) < 0):
        raise ValueError("offsets must be non-negative and non-decreasing")
    if np.any(offsets[:-1] + np.maximum(n_i, 0) > offsets[1:]):
        raise ValueError("offsets[r] + n_i[r] must not exceed offsets[r+1]")


def launch_count() -> int:
    return int(lib().fgq_launch_count())


# --------------------------------------------------------------------------------------------
# Gauss-Legendre  (reference src/gausslegendre.jl:22-69)
# --------------------------------------------------------------------------------------------

def gausslegendre(n, *, pinned: bool = False, out: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                  index_range: Optional[Tuple[int, int]] = None):
    """``gausslegendre(n) -> x, w`` (reference src/gausslegendre.jl:22,69).

    n < 0 raises DomainError (gausslegendre.jl:26); n == 0 returns two empty arrays (:28).
    ``out=(x, w)`` reuses caller buffers (e.g. pinned ones) instead of allocating.
    ``index_range=(lo, hi)`` returns only nodes lo..hi-1 (0-based) — one process per GPU can
    fetch its own part of a large rule.
    """
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    lo, hi = (0, n) if index_range is None else (int(index_range[0]), int(index_range[1]))
    if not (0 <= lo <= hi <= n):
        raise ValueError(f"invalid index_range {index_range} for n={n}")
    cnt = hi - lo
    if out is not None:
        x, w = out
        _check_out(x, w, cnt)
    else:
        x, w = _host_pair(cnt, pinned)
    if cnt == 0:
        return x, w
    if index_range is None:
        check(lib().fgq_gausslegendre_host(n, _ptr(x), _ptr(w)))
    else:
        check(lib().fgq_gausslegendre_host_slice(n, lo, hi, _ptr(x), _ptr(w)))
    return x, w


def gausslegendre_host_pair(n, rank: int = 0, world: int = 1, out=None, pinned: bool = True):
    """This rank's mirror-pair shard of ``gausslegendre(n)`` in HOST memory: returns
    ``[(lo, hi, x, w), (lo, hi, x, w)]`` (left segment, mirrored right segment) like
    ``gausslegendre_device(mode="pair")``.  Only the left segment crosses PCIe."""
    n = _as_int(n)
    k_lo, k_hi = half_index_partition(n, world)[rank]
    cnt = k_hi - k_lo
    m = (n + 1) // 2
    skip = 1 if (n % 2 == 1 and k_hi - 1 == m and cnt > 0) else 0
    if out is None:
        xl, wl = _host_pair(cnt, pinned)
        xr, wr = _host_pair(cnt, pinned)
    else:
        xl, wl, xr, wr = out
        _check_out(xl, wl, cnt)
        _check_out(xr, wr, cnt)
    if cnt:
        check(lib().fgq_gausslegendre_host_pair(n, k_lo, k_hi, _ptr(xl), _ptr(wl), _ptr(xr), _ptr(wr)))
    return [(k_lo - 1, k_hi - 1, xl, wl), (n - k_hi + 1 + skip, n - k_lo + 1, xr[skip:], wr[skip:])], (xl, wl, xr, wr)


def gausslegendre_batch(n_i, offsets=None):
    """Batch of small rules ``gausslegendre(n_i[r])``, 0 <= n_i <= 60, CSR layout.

    Returns ``(offsets, x, w)``: rule r occupies ``x[offsets[r]:offsets[r+1]]``.  One warp per
    rule on the device (closed forms / Newton on the recurrence, gausslegendre.jl:27-60,233-309).
    """
    n_i = np.ascontiguousarray(n_i, dtype=np.int32)
    if n_i.ndim != 1:
        raise ValueError("n_i must be one-dimensional")
    if offsets is None:
        offsets = np.zeros(len(n_i) + 1, dtype=np.int64)
        np.cumsum(np.maximum(n_i, 0), out=offsets[1:])
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    _check_offsets(n_i, offsets)
    total = int(offsets[-1]) if len(offsets) else 0
    x = np.empty(total, dtype=np.float64)
    w = np.empty(total, dtype=np.float64)
    if len(n_i):
        check(lib().fgq_gausslegendre_batch_host(len(n_i), _ptr(n_i), _ptr(offsets), _ptr(x), _ptr(w)))
    return offsets, x, w


def _batch_layout(n_i, offsets):
    n_i = np.ascontiguousarray(n_i, dtype=np.int32)
    if n_i.ndim != 1:
        raise ValueError("n_i must be one-dimensional")
    if offsets is None:
        offsets = np.zeros(len(n_i) + 1, dtype=np.int64)
        np.cumsum(np.maximum(n_i, 0), out=offsets[1:])
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    _check_offsets(n_i, offsets)
    total = int(offsets[-1]) if len(offsets) else 0
    return n_i, offsets, np.empty(total, dtype=np.float64), np.empty(total, dtype=np.float64)


def _batch_param(v, count):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (count,)))
    return a


def gaussjacobi_batch(n_i, alpha, beta, offsets=None):
    """Batch of small rules ``gaussjacobi(n_i[r], alpha[r], beta[r])``, 0 <= n_i <= 100, CSR layout
    (``alpha`` / ``beta`` broadcast against ``n_i``).  Returns ``(offsets, x, w)``.  Every rule takes the
    branch gaussjacobi.jl:23-55 dispatches it to; one block per ``jacobi_rec`` rule (:61-155), bit-identical
    to the single calls."""
    n_i, offsets, x, w = _batch_layout(n_i, offsets)
    a, b = _batch_param(alpha, len(n_i)), _batch_param(beta, len(n_i))
    if len(n_i):
        check(lib().fgq_gaussjacobi_batch_host(len(n_i), _ptr(n_i), _ptr(a), _ptr(b), _ptr(offsets), _ptr(x), _ptr(w)))
    return offsets, x, w


def gausshermite_batch(n_i, offsets=None, normalize=False):
    """Batch of small rules ``gausshermite(n_i[r]; normalize)``, 0 <= n_i <= 200, CSR layout; one block per
    rule (hermite_gw / hermite_rec, gausshermite.jl:91-137,325-341), bit-identical to the single calls."""
    n_i, offsets, x, w = _batch_layout(n_i, offsets)
    if len(n_i):
        check(lib().fgq_gausshermite_batch_host(len(n_i), _ptr(n_i), _ptr(offsets), int(bool(normalize)), _ptr(x), _ptr(w)))
    return offsets, x, w


def gausslaguerre_batch(n_i, alpha=0.0, offsets=None):
    """Batch of small rules ``gausslaguerre(n_i[r], alpha[r])``, 0 <= n_i <= 127, CSR layout; one thread per
    rule (closed forms, gausslaguerre_GW, gausslaguerre_rec: gausslaguerre.jl:70-95), bit-identical to the
    single calls."""
    n_i, offsets, x, w = _batch_layout(n_i, offsets)
    a = _batch_param(alpha, len(n_i))
    if len(n_i):
        check(lib().fgq_gausslaguerre_batch_host(len(n_i), _ptr(n_i), _ptr(a), _ptr(offsets), _ptr(x), _ptr(w)))
    return offsets, x, w


# ---- device-resident, sharded -----------------------------------------------------------------

def index_partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous node-index ranges [lo, hi) per rank, boundaries on even indices so that every
    shard's 16-byte vector stores stay aligned."""
    b = [((r * n) // world) & ~1 for r in range(world)] + [n]
    return [(b[r], b[r + 1]) for r in range(world)]


def half_index_partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Mirror-pair sharding: rank r owns half-indices [k_lo, k_hi) (1-based), i.e. the contiguous
    node-index range [k_lo-1, k_hi-1) AND its mirror image [n-k_hi+1, n-k_lo+1)."""
    m = (n + 1) // 2
    b = [((r * m) // world) & ~1 for r in range(world)] + [m]
    return [(b[r] + 1, b[r + 1] + 1) for r in range(world)]


@dataclass
class LegendreShard:
    """What one rank holds of an n-point rule: a list of (lo, hi, x, w) segments, x and w torch
    CUDA tensors holding output indices [lo, hi)."""
    n: int
    rank: int
    world: int
    mode: str
    segments: List[tuple] = field(default_factory=list)

    @property
    def count(self) -> int:
        return sum(hi - lo for lo, hi, _, _ in self.segments)


def _stream_ptr(stream) -> int:
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return int(s.cuda_stream)


def gausslegendre_device(n, rank: int = 0, world: int = 1, mode: str = "pair", out: Optional[LegendreShard] = None,
                         stream=None) -> LegendreShard:
    """Device-resident ``gausslegendre(n)``: this rank's shard stays in HBM.

    mode="pair": rank owns a half-index range; each half-index is evaluated once and stored
    as +-x (two segments).  mode="contiguous": rank owns one node-index range [lo, hi).
    Asynchronous on the current torch stream; pass ``out`` (a previous result) to reuse buffers.
    No collective is involved: every node is a function of (n, index) alone.
    """
    import torch
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    if not torch.cuda.is_available():
        raise _lib.FGQError("gausslegendre_device needs a CUDA device (no CPU fallback)")
    L = lib()
    sp = _stream_ptr(stream)
    if mode == "pair" and n <= 60:
        mode = "contiguous"  # tiny rules: the recurrence kernel writes plain index ranges
    if mode == "contiguous":
        lo, hi = index_partition(n, world)[rank]
        if out is None:
            x = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
            w = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
            out = LegendreShard(n, rank, world, mode, [(lo, hi, x, w)])
        _, _, x, w = out.segments[0]
        check(L.fgq_gausslegendre_device(n, lo, hi, x.data_ptr(), w.data_ptr(), sp))
        return out
    if mode != "pair":
        raise ValueError("mode must be 'pair' or 'contiguous'")
    k_lo, k_hi = half_index_partition(n, world)[rank]
    cnt = k_hi - k_lo
    m, mr = (n + 1) // 2, n // 2
    has_centre = (n % 2 == 1) and (k_hi - 1 == m) and cnt > 0
    if out is None:
        pad = cnt + (cnt & 1)  # keep the right segment 16-byte aligned inside one allocation
        xb = torch.empty(2 * pad, dtype=torch.float64, device="cuda")
        wb = torch.empty(2 * pad, dtype=torch.float64, device="cuda")
        xl, wl = xb[:cnt], wb[:cnt]
        xr_full, wr_full = xb[pad:pad + cnt], wb[pad:pad + cnt]
        skip = 1 if has_centre else 0
        segs = [(k_lo - 1, k_hi - 1, xl, wl),
                (n - k_hi + 1 + skip, n - k_lo + 1, xr_full[skip:], wr_full[skip:])]
        out = LegendreShard(n, rank, world, mode, segs)
        out._right_full = (xr_full, wr_full)
    xl, wl = out.segments[0][2], out.segments[0][3]
    xr_full, wr_full = out._right_full
    if cnt > 0:
        check(L.fgq_gausslegendre_device_pair(n, k_lo, k_hi, xl.data_ptr(), wl.data_ptr(),
                                              xr_full.data_ptr(), wr_full.data_ptr(), sp))
    return out


# --------------------------------------------------------------------------------------------
# Gauss-Jacobi, Radau, Lobatto  (reference src/gaussjacobi.jl:23-55, gaussradau.jl, gausslobatto.jl)
# --------------------------------------------------------------------------------------------

def gaussjacobi(n, a, b):
    """``gaussjacobi(n, α, β) -> x, w`` (reference src/gaussjacobi.jl:23).

    Integer α, β are promoted to Float64 (regression #117).  Errors as in the reference: n < 0 or
    min(α,β) <= -1 raise DomainError (:26-27, :44-45).  The device path covers the asymptotic
    branch (n > 100), the recurrence (n <= 100), the Golub-Welsch region (max(α,β) >= 5, n <= 4000),
    (0,0) -> Legendre and (±1/2, ±1/2) -> Chebyshev; max(α,β) >= 5 with n > 4000 raises, as in the reference."""
    n = _as_int(n)
    a = float(a)
    b = float(b)
    if n < 0:
        raise DomainError(f"DomainError({n}): gaussjacobi({n},{a},{b}) not defined: n must be non-negative.")
    if a == 0.0 and b == 0.0:
        return gausslegendre(n)
    if abs(a) == 0.5 and abs(b) == 0.5:      # gaussjacobi.jl:30-39
        kind = 1 if (a < 0 and b < 0) else 2 if (a > 0 and b > 0) else 3 if a < 0 else 4
        return _chebyshev(kind, n)
    x, w = _host_pair(n, False)
    if n == 0:
        return x, w
    check(lib().fgq_gaussjacobi_host(n, a, b, _ptr(x), _ptr(w)))
    return x, w


def gaussjacobi_device(n, a, b, stream=None):
    """Device-resident ``gaussjacobi``: (x, w, workspace) as torch CUDA tensors."""
    import torch
    n = _as_int(n)
    L = lib()
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(max(n, 0))) + 16, dtype=torch.uint8, device="cuda")
    check(L.fgq_gaussjacobi_device(n, float(a), float(b), 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(),
                                   _stream_ptr(stream)))
    return x, w, ws


def gaussradau(n, a=None, b=None):
    """``gaussradau(n) -> x, w`` (reference src/gaussradau.jl:31-49) and ``gaussradau(n, α, β)``
    (:52-69, general Jacobi weight); x[0] == -1 exactly."""
    n = _as_int(n)
    if n <= 0:
        raise DomainError(f"DomainError({n}): Input n must be a positive integer")
    x, w = _host_pair(n, False)
    if a is None and b is None:
        check(lib().fgq_gaussradau_host(n, _ptr(x), _ptr(w)))
    else:
        check(lib().fgq_gaussradau_jacobi_host(n, float(a), float(b), _ptr(x), _ptr(w)))
    return x, w


def gausslobatto(n):
    """``gausslobatto(n) -> x, w`` (reference src/gausslobatto.jl:31-52); x[0] == -1, x[-1] == 1."""
    n = _as_int(n)
    if n <= 1:
        raise DomainError(f"DomainError({n}): Lobatto undefined for n <= 1.")
    x, w = _host_pair(n, False)
    check(lib().fgq_gausslobatto_host(n, _ptr(x), _ptr(w)))
    return x, w


def _chebyshev(kind: int, n):
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    x, w = _host_pair(n, False)
    if n:
        check(lib().fgq_gausschebyshev_host(kind, n, _ptr(x), _ptr(w)))
    return x, w


def gausschebyshevt(n):
    """``gausschebyshevt(n)`` (reference src/gausschebyshev.jl:22-26)."""
    return _chebyshev(1, n)


def gausschebyshevu(n):
    """``gausschebyshevu(n)`` (reference src/gausschebyshev.jl:51-55)."""
    return _chebyshev(2, n)


def gausschebyshevv(n):
    """``gausschebyshevv(n)`` (reference src/gausschebyshev.jl:80-84)."""
    return _chebyshev(3, n)


def gausschebyshevw(n):
    """``gausschebyshevw(n)`` (reference src/gausschebyshev.jl:109-113)."""
    return _chebyshev(4, n)


def gausschebyshev(n, kind=1):
    """Deprecated ``gausschebyshev(n, kind=1)`` (reference src/gausschebyshev.jl:117-144): kind 1..4 forwards to
    gausschebyshevt/u/v/w with a DeprecationWarning (the reference's ``Base.depwarn``); any other kind raises
    ``ValueError`` (the reference's ``ArgumentError``); n < 0 is a DomainError first, as there."""
    import warnings
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    if kind not in (1, 2, 3, 4):
        raise ValueError("Chebyshev kind should be 1, 2, 3, or 4")
    name = "tuvw"[kind - 1]
    warnings.warn(f"`gausschebyshev(n, {kind})` is deprecated and will be removed in the next breaking release. "
                  f"Please use `gausschebyshev{name}(n)` instead.", DeprecationWarning, stacklevel=2)
    return _chebyshev(int(kind), n)


def besselroots(nu, n):
    """Deprecated alias of ``approx_besselroots`` (reference src/besselroots.jl:247)."""
    import warnings
    warnings.warn("`besselroots(nu, n)` is deprecated, use `approx_besselroots(nu, n)` instead.", DeprecationWarning,
                  stacklevel=2)
    return approx_besselroots(nu, n)


def approx_besselroots(nu, n):
    """``approx_besselroots(ν, n)`` (reference src/besselroots.jl:23-67)."""
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    x = np.zeros(n, dtype=np.float64)
    check(lib().fgq_approx_besselroots(float(nu), n, _ptr(x)))
    return x


# --------------------------------------------------------------------------------------------
# Gauss-Hermite  (reference src/gausshermite.jl:28-33)
# --------------------------------------------------------------------------------------------

def gausshermite(n, *, normalize: bool = False):
    """``gausshermite(n; normalize=false) -> x, w`` (reference src/gausshermite.jl:28,33).

    n < 0 raises DomainError (:38); n == 0 returns empty arrays (:40).  All n are served on the
    device: n <= 20 by a tiny eigen-solve (the reference's Golub-Welsch, :43-45), n <= 200 by Newton
    on the recurrence (:46-48), larger n by the Airy-type asymptotics (:49-50)."""
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    x, w = _host_pair(n, False)
    if n == 0:
        return x, w
    check(lib().fgq_gausshermite_host(n, int(bool(normalize)), _ptr(x), _ptr(w)))
    return x, w


def gausshermite_reduced(n, *, normalize: bool = False):
    """Truncated Gauss-Hermite rule: ``(first, x, w)`` covering positions ``first : first+len(x)`` (0-based) of
    ``gausshermite(n)`` — every weight >= 10 floatmin; all omitted weights are 0.0 in the full rule (exp(-x^2)
    underflows, gausshermite.jl:30).  Only that block is computed (2.4 % of the nodes at n = 1e6), with the same
    kernels, Newton stopping rule and normalisation sum taken over the block: the entries agree with the full rule
    to the Newton tolerance (a few ulp of sqrt(2n+1); the global stopping rule sees fewer nodes and may stop one
    sweep apart) and are bit-identical only when nothing is truncated (n <= 200).  The analogue of
    ``gausslaguerre(n; reduced=true)`` (gausslaguerre.jl:163-169)."""
    import ctypes
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    L = lib()
    bound = int(L.fgq_gausshermite_reduced_bound(n))
    x, w = _host_pair(bound, False)
    first, n_out = ctypes.c_int64(0), ctypes.c_int64(0)
    if n > 0:
        check(L.fgq_gausshermite_reduced_host(n, int(bool(normalize)), _ptr(x), _ptr(w), ctypes.byref(first),
                                              ctypes.byref(n_out)))
    return int(first.value), x[:n_out.value], w[:n_out.value]


def gausshermite_reduced_device(n, *, normalize: bool = False, stream=None):
    """Device-resident truncated rule: ``(first, x, w, workspace)`` with x, w torch CUDA tensors holding the
    entries ``first : first+len(x)`` of the full rule (the block inside |x| <= 27.5, untrimmed)."""
    import ctypes
    import torch
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    L = lib()
    bound = int(L.fgq_gausshermite_reduced_bound(n))
    x = torch.empty(bound, dtype=torch.float64, device="cuda")
    w = torch.empty(bound, dtype=torch.float64, device="cuda")
    ws = torch.empty(max(int(L.fgq_gausshermite_workspace_bytes(n)), 8), dtype=torch.uint8, device="cuda")
    first, count = ctypes.c_int64(0), ctypes.c_int64(0)
    check(L.fgq_gausshermite_reduced_device(n, int(bool(normalize)), x.data_ptr(), w.data_ptr(), ws.data_ptr(),
                                            _stream_ptr(stream), ctypes.byref(first), ctypes.byref(count)))
    return int(first.value), x[:count.value], w[:count.value], ws


def gausshermite_device(n, *, normalize: bool = False, stream=None):
    """Device-resident ``gausshermite(n)``: returns (x, w, info) with x, w torch CUDA tensors."""
    import torch
    n = _as_int(n)
    if n < 0:
        raise DomainError(f"DomainError({n}): Input n must be a non-negative integer")
    L = lib()
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    check(L.fgq_gausshermite_device(n, int(bool(normalize)), 0, n, x.data_ptr(), w.data_ptr(),
                                    ws.data_ptr(), _stream_ptr(stream)))
    return x, w, ws


# --------------------------------------------------------------------------------------------
# Gauss-Laguerre  (reference src/gausslaguerre.jl:1,23,59-96)
# --------------------------------------------------------------------------------------------

def gausslaguerre(n, alpha=0.0, *, reduced: bool = False):
    """``gausslaguerre(n[, α]; reduced=false) -> x, w`` (reference src/gausslaguerre.jl:1,23,59).

    α <= -1 or n < 0 raise DomainError (:60-65); n == 0 returns empty arrays (:69).  n >= 128 with
    α < 5 is the explicit-asymptotics branch; n < 128 (closed forms, Golub-Welsch, recurrence rule)
    and α >= 5 (recurrence rule, up to n = 4096) are served by a single-thread kernel, as the work is
    sequential there.  ``reduced=True`` drops the tail of underflowed weights."""
    n = _as_int(n)
    alpha = float(alpha)
    if alpha <= -1:
        raise DomainError(f"DomainError({alpha}): The Laguerre parameter α ≤ -1 corresponds to a nonintegrable weight function")
    if n < 0:
        raise DomainError(f"DomainError({n}): gausslaguerre({n},{alpha}) not defined: n must be positive.")
    x, w = _host_pair(n, False)
    if n == 0:
        return x, w
    n_out = ctypes.c_int64(0)
    check(lib().fgq_gausslaguerre_host(n, alpha, int(bool(reduced)), _ptr(x), _ptr(w), ctypes.byref(n_out)))
    if reduced:
        return x[:n_out.value].copy(), w[:n_out.value].copy()
    return x, w


def gausslaguerre_device(n, alpha=0.0, stream=None):
    """Device-resident ``gausslaguerre``: (x, w, workspace) as torch CUDA tensors."""
    import torch
    n = _as_int(n)
    L = lib()
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(max(n, 0))) + 16, dtype=torch.uint8, device="cuda")
    check(L.fgq_gausslaguerre_device(n, float(alpha), x.data_ptr(), w.data_ptr(), ws.data_ptr(), _stream_ptr(stream)))
    return x, w, ws


def rule_moments_device(shard: LegendreShard, stream=None):
    """(sum w, sum w x^2) of this rank's shard as a 2-element CUDA tensor (a fused-reduction
    consumer; callers all-reduce it across ranks for the Σw = 2, ∫x² = 2/3 check)."""
    import torch
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    sp = _stream_ptr(stream)
    for lo, hi, x, w in shard.segments:
        if hi > lo:
            check(lib().fgq_rule_moments_device(x.data_ptr(), w.data_ptr(), hi - lo, sums.data_ptr(), sp))
    return sums


def gausslegendre_moments_fused(n, rank: int = 0, world: int = 1, with_exp: bool = False, stream=None):
    """(sum w, sum w x^2[, sum w exp(x)]) of ``gausslegendre(n)`` for this rank's half-index shard,
    generated and reduced in one kernel — nothing is stored (consumer fusion).
    Returns a 3-element CUDA tensor; all-reduce it across ranks."""
    import torch
    n = _as_int(n)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    k_lo, k_hi = half_index_partition(n, world)[rank]
    check(lib().fgq_gausslegendre_moments_fused_device(n, k_lo, k_hi, int(bool(with_exp)), sums.data_ptr(),
                                                       _stream_ptr(stream)))
    return sums


def _check_integrand(f: str) -> bytes:
    """Syntax screening of a user integrand by the library (no device work), then the no-CPU-fallback rule."""
    if not isinstance(f, str):
        raise TypeError("the integrand must be a string: an expression in x")
    dummy = (ctypes.c_double * 2)()
    check(lib().fgq_rule_integrate_device(None, None, 0, f.encode(), dummy, None))
    if lib().fgq_device_count() == 0:
        raise _lib.FGQError("user integrands need a CUDA device (no CPU fallback)")
    return f.encode()


def gausslegendre_integrate(n, f: str, rank: int = 0, world: int = 1, stream=None):
    """``dot(w, f.(x))`` over ``gausslegendre(n)`` (n > 60) for a USER integrand, generated and reduced in one
    kernel — x and w are never stored (README.md:27-30 is exactly this reduction).

    ``f`` is an expression in ``x`` in CUDA C++ syntax, e.g. ``"exp(x) * cos(3 * x)"`` or ``"1 / (1 + 25 * x * x)"``;
    it is compiled once per (expression, series branch) with NVRTC together with the library's own node-generation
    source and cached.  Returns a 2-element CUDA tensor ``[sum w f(x), sum w]`` for this rank's half-index shard;
    all-reduce it across ranks."""
    import torch
    n = _as_int(n)
    fb = _check_integrand(f)
    if n <= 60:
        raise ValueError(f"fused integration needs the asymptotic path (n > 60), got n={n}; use rule_integrate on a stored rule")
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    k_lo, k_hi = half_index_partition(n, world)[rank]
    check(lib().fgq_gausslegendre_integrate_device(n, k_lo, k_hi, fb, sums.data_ptr(), _stream_ptr(stream)))
    return sums


def rule_integrate(x, w, f: str, stream=None):
    """``dot(w, f.(x))`` over a STORED rule (torch CUDA tensors of any family: Jacobi, Hermite, Laguerre, ...) for a
    user integrand ``f`` (expression in ``x``, as in :func:`gausslegendre_integrate`).  Returns
    ``[sum w f(x), sum w]`` as a CUDA tensor."""
    import torch
    if x.numel() != w.numel() or x.dtype != torch.float64 or w.dtype != torch.float64 or not x.is_cuda or not w.is_cuda:
        raise ValueError("x and w must be float64 CUDA tensors of equal length")
    fb = _check_integrand(f)
    x, w = x.contiguous(), w.contiguous()
    sums = torch.zeros(2, dtype=torch.float64, device=x.device)
    check(lib().fgq_rule_integrate_device(x.data_ptr(), w.data_ptr(), x.numel(), fb, sums.data_ptr(),
                                          _stream_ptr(stream)))
    return sums


def measure_peaks(store_bytes: int = 4 << 30):
    """Roofline denominators on the current device: (FP64 TFLOP/s from the DFMA microbenchmark,
    store-only HBM GB/s)."""
    tf = ctypes.c_double(0.0)
    gb = ctypes.c_double(0.0)
    check(lib().fgq_measure_dfma_peak(ctypes.byref(tf)))
    check(lib().fgq_measure_store_peak(int(store_bytes), ctypes.byref(gb)))
    return tf.value, gb.value
