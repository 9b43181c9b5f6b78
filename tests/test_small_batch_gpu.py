"""Batched small rules of the Jacobi / Hermite / Laguerre families:
`fgq_gauss{jacobi,hermite,laguerre}_batch_*`.  The single-rule calls are parity-tested against the oracles
in the per-family files; here a batch must be BIT-IDENTICAL to the concatenation of single calls (it runs
the same device functions), plus a direct oracle check on a few rules and the error behaviour."""
import numpy as np
import pytest

from oracle import hermite_oracle as ho
from oracle import jacobi_oracle as jo
from oracle import laguerre_oracle as lo

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _assert_batch_equals_singles(offsets, x, w, singles):
    assert len(offsets) == len(singles) + 1
    for r, (xs, ws) in enumerate(singles):
        lo_, hi_ = offsets[r], offsets[r + 1]
        assert hi_ - lo_ == len(xs), r
        assert np.array_equal(x[lo_:hi_], xs), r
        assert np.array_equal(w[lo_:hi_], ws), r


def test_jacobi_batch_equals_single_calls(fgq):
    rng = np.random.default_rng(21)
    n_i = np.concatenate([np.arange(0, 101), rng.integers(0, 101, 200)]).astype(np.int32)
    a = rng.uniform(-0.95, 4.9, len(n_i))
    b = rng.uniform(-0.95, 4.9, len(n_i))
    # the special routes of gaussjacobi.jl:28-39,50-51 inside a batch: Legendre (rec and asy), the four
    # Chebyshev kinds, the Golub-Welsch region, integer parameters, alpha == beta
    a[5], b[5] = 0.0, 0.0
    a[70], b[70] = 0.0, 0.0
    a[33], b[33] = -0.5, -0.5
    a[34], b[34] = 0.5, 0.5
    a[35], b[35] = -0.5, 0.5
    a[36], b[36] = 0.5, -0.5
    a[40], b[40] = 19.0, 21.0
    a[41], b[41] = 5.0, -0.3
    a[50], b[50] = 1.0, 2.0
    a[60:66] = b[60:66]
    off, x, w = fgq.gaussjacobi_batch(n_i, a, b)
    singles = [fgq.gaussjacobi(int(n), float(ai), float(bi)) for n, ai, bi in zip(n_i, a, b)]
    _assert_batch_equals_singles(off, x, w, singles)
    # and straight against the oracle on a few rec rules
    for r in (12, 57, 99, 150):
        n, ai, bi = int(n_i[r]), float(a[r]), float(b[r])
        if n < 2:
            continue
        xo, wo = jo.gaussjacobi(n, ai, bi)
        assert np.abs(x[off[r]:off[r + 1]] - xo).max() <= 4 * EPS
        assert np.abs(w[off[r]:off[r + 1]] / wo - 1).max() < 1e-11


def test_jacobi_batch_scalar_parameters_and_gaps(fgq):
    """Scalar alpha/beta broadcast; caller-supplied offsets with gaps leave the gaps untouched."""
    n_i = np.array([7, 0, 100, 1, 42], dtype=np.int32)
    offsets = np.array([3, 20, 20, 130, 140, 190], dtype=np.int64)  # rule r at offsets[r]; last entry = length
    off, x, w = fgq.gaussjacobi_batch(n_i, -0.1, 0.3, offsets=offsets)
    for r, n in enumerate(n_i):
        xs, ws = fgq.gaussjacobi(int(n), -0.1, 0.3)
        assert np.array_equal(x[off[r]:off[r] + n], xs) and np.array_equal(w[off[r]:off[r] + n], ws)
    # the reference golden gaussjacobi(42, -0.1, 0.3) (test_gaussjacobi.jl:149-159)
    assert abs(x[140 + 36] - 0.912883347814032) < 1e-14 and abs(w[140 + 36] - 0.046661910947553) < 1e-14


def test_hermite_batch_equals_single_calls(fgq):
    rng = np.random.default_rng(22)
    n_i = np.concatenate([np.arange(0, 201), rng.integers(0, 201, 100)]).astype(np.int32)
    for normalize in (False, True):
        off, x, w = fgq.gausshermite_batch(n_i, normalize=normalize)
        singles = [fgq.gausshermite(int(n), normalize=normalize) for n in n_i]
        _assert_batch_equals_singles(off, x, w, singles)
    off, x, w = fgq.gausshermite_batch(n_i)
    for r in (42, 21, 200, 137):
        xo, wo = ho.gausshermite(int(n_i[r]))
        assert np.abs(x[off[r]:off[r + 1]] - xo).max() <= 8 * np.spacing(np.sqrt(2.0 * n_i[r] + 1))
        assert abs(w[off[r]:off[r + 1]].sum() - np.sqrt(np.pi)) < 1e-13


def test_laguerre_batch_equals_single_calls(fgq):
    rng = np.random.default_rng(23)
    n_i = np.concatenate([np.arange(0, 128), rng.integers(0, 128, 100)]).astype(np.int32)
    alpha = rng.uniform(-0.95, 4.9, len(n_i))
    alpha[::7] = 0.0
    alpha[3] = 7.5      # alpha >= 5: the recurrence rule for every n (gausslaguerre.jl:93-94)
    alpha[100] = 12.0
    off, x, w = fgq.gausslaguerre_batch(n_i, alpha)
    singles = [fgq.gausslaguerre(int(n), float(al)) for n, al in zip(n_i, alpha)]
    _assert_batch_equals_singles(off, x, w, singles)
    # scalar alpha, and the reference golden gausslaguerre(42) (test_gausslaguerre.jl)
    off, x, w = fgq.gausslaguerre_batch([42, 5, 100], 0.0)
    for r, n in enumerate((42, 5, 100)):
        xo, wo = lo.gausslaguerre(n, 0.0)
        xs = x[off[r]:off[r + 1]]
        ws = w[off[r]:off[r + 1]]
        assert (np.abs(xs - xo) / xo).max() < 1e-12
        big = wo > 1e-290
        assert np.abs(ws[big] / wo[big] - 1).max() < 1e-9  # the recurrence rule itself is only ~1e-12 accurate
        assert abs(ws.sum() - 1.0) < 2e-12


def test_batch_errors(fgq):
    with pytest.raises(fgq.DomainError):
        fgq.gaussjacobi_batch([5, -1], 0.3, 0.2, offsets=np.array([0, 5, 5], dtype=np.int64))
    with pytest.raises(fgq.DomainError):
        fgq.gaussjacobi_batch([5, 6], [0.3, -1.0], 0.2)
    with pytest.raises(ValueError):
        fgq.gaussjacobi_batch([5, 101], 0.3, 0.2)
    with pytest.raises(ValueError):
        fgq.gausshermite_batch([5, 201])
    with pytest.raises(fgq.DomainError):
        fgq.gausshermite_batch([5, -2], offsets=np.array([0, 5, 5], dtype=np.int64))
    with pytest.raises(ValueError):
        fgq.gausslaguerre_batch([5, 128], 0.0)
    with pytest.raises(fgq.DomainError):
        fgq.gausslaguerre_batch([5, 6], [-1.0, 0.0])
    # empty batches
    for off, x, w in (fgq.gaussjacobi_batch([], 0.1, 0.2), fgq.gausshermite_batch([]), fgq.gausslaguerre_batch([])):
        assert len(x) == 0 and len(w) == 0
    off, x, w = fgq.gausshermite_batch([0, 0])
    assert len(x) == 0


def test_batch_device_variant_on_user_stream(fgq):
    """The device entry points fill caller-owned device buffers on the caller's stream."""
    import torch
    L = fgq.lib()
    n_i = np.array([30, 64, 9, 100], dtype=np.int32)
    a = np.array([0.3, -0.2, 1.5, 2.0])
    b = np.array([-0.7, 0.4, 0.0, 2.0])
    offsets = np.zeros(5, dtype=np.int64)
    np.cumsum(n_i, out=offsets[1:])
    total = int(offsets[-1])
    out = torch.full((2, total), float("nan"), dtype=torch.float64, device="cuda")
    ws = torch.empty(L.fgq_gaussjacobi_batch_workspace_bytes(len(n_i)), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        rc = L.fgq_gaussjacobi_batch_device(len(n_i), n_i.ctypes.data, a.ctypes.data, b.ctypes.data,
                                            offsets.ctypes.data, out[0].data_ptr(), out[1].data_ptr(),
                                            ws.data_ptr(), st.cuda_stream)
    assert rc == 0
    st.synchronize()
    x, w = out[0].cpu().numpy(), out[1].cpu().numpy()
    for r in range(len(n_i)):
        xs, wsgl = fgq.gaussjacobi(int(n_i[r]), float(a[r]), float(b[r]))
        assert np.array_equal(x[offsets[r]:offsets[r + 1]], xs) and np.array_equal(w[offsets[r]:offsets[r + 1]], wsgl)
