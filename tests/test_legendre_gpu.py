"""GPU parity tests for the Gauss-Legendre path, through the C ABI (ctypes) of libfgq_cuda.so.

Bar (BASELINE.json north_star): nodes within 2 ulp, weights within 4 ulp of the reference's
Float64 results; theta bit-exact; zero entries exactly zero.  Oracle modes:
  julia — the C restatement of the path with Julia Base's sin/cos/tan restated operation by
          operation (oracle/julia_libm.h): the functions the reference really calls.  The kernels
          follow the same association, so the expectation is BIT-FOR-BIT (0 ulp <= 2 / 4 ulp);
  twin  — the library's own fgq_math.h compiled for the host: bit-for-bit as well (and equal to
          the julia mode: tests/test_oracle_legendre.py::test_twin_equals_julia_mode);
  libm  — glibc sin/cos/tan, an unrelated < 1 ulp libm: sanity band for the nodes only.  Two
          different correct libms put a 1-ulp difference of sin(a) through the four roundings of
          w = 2/(J1^2 (a/sin a) W), so glibc and Julia weights differ by up to 5 ulp from EACH
          OTHER (tests/test_oracle_legendre.py::test_glibc_vs_julia_band); that is a property of
          the two oracles, not of the kernel, and is not the parity bar.
"""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "legendre_reference_tests.json")))

X_ULP, W_ULP = 2, 4


def _ulps(oracle, got, ref):
    nz = ref != 0
    assert np.array_equal(got[~nz], ref[~nz]), "entries that are exactly 0.0 in the reference must be 0.0"
    return oracle.ulp_diff(got[nz], ref[nz]).max() if nz.any() else 0.0


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}")
def test_reference_golden_values(fgq, case):
    x, w = fgq.gausslegendre(case["n"])
    for i, v in case["x"].items():
        assert abs(x[int(i) - 1] - v) < GOLD["tol"]
    for i, v in case["w"].items():
        assert abs(w[int(i) - 1] - v) < GOLD["tol"]
    assert abs(np.dot(w, x * x) - 2 / 3) < 1e-14
    assert abs(np.dot(w, np.exp(x)) - (math.e - 1 / math.e)) < 1e-14


def test_special_cases(fgq):
    # test_gausslegendre.jl:4-14
    with pytest.raises(fgq.DomainError):
        fgq.gausslegendre(-1)
    x, w = fgq.gausslegendre(0)
    assert x.shape == (0,) and w.shape == (0,)
    for n in range(0, 11):
        x, w = fgq.gausslegendre(n)
        assert x.dtype == np.float64 and x.shape == (n,)
        for i in range(n):
            r = 2 * i
            assert abs(np.dot(w, x ** r) - 2 / (r + 1)) < 10 * np.finfo(float).eps


@pytest.mark.parametrize("n", list(range(1, 62)))
def test_small_rules_bit_exact_vs_twin(fgq, oracle_mod, n):
    """Closed forms (n<=5) and the Newton/recurrence path (n<=60): bit-for-bit against the twin
    oracle; against the libm oracle the nodes stay within 2 ulp."""
    x, w = fgq.gausslegendre(n)
    xt, wt = oracle_mod.gausslegendre(n, twin=True)
    assert np.array_equal(x, xt) and np.array_equal(w, wt)
    xj, wj = oracle_mod.gausslegendre(n, twin="julia")
    assert np.array_equal(x, xj) and np.array_equal(w, wj)
    xo, wo = oracle_mod.gausslegendre(n)
    assert _ulps(oracle_mod, x, xo) <= X_ULP


ASY_N = [61, 100, 170, 171, 255, 256, 1013, 1500, 1501, 3950, 3951, 10013, 13191 * 2, 13192 * 2 + 1,
         100000, 100001, 850000, 850001, 1000003, (1 << 25) - 1, 1 << 25, (1 << 25) + 1]


@pytest.mark.parametrize("n", ASY_N)
def test_asy_parity(fgq, oracle_mod, n):
    x, w = fgq.gausslegendre(n)
    xt, wt = oracle_mod.gausslegendre(n, twin=True)
    assert np.array_equal(x, xt), "nodes must match the twin oracle bit-for-bit"
    assert np.array_equal(w, wt), "weights must match the twin oracle bit-for-bit"
    # the reference's own elementary functions (restated Julia Base libm): every bit
    xj, wj = oracle_mod.gausslegendre(n, twin="julia")
    assert np.array_equal(x, xj), "nodes must match the julia-mode oracle bit-for-bit"
    assert np.array_equal(w, wj), "weights must match the julia-mode oracle bit-for-bit"
    assert _ulps(oracle_mod, x, xj) <= X_ULP and _ulps(oracle_mod, w, wj) <= W_ULP
    # an unrelated libm (glibc): nodes within the bar as well
    xo, wo = oracle_mod.gausslegendre(n)
    assert _ulps(oracle_mod, x, xo) <= X_ULP
    if n % 2:
        assert x[n // 2] == 0.0
    assert np.all(np.diff(x) >= 0)


@pytest.mark.parametrize("n", [100, 1013, 10013, 100000, 850001, (1 << 25) - 1, 1 << 25, 10 ** 9])
def test_theta_bit_exact(fgq, oracle_mod, n):
    """The angle theta_k is reproduced bit-for-bit (SURVEY.md §7 hard part 1).  For n >= 2^25 the
    kernel returns a_k itself and the oracle evaluates the literal a + correction."""
    import torch
    m = (n + 1) // 2
    ranges = [(1, min(m, 40000) + 1)]
    if m > 400000:
        ranges += [(m // 2, m // 2 + 100000), (m - 100000, m + 1)]
    for k_lo, k_hi in ranges:
        th = torch.empty(k_hi - k_lo, dtype=torch.float64, device="cuda")
        fgq._lib.check(fgq.lib().fgq_debug_legendre_theta_device(n, k_lo, k_hi, th.data_ptr(), 0))
        torch.cuda.synchronize()
        got = th.cpu().numpy()
        assert np.array_equal(got, oracle_mod.legendre_theta(n, k_lo, k_hi - 1, twin=True))
        # cot = inv(tan) with Julia Base's tan (gausslegendre.jl:100): every theta, every bit
        assert np.array_equal(got, oracle_mod.legendre_theta(n, k_lo, k_hi - 1, twin="julia"))
        # an unrelated libm (glibc tan): within one ulp of theta
        ref = oracle_mod.legendre_theta(n, k_lo, k_hi - 1)
        assert np.abs(got - ref).max() <= np.spacing(ref).max()


def _device_rule(fgq, n, world, mode):
    import torch
    x = np.full(n, np.nan)
    w = np.full(n, np.nan)
    for r in range(world):
        sh = fgq.gausslegendre_device(n, rank=r, world=world, mode=mode)
        torch.cuda.synchronize()
        for lo, hi, xs, ws in sh.segments:
            assert np.all(np.isnan(x[lo:hi])), "segments of different ranks must not overlap"
            x[lo:hi] = xs.cpu().numpy()
            w[lo:hi] = ws.cpu().numpy()
    return x, w


@pytest.mark.parametrize("mode", ["pair", "contiguous"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [40, 61, 1013, 100000, 100001, 26383, 1000003])
def test_sharded_equals_single(fgq, n, world, mode):
    """Multi-GPU layout check on one device: the shards of every rank, concatenated, are
    bit-identical to the single-call rule (no halo, no exchange: SURVEY.md §8e)."""
    x1, w1 = fgq.gausslegendre(n)
    x, w = _device_rule(fgq, n, world, mode)
    assert np.array_equal(x, x1) and np.array_equal(w, w1)


def test_arbitrary_slices(fgq):
    """Odd/unaligned [lo, hi) ranges through the raw C ABI."""
    import torch
    rng = np.random.default_rng(1)
    for n in (1013, 26385, 100000):
        x1, w1 = fgq.gausslegendre(n)
        for _ in range(12):
            lo = int(rng.integers(0, n))
            hi = int(rng.integers(lo, n + 1))
            xs = torch.full((hi - lo + 3,), -7.0, dtype=torch.float64, device="cuda")
            ws = torch.full((hi - lo + 3,), -7.0, dtype=torch.float64, device="cuda")
            off = int(rng.integers(0, 2))  # also misalign the base pointer by 8 bytes
            fgq._lib.check(fgq.lib().fgq_gausslegendre_device(
                n, lo, hi, xs.data_ptr() + 8 * off, ws.data_ptr() + 8 * off, 0))
            torch.cuda.synchronize()
            xh, wh = xs.cpu().numpy(), ws.cpu().numpy()
            assert np.array_equal(xh[off:off + hi - lo], x1[lo:hi])
            assert np.array_equal(wh[off:off + hi - lo], w1[lo:hi])
            # nothing outside the slice is touched
            assert np.all(xh[:off] == -7.0) and np.all(xh[off + hi - lo:] == -7.0)
            assert np.all(wh[:off] == -7.0) and np.all(wh[off + hi - lo:] == -7.0)


def test_invalid_arguments(fgq):
    L = fgq.lib()
    assert L.fgq_gausslegendre_device(-1, 0, 0, 0, 0, 0) == 1
    assert b"DomainError" in L.fgq_last_error()
    assert L.fgq_gausslegendre_device(100, 5, 3, 0, 0, 0) == 3
    assert L.fgq_gausslegendre_device(100, 0, 101, 0, 0, 0) == 3
    assert L.fgq_gausslegendre_device(100, 0, 10, 0, 0, 0) == 3  # null pointers
    assert L.fgq_gausslegendre_device_pair(40, 1, 5, 0, 0, 0, 0, 0) == 3
    n_i = np.array([5, 61], dtype=np.int32)
    off = np.array([0, 5], dtype=np.int64)
    buf = np.empty(70)
    assert L.fgq_gausslegendre_batch_host(2, n_i.ctypes.data, off.ctypes.data, buf.ctypes.data, buf.ctypes.data) == 3
    n_i[1] = -2
    assert L.fgq_gausslegendre_batch_host(2, n_i.ctypes.data, off.ctypes.data, buf.ctypes.data, buf.ctypes.data) == 1


def test_host_pipeline_multi_chunk(fgq, oracle_mod):
    """n larger than the D2H pipeline chunk (2^24 nodes): buffers are recycled; pinned and
    pageable destinations give the same bits."""
    n = 3 * (1 << 24) + 7
    x, w = fgq.gausslegendre(n)
    xp, wp = fgq.gausslegendre(n, pinned=True)
    assert np.array_equal(x, xp) and np.array_equal(w, wp)
    xt, wt = oracle_mod.gausslegendre(n, twin=True)
    assert np.array_equal(x, xt) and np.array_equal(w, wt)


def test_library_pinned_destinations(fgq):
    """fgq_host_alloc / fgq_host_free: the library's page-locked destinations (what `pinned=True` and the Julia
    wrapper's `pinned_vector` hand out) are recognised as pinned by the runtime, reusable across calls via `out=`,
    and released with their last view."""
    import gc
    import torch
    n = (1 << 21) + 5
    x, w = fgq.empty_pinned(n), fgq.empty_pinned(n)
    assert x.dtype == np.float64 and x.shape == (n,) and x.flags.c_contiguous and x.flags.writeable
    assert torch.from_numpy(x).is_pinned() and torch.from_numpy(w).is_pinned()
    x[:] = 7.0
    fgq.gausslegendre(n, out=(x, w))
    xr, wr = fgq.gausslegendre(n)
    assert np.array_equal(x, xr) and np.array_equal(w, wr)
    fgq.gausslegendre(n, out=(x, w))     # reuse
    assert np.array_equal(x, xr)
    view = x[10:20]
    del x
    gc.collect()
    assert view[0] == xr[10]              # the block lives as long as a view does
    assert fgq.empty_pinned(0).shape == (0,)
    L = fgq.lib()
    assert L.fgq_host_free(None) == 0
    assert L.fgq_host_alloc(-1, None) == 3


def test_batch_small_rules(fgq, oracle_mod):
    """BASELINE config C5 (reduced count here; the full 1e6-rule batch is in test_full_size)."""
    rng = np.random.default_rng(0)
    n_i = rng.integers(2, 61, size=50000)
    n_i[:61] = np.arange(61)  # every n incl. 0 and 1
    off, x, w = fgq.gausslegendre_batch(n_i)
    off_o, xt, wt = oracle_mod.gausslegendre_batch(n_i, twin=True)
    assert np.array_equal(off, off_o)
    assert np.array_equal(x, xt) and np.array_equal(w, wt)
    _, xo, wo = oracle_mod.gausslegendre_batch(n_i)
    assert _ulps(oracle_mod, x, xo) <= X_ULP
    # weights of the recurrence path are ill-conditioned in the Newton iterate (SURVEY.md §7
    # hard part 2): against a different libm they are pinned in absolute terms only
    assert np.abs(w - wo).max() < 1e-14


def test_device_entry_points_are_graph_capturable(fgq):
    """include/fgq.h promises that the *_device entry points only enqueue work (no allocation, no
    synchronisation): capture them in a CUDA graph and replay."""
    import torch
    L = fgq.lib()
    n = 200001
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    w = torch.zeros(n, dtype=torch.float64, device="cuda")
    wsj = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    xj = torch.zeros(n, dtype=torch.float64, device="cuda")
    wj = torch.zeros(n, dtype=torch.float64, device="cuda")
    wsl = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    xl = torch.zeros(n, dtype=torch.float64, device="cuda")
    wl = torch.zeros(n, dtype=torch.float64, device="cuda")
    wsh = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    xh = torch.zeros(n, dtype=torch.float64, device="cuda")
    wh = torch.zeros(n, dtype=torch.float64, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            sp = torch.cuda.current_stream().cuda_stream
            fgq._lib.check(L.fgq_gausslegendre_device(n, 0, n, x.data_ptr(), w.data_ptr(), sp))
            fgq._lib.check(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, xj.data_ptr(), wj.data_ptr(), wsj.data_ptr(), sp))
            fgq._lib.check(L.fgq_gausslaguerre_device(n, 0.5, xl.data_ptr(), wl.data_ptr(), wsl.data_ptr(), sp))
            fgq._lib.check(L.fgq_gausshermite_device(n, 0, 0, n, xh.data_ptr(), wh.data_ptr(), wsh.data_ptr(), sp))
    for t in (x, w, xj, wj, xl, wl, xh, wh):
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    x1, w1 = fgq.gausslegendre(n)
    assert np.array_equal(x.cpu().numpy(), x1) and np.array_equal(w.cpu().numpy(), w1)
    x2, w2 = fgq.gaussjacobi(n, 0.3, -0.7)
    assert np.array_equal(xj.cpu().numpy(), x2) and np.array_equal(wj.cpu().numpy(), w2)
    x3, w3 = fgq.gausslaguerre(n, 0.5)
    assert np.array_equal(xl.cpu().numpy(), x3) and np.array_equal(wl.cpu().numpy(), w3)
    x4, w4 = fgq.gausshermite(n)
    assert np.array_equal(xh.cpu().numpy(), x4) and np.array_equal(wh.cpu().numpy(), w4)


@pytest.mark.parametrize("n", [(1 << 21) + 4, (1 << 21) + 5, 3 * (1 << 24) + 7, 2 * (1 << 25) + 10])
def test_host_symmetric_path(fgq, oracle_mod, n):
    """n >= 2^21: the host-returning call ships only the left half over PCIe and mirrors it on the host;
    the result must still be the twin oracle's rule bit for bit (incl. the centre of odd rules)."""
    x, w = fgq.gausslegendre(n)
    xt, wt = oracle_mod.gausslegendre(n, twin=True)
    assert np.array_equal(x, xt) and np.array_equal(w, wt)
    for world in (1, 3):
        xa = np.full(n, np.nan)
        wa = np.full(n, np.nan)
        for r in range(world):
            segs, _ = fgq.gausslegendre_host_pair(n, r, world, pinned=False)
            for lo, hi, xs, ws in segs:
                assert np.all(np.isnan(xa[lo:hi]))
                xa[lo:hi] = xs
                wa[lo:hi] = ws
        assert np.array_equal(xa, xt) and np.array_equal(wa, wt)


@pytest.mark.parametrize("share", ["0", "1", "0.37", None])
@pytest.mark.parametrize("n", [(1 << 21) + 5, 3 * (1 << 24) + 7, 3 * (1 << 24) + 8])
def test_host_symmetric_path_pinned(fgq, oracle_mod, n, share, monkeypatch):
    """Pinned destinations: the mirror image of a piece is written either by host threads or by a second
    pair of DMA copies of the image the kernel left in device memory; the split is adaptive at run time
    (share=None) and pinned here to 0, 1 and a mixed value: the bits must not depend on it."""
    if share is None:
        monkeypatch.delenv("FGQ_HOST_MIRROR_DMA_SHARE", raising=False)
    else:
        monkeypatch.setenv("FGQ_HOST_MIRROR_DMA_SHARE", share)
    xt, wt = oracle_mod.gausslegendre(n, twin=True)
    x, w = fgq.gausslegendre(n, pinned=True)
    assert np.array_equal(x, xt) and np.array_equal(w, wt)
    world = 3
    xa = np.full(n, np.nan)
    wa = np.full(n, np.nan)
    for r in range(world):
        segs, _ = fgq.gausslegendre_host_pair(n, r, world, pinned=True)
        for lo, hi, xs, ws in segs:
            assert np.all(np.isnan(xa[lo:hi]))
            xa[lo:hi] = xs
            wa[lo:hi] = ws
    assert np.array_equal(xa, xt) and np.array_equal(wa, wt)
