"""GPU parity tests for the Gauss-Hermite asymptotic path (C ABI via ctypes).

Parity metric (SURVEY.md §7 hard part 1, §8d): the node is x = sqrt(2n+1) cos(theta) with theta
fixed by a Newton iteration on an Airy-type expansion; one ulp of theta moves a node near x = 0 by
many ulps of x itself, and the oracle's Airy function (AMOS) carries ~zeta*1e-16 of phase noise.
Nodes are therefore compared in units of ulp(sqrt(2n+1)) (absolute scale of the rule) and weights
relatively, where they are not underflowed; exact zeros must match exactly."""
import json
import os

import numpy as np
import pytest

from oracle import hermite_oracle as ho

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hermite_reference_tests.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}")
def test_reference_golden_values(fgq, case):
    n = case["n"]
    x, w = fgq.gausshermite(n)
    assert x.shape == (n,) and w.shape == (n,)
    for i, (v, tol) in case["x"].items():
        assert abs(x[int(i) - 1] - v) < tol
    for i, (v, tol) in case["w"].items():
        assert abs(w[int(i) - 1] - v) < tol
    assert np.dot(w, x) < case["dot_wx_lt"]
    assert abs(np.dot(w, x * x) - np.sqrt(np.pi) / 2) < case["dot_wx2_tol"]
    if "poly_integral" in case:   # test_gausshermite.jl:35-36
        assert abs(np.dot(w, 1 + x + x ** 2 + x ** 3) - case["poly_integral"]) < 1.5e-8 * case["poly_integral"]  # Julia isapprox, rtol = sqrt(eps)


def test_all_small_n(fgq):
    """test_gausshermite.jl:97-106 ("All"): n = 2..220 spans Golub-Welsch (<= 20), recurrence (<= 200)
    and asymptotics; plus parity with the oracle (LAPACK eigen-solve / recurrence / AMOS)."""
    tol = 1e-14
    for n in range(2, 221):
        x, w = fgq.gausshermite(n)
        assert len(x) == n and len(w) == n
        assert np.dot(w, x) < tol and abs(np.dot(w, x ** 2) - np.sqrt(np.pi) / 2) < 1000 * tol
        xo, wo = ho.gausshermite(n)
        assert np.abs(x - xo).max() <= 8 * np.spacing(np.sqrt(2 * n + 1)), n
        nz = wo > 1e-290
        assert np.all(np.abs(w[nz] - wo[nz]) / wo[nz] <= 1e-13 + 4 * np.abs(xo[nz]) * 8 * np.spacing(np.sqrt(2 * n + 1))), n


def test_normalize_small_n(fgq):
    # test_gausshermite.jl:80-95
    for t in range(1, 7):
        x, w = fgq.gausshermite(t, normalize=True)
        v = 1.0
        for k in range(2, 2 * t, 2):
            v *= (k - 1)
            assert abs(v - np.dot(x ** k, w)) < 1e-13 * v
        for k in range(1, 2 * t, 2):
            assert abs(np.dot(x ** k, w)) < 1e-13


def test_api_edges(fgq):
    with pytest.raises(fgq.DomainError):      # test_gausshermite.jl:3
        fgq.gausshermite(-1)
    x, w = fgq.gausshermite(0)                 # :7
    assert len(x) == 0 and len(w) == 0
    x, w = fgq.gausshermite(1)                 # :8-10
    assert x[0] == 0.0 and abs(w[0] - np.sqrt(np.pi)) < 1e-15
    x, w = fgq.gausshermite(2)
    assert np.allclose(x, [-np.sqrt(0.5), np.sqrt(0.5)], rtol=4e-16) and np.allclose(w, [np.sqrt(np.pi) / 2] * 2, rtol=4e-16)


@pytest.mark.parametrize("n", [201, 251, 500, 1000, 1001, 4001, 100000, 100001, 1000000, 1000001])
def test_parity_with_oracle(fgq, n):
    x, w = fgq.gausshermite(n)
    xo, wo = ho.gausshermite(n)
    scale = np.sqrt(2 * n + 1)
    ex = np.abs(x - xo) / np.spacing(scale)
    print("n", n, "node err / ulp(scale): max", ex.max(), "mean", ex.mean())
    assert ex.max() <= 8, f"nodes: {ex.max()} ulp(sqrt(2n+1))"
    # weights: underflowed entries must be exactly zero in both; others agree relatively
    assert np.count_nonzero((w == 0) != (wo == 0)) <= 2
    nz = (wo > 1e-290) & (w > 1e-290)
    rel = np.abs(w[nz] - wo[nz]) / wo[nz]
    # the weight of node x carries exp(-x^2): a node difference dx changes it by 2 x dx relatively
    bound = 512 * np.finfo(float).eps + 2 * np.abs(xo[nz]) * 8 * np.spacing(scale)
    print("   weight rel err max", rel.max(), "worst/bound", np.max(rel / bound))
    assert np.all(rel <= bound), f"weights: worst {np.max(rel / bound)} x bound"
    assert np.all(np.diff(x) > 0)
    assert abs(w.sum() - np.sqrt(np.pi)) < 1e-13
    assert abs(np.dot(w, x * x) - np.sqrt(np.pi) / 2) < 2e-11


def test_sweep_count_matches_reference_rule(fgq):
    """The device-side global stopping rule must stop after the same number of sweeps as the
    reference's norm(dtheta, Inf) test (3 at n = 1e6)."""
    import ctypes
    import torch
    for n in (251, 1000000):
        info = {}
        ho.gausshermite(n, info=info)
        x, w, ws = fgq.gausshermite_device(n)
        torch.cuda.synchronize()
        s = ctypes.c_int(0)
        fgq._lib.check(fgq.lib().fgq_gausshermite_sweeps(ws.data_ptr(), ctypes.byref(s)))
        assert s.value == info["sweeps"]


def test_normalize_and_slices(fgq):
    import torch
    n = 5001
    x, w = fgq.gausshermite(n)
    xn, wn = fgq.gausshermite(n, normalize=True)
    assert np.allclose(xn, np.sqrt(2.0) * x, rtol=1e-15, atol=0) and np.allclose(wn, w / np.sqrt(np.pi), rtol=4e-16, atol=0)
    L = fgq.lib()
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    for lo, hi in ((0, 17), (2490, 2513), (4000, 5001)):
        xs = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        wsl = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        fgq._lib.check(L.fgq_gausshermite_device(n, 0, lo, hi, xs.data_ptr(), wsl.data_ptr(), ws.data_ptr(), 0))
        torch.cuda.synchronize()
        assert np.array_equal(xs.cpu().numpy(), x[lo:hi]) and np.array_equal(wsl.cpu().numpy(), w[lo:hi])
