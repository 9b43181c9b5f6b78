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


@pytest.mark.parametrize("n", [1, 2, 20, 21, 200, 201, 251, 372, 373, 5000, 100_001, 1_000_000])
@pytest.mark.parametrize("normalize", [False, True])
def test_reduced_rule_is_the_centre_block_of_the_full_rule(fgq, n, normalize):
    """Truncated rule: the returned block matches the full rule to the Newton tolerance,
    every omitted weight of the full rule is below 10 floatmin, and the moments still hold."""
    xf, wf = fgq.gausshermite(n, normalize=normalize)
    first, x, w = fgq.gausshermite_reduced(n, normalize=normalize)
    m = len(x)
    assert first >= 0 and first + m <= n and first == n - (first + m)  # symmetric block
    assert np.all(wf[:first] < 10 * np.finfo(float).tiny) and np.all(wf[first + m:] < 10 * np.finfo(float).tiny)
    assert np.all(w >= 10 * np.finfo(float).tiny)
    if first > 0:   # maximal: the next weight outside is below the threshold in the full rule too
        assert wf[first - 1] < 10 * np.finfo(float).tiny
    xs, ws = xf[first:first + m], wf[first:first + m]
    scale = np.sqrt(2.0 * n + 1) * (np.sqrt(2.0) if normalize else 1.0)
    assert np.abs(x - xs).max() <= 8 * np.spacing(scale)
    big = ws > 1e-280
    # One ulp of theta moves a node by ~ulp(scale) (the two runs may stop Newton one sweep apart); through
    # exp(-x^2) and the residual term x*val of the weight formula (gausshermite.jl:83-86) that shows up
    # ~4|x| times larger, relatively, in the weight
    dx = np.maximum(np.abs(x - xs), 2 * np.spacing(scale))
    assert np.all(np.abs(w[big] / ws[big] - 1) <= 5e-12 + 8 * np.abs(xs[big]) * dx[big])
    half = m // 2   # (the centre node of an odd rule is whatever Newton left of cos(pi/2): not compared)
    assert np.array_equal(x[:half], -x[::-1][:half]) and np.array_equal(w, w[::-1])
    if n <= 200 or fgq.lib().fgq_gausshermite_reduced_bound(n) == n:
        assert np.array_equal(x, xs) and np.array_equal(w, ws)  # nothing truncated: the very same launches
    mu0 = 1.0 if normalize else np.sqrt(np.pi)
    mu2 = 1.0 if normalize else np.sqrt(np.pi) / 2
    assert abs(w.sum() - mu0) < 2e-13 * max(1.0, mu0)
    if n >= 2:
        assert abs(np.dot(w, x * x) - mu2) < 3e-12


def test_reduced_rule_size(fgq):
    """2.4 % of the nodes at n = 1e6."""
    first, x, w = fgq.gausshermite_reduced(1_000_000)
    assert 0.02 * 1e6 < len(x) < 0.03 * 1e6
    assert fgq.lib().fgq_gausshermite_reduced_bound(1_000_000) < 0.03 * 1e6
    f2, xd, wd, _ = fgq.gausshermite_reduced_device(1_000_000)
    xd, wd = xd.cpu().numpy(), wd.cpu().numpy()
    off = first - f2
    assert off >= 0 and np.array_equal(xd[off:off + len(x)], x) and np.array_equal(wd[off:off + len(x)], w)
    assert np.all(wd[:off] < 10 * np.finfo(float).tiny)
    with pytest.raises(fgq.DomainError):
        fgq.gausshermite_reduced(-1)
    assert len(fgq.gausshermite_reduced(0)[1]) == 0


def test_fused_launch_equals_multi_launch_form():
    """n > 200 runs as one cooperative launch (hermite_fused_kernel); FGQ_HERMITE_FUSED=0 selects the
    multi-launch form it replaced.  Both must give the same bits (full and truncated rule, both normalisations,
    even and odd n) — checked in two subprocesses because the switch is read once per process."""
    import subprocess
    import sys
    code = (
        "import hashlib, sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(HERE)!r})\n"
        "import fgq_b200\n"
        "h = hashlib.sha256()\n"
        "for n in (201, 1000, 5001, 100000, 1000001):\n"
        "    for nz in (False, True):\n"
        "        for a in fgq_b200.gausshermite(n, normalize=nz): h.update(a.tobytes())\n"
        "        for a in fgq_b200.gausshermite_reduced(n, normalize=nz)[1:]: h.update(a.tobytes())\n"
        "print(h.hexdigest())\n")
    outs = []
    for fused in ("0", "1"):
        env = dict(os.environ, FGQ_HERMITE_FUSED=fused)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] and len(outs[0]) == 64
