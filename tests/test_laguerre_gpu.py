"""GPU parity tests for the explicit-asymptotics Gauss-Laguerre path (C ABI via ctypes).

Nodes are explicit series: they are compared relative to the node itself (ulps of x); weights are
compared relatively where they have not underflowed, and must underflow to exactly the same zeros.
The hand-over indices Bessel->bulk->Airy and the set of recomputed nodes must be the reference's."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import laguerre_oracle as lo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "laguerre_reference_tests.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}_a{c['alpha']}")
def test_reference_golden_values(fgq, case):
    x, w = fgq.gausslaguerre(case["n"], case["alpha"])
    assert x.shape == (case["n"],)
    for i, v in case["x"].items():
        assert abs(x[int(i) - 1] - v) < GOLD["tol"]
    for i, v in case["w"].items():
        assert abs(w[int(i) - 1] - v) < GOLD["tol"]
    if "cos_a" in case:
        assert abs(np.dot(w, np.cos(case["cos_a"] * x)) - case["cos_integral"]) < GOLD["tol"]


def test_api_edges(fgq):
    with pytest.raises(fgq.DomainError):       # test_gausslaguerre.jl:37-38
        fgq.gausslaguerre(-1)
    with pytest.raises(fgq.DomainError):
        fgq.gausslaguerre(200, -1.0)
    x, w = fgq.gausslaguerre(0)
    assert len(x) == 0
    with pytest.raises(fgq.NotImplementedOnDevice):
        fgq.gausslaguerre(5000, 6.0)            # alpha >= 5 -> O(n^2) recurrence rule, device path stops at n = 4096
    for c in GOLD["closed_forms"]:              # test_gausslaguerre.jl:66-72
        x, w = fgq.gausslaguerre(c["n"], c["alpha"])
        assert np.allclose(x, c["x"], rtol=1e-15, atol=0) and np.allclose(w, c["w"], rtol=1e-14, atol=0)
    xr, wr = fgq.gausslaguerre(350000, 0.44, reduced=True)   # test_gausslaguerre.jl:142-144
    assert wr[-1] < 100 * np.finfo(float).tiny and len(xr) < 350000
    xo, wo = lo.gausslaguerre(350000, 0.44, reduced=True)
    assert len(xr) == len(xo)


@pytest.mark.parametrize("n,alpha", [(128, 0.0), (200, 0.0), (251, 0.0), (256, 0.0), (400, 0.0), (251, 0.5), (1000, -0.7), (1000, 4.5), (5000, 1.0),
                                     (350000, 0.0), (350000, 0.44), (1000000, 0.0), (1000001, 2.5)])
def test_parity_with_oracle(fgq, n, alpha):
    import torch
    info = {}
    xo, wo = lo.gausslaguerre(n, alpha, info=info)
    x, w, ws = fgq.gausslaguerre_device(n, alpha)
    torch.cuda.synchronize()
    x = x.cpu().numpy()
    w = w.cpu().numpy()
    kb, ka, nf, nr = (ctypes.c_int64() for _ in range(4))
    fgq._lib.check(fgq.lib().fgq_gausslaguerre_stats(ws.data_ptr(), ctypes.byref(kb), ctypes.byref(ka),
                                                    ctypes.byref(nf), ctypes.byref(nr)))
    ka_o = info["k_a"] if info["k_a"] is not None else n + 1
    eps = np.finfo(float).eps
    d = 1.0 / (4 * n + 2 * alpha + 2)
    # Tolerance model.  Most nodes are explicit series and agree to a few ulp.  In the bulk the node
    # is x = t/d + ... with t the root of  pt*pi + 2 sqrt(t - t^2) - acos(2t - 1)  (gausslaguerre.jl:420):
    # the pi-sized terms cancel, so the root is only defined to ~eps*sqrt(t) absolutely, i.e. to
    # eps/sqrt(t) = eps/sqrt(x d) relatively (and to eps/sqrt(1-t) near the soft edge, where the
    # derivative 2 sqrt((1-t)/t) vanishes) — two libms land anywhere inside that band.  Measured:
    # Bessel- and Airy-region nodes agree bit-for-bit; bulk nodes sit at <= 9 eps/sqrt(t), 5 eps/sqrt(1-t).
    tt = np.clip(xo * d, 1e-300, 1 - 1e-12)
    rel_tol = (16 + 32 / np.sqrt(tt) + 32 / np.sqrt(1 - tt)) * eps
    rx = np.abs(x - xo) / xo
    nz = (wo > 1e-290)
    rw = np.abs(w[nz] - wo[nz]) / wo[nz]
    print(f"n={n} alpha={alpha}: k_b {kb.value}/{info['k_b']} k_a {ka.value}/{ka_o} flagged {nf.value}/{len(info['flagged'])}; "
          f"node rel err max {rx.max() / eps:.1f} eps, worst/tol {np.max(rx / rel_tol):.2f}; weight rel err max {rw.max():.2e}")
    assert kb.value == info["k_b"] and ka.value == ka_o
    assert nf.value == len(info["flagged"])
    assert np.all(rx <= rel_tol), f"nodes: worst {np.max(rx / rel_tol)} x tolerance"
    # weights carry x^alpha e^{-x} sqrt(t/(1-t)): a relative node difference r changes them by ~(x + 1) r
    bound = 4e-13 + 2 * (xo[nz] + 1) * rel_tol[nz]
    assert np.all(rw <= bound), f"weights: worst {np.max(rw / bound)} x bound"
    assert np.count_nonzero((w == 0) != (wo == 0)) <= 2
    assert np.all(np.diff(x) > 0)
    if alpha == 0.0:
        assert abs(w.sum() - 1) < 1e-13 and abs(np.dot(w, x) - 1) < 1e-13
    # Bessel region (k < k_b): explicit series in the Bessel zero — the reference takes the tabulated / Piessens zero
    # only for k < k_bessel = max(ceil(sqrt(n)), 7) and McMahon's expansion from k_bessel on (gausslaguerre.jl:131,144;
    # for alpha = 0 and n < 441 that is inside the 20 tabulated zeros, which differ from McMahon by an ulp): bit for bit
    if alpha == 0.0:   # (for alpha != 0 the McMahon coefficients are planned in a different association: a few ulp)
        nb = info["k_b"] - 1
        assert np.array_equal(x[:nb], xo[:nb]), "Bessel-region nodes must equal the oracle's bit for bit"


@pytest.mark.parametrize("alpha", [0.0, 0.5, -0.7, 2.5, 6.0])
def test_small_rules_parity(fgq, alpha):
    """n < 128 (and alpha >= 5): closed forms, Golub-Welsch (on-device QL vs the oracle's LAPACK) and the
    sequential recurrence rule (gausslaguerre.jl:70-95, 535-542, 582-618)."""
    from math import gamma
    for n in [1, 2, 3, 5, 10, 14, 15, 16, 31, 42, 64, 100, 127] + ([128, 300] if alpha >= 5 else []):
        x, w = fgq.gausslaguerre(n, alpha)
        xo, wo = lo.gausslaguerre(n, alpha)
        eps = np.finfo(float).eps
        if 3 <= n < 15:   # eigenvalues of the Jacobi matrix: absolute accuracy eps*||J|| in LAPACK and in the QL kernel
            assert np.abs(x - xo).max() <= 8 * eps * xo.max(), (n, alpha)
        else:
            # the rule stops Newton at |step| <= 40 eps x (gausslaguerre.jl:556): two libms stop up to ~2 x 40 eps apart
            assert (np.abs(x - xo) / xo).max() <= 256 * eps, (n, alpha)
        nz = wo > 1e-290
        # recurrence rule: w = c / (p_{n-1}(x) p_n'(x~)) with p_n' taken at the PREVIOUS Newton iterate
        # (gausslaguerre.jl:556-575); one Newton step more or less moves it at the 1e-12 level
        assert np.all(np.abs(w[nz] - wo[nz]) / wo[nz] <= 1e-11 + 2 * np.abs(x - xo)[nz]), (n, alpha)
        assert np.count_nonzero((w == 0) != (wo == 0)) == 0
        # sum w = Gamma(alpha + 1) wherever the reference algorithm itself delivers it (at n = 15, alpha = 6 its
        # extrapolated initial guesses make Newton find one root twice: sum w = 820 instead of 720, reproduced)
        if abs(wo.sum() - gamma(alpha + 1)) < 1e-12 * gamma(alpha + 1):
            assert abs(w.sum() - gamma(alpha + 1)) < 2e-12 * gamma(alpha + 1)
    # `reduced` only exists inside gausslaguerre_asy (gausslaguerre.jl:88-95): the small-n / alpha >= 5 branches
    # return all n entries whatever the flag says
    xr, wr = fgq.gausslaguerre(100, alpha, reduced=True)
    xo, wo = lo.gausslaguerre(100, alpha, reduced=True)
    assert len(xr) == len(xo) == 100
    if alpha >= 5:
        xr, wr = fgq.gausslaguerre(2000, alpha, reduced=True)
        assert len(xr) == 2000 and wr[-1] == 0.0


def test_large_rule_launch_forms_agree():
    """The asymptotic rule runs the hand-over searches beside the bulk nodes (laguerre_edge_eval / edge_select / bulk_nodes /
    finish).  Its results, hand-over indices, flagged counts and reduced lengths must equal, bit for bit (digest), the
    sequential launch chain (FGQ_LAGUERRE_FAST=0) and its own fall-back, reached by shrinking the Bessel search window
    below k_b (FGQ_LAGUERRE_WINDOW=64).  The switches are read once per process, hence the subprocesses."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(HERE), "tools", "laguerre_time.py")
    outs = []
    for extra in ({}, {"FGQ_LAGUERRE_FAST": "0"}, {"FGQ_LAGUERRE_WINDOW": "64"}):
        env = dict(os.environ)
        env.update(extra)
        r = subprocess.run([sys.executable, tool], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0].count("digest") == 12
    assert outs[0] == outs[1], "large-rule path != sequential chain"
    assert outs[2] == outs[1], "fall-back != sequential chain"
