"""GPU parity tests for Gauss-Jacobi / Radau / Lobatto (asymptotic branch), through the C ABI.

Parity metric (SURVEY.md §7 hard part 1): bit-exact theta is not attainable here (BLAS summation
order inside the reference's feval_asy1, AMOS besselj at the boundary, sin/cos of 1e7-radian
phases), so nodes are compared in ulp(1) — the absolute scale of a rule on [-1, 1] — and weights
relatively.  Fixed endpoints of Radau/Lobatto must be exact."""
import ctypes
import json
import os

import numpy as np
import pytest
from scipy.special import beta, jv

from oracle import jacobi_oracle as jo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "jacobi_reference_tests.json")))
EPS = np.finfo(float).eps


def _at(arr, i):
    i = int(i)
    return arr[i - 1] if i > 0 else arr[i]


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}_{c['a']}_{c['b']}")
def test_reference_golden_values(fgq, case):
    x, w = fgq.gaussjacobi(case["n"], case["a"], case["b"])
    assert x.shape == (case["n"],) and x.dtype == np.float64
    for i, v in case["x"].items():
        assert abs(_at(x, i) - v) < GOLD["tol"]
    for i, v in case["w"].items():
        assert abs(_at(w, i) - v) < GOLD["tol"]
    if "exp_integral" in case:
        assert abs(np.dot(w, np.exp(x)) - case["exp_integral"]) < 1.5e-8 * case["exp_integral"]
    if "sum_w" in case:
        assert abs(w.sum() - 2 ** (case["a"] + 1) / (case["a"] + 1)) < GOLD["tol"]


def test_dispatch_and_errors(fgq):
    with pytest.raises(fgq.DomainError):       # test_gaussjacobi.jl:2-6
        fgq.gaussjacobi(-1, 0.1, 0.2)
    with pytest.raises(fgq.DomainError):
        fgq.gaussjacobi(200, -1.0, 0.2)
    with pytest.raises(fgq.DomainError):
        fgq.gaussjacobi(200, 0.3, -1.5)
    x, w = fgq.gaussjacobi(20, 11.0, 0.0)          # test_gaussjacobi.jl:212-216 (Golub-Welsch region)
    assert x.shape == (20,) and x.dtype == np.float64
    x, w = fgq.gaussjacobi(1, 1.0, 2.0)            # test_gaussjacobi.jl:171-179
    assert abs(x[0] - 0.2) < 1e-14 and abs(w[0] - 4 / 3) < 1e-14
    with pytest.raises(fgq.NotImplementedOnDevice):  # gaussjacobi.jl:53 (ErrorException in the reference)
        fgq.gaussjacobi(5000, 6.0, 0.0)
    x, w = fgq.gaussjacobi(0, 0.3, 0.1)
    assert len(x) == 0
    xl, wl = fgq.gausslegendre(300)             # (0,0) -> Legendre (:28-29)
    xj, wj = fgq.gaussjacobi(300, 0, 0)
    assert np.array_equal(xl, xj) and np.array_equal(wl, wj)
    x1, w1 = fgq.gaussjacobi(500, 0, -1 / 2)    # integer promotion (#117)
    x2, w2 = fgq.gaussjacobi(500, 0.0, -1 / 2)
    assert np.array_equal(x1, x2) and np.array_equal(w1, w2)
    with pytest.raises(fgq.DomainError):        # test_gausslobatto.jl:3-5, test_gaussradau.jl:3-4
        fgq.gausslobatto(1)
    with pytest.raises(fgq.DomainError):
        fgq.gaussradau(0)


def test_besselroots_and_besselj(fgq):
    # test/test_besselroots.jl:7-12 on the library's approx_besselroots (one thread per zero on the device)
    for nu in np.arange(0, 5.01, 0.1):
        r = fgq.approx_besselroots(float(nu), 10)
        assert np.max(np.abs(jv(nu, r))) < 1e-11
        assert np.allclose(r, jo.approx_besselroots(float(nu), 10), rtol=1e-14, atol=0)
    # every branch of besselroots.jl:28-67 at lengths beyond the host-planned head
    for nu in (0.0, -1.0, -0.5, 0.3, 5.0, 5.5, 12.25):
        for n in (0, 1, 5, 6, 7, 20, 21, 1000, 100003):
            r = fgq.approx_besselroots(nu, n)
            ro = jo.approx_besselroots(nu, n)
            assert r.shape == ro.shape
            assert np.all(np.abs(r - ro) <= 2 * np.spacing(np.abs(ro))), (nu, n)
    assert np.array_equal(fgq.approx_besselroots(-1.5, 8), np.zeros(8))   # no branch of the reference matches: zeros
    with pytest.raises(fgq.DomainError):
        fgq.approx_besselroots(0.0, -1)
    # device-resident variant
    import torch
    out = torch.empty(1000, dtype=torch.float64, device="cuda")
    assert fgq.lib().fgq_approx_besselroots_device(0.3, 1000, out.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), fgq.approx_besselroots(0.3, 1000))
    # besselj stand-in (double-double ascending series) against mpmath; AMOS (scipy.special.jv, what
    # the reference calls) is itself only good to ~5e-14 of the local amplitude on this grid
    import mpmath as mp
    mp.mp.dps = 40
    L = fgq.lib()
    for nu in (0.0, 0.3, -0.7, 1.0, -0.9, 4.5, -29.7, 30.3, -30.0, 31.0):
        for xx in (0.01, 2.4, 7.0, 15.0, 31.4, 36.0):
            ref = float(mp.besselj(mp.mpf(nu), mp.mpf(xx)))
            amp = abs(ref) if xx < abs(nu) else max(abs(ref), 0.1 * np.sqrt(2 / (np.pi * xx)))
            assert abs(L.fgq_besselj(nu, xx) - ref) <= 1e-15 * amp
            assert abs(L.fgq_besselj(nu, xx) - float(jv(nu, xx))) <= 2e-13 * amp


PAIRS = [(251, -0.7, 1.3), (1013, 0.9, -0.1), (10013, 0.9, -0.1), (10000, 0.1, 0.2), (1024, 0.25, 0.0),
         (65536, -0.9, 0.0), (101, 0.3, -0.7), (5000, 4.5, 2.0), (100000, 0.3, -0.7), (1000000, 0.3, -0.7),
         (2001, 1.0, 1.0), (2000, 0.0, 1.0)]


@pytest.mark.parametrize("n,a,b", PAIRS)
def test_parity_with_oracle(fgq, n, a, b):
    info = {}
    xo, wo = jo.gaussjacobi(n, a, b, info)
    x, w, ws = fgq.gaussjacobi_device(n, a, b)
    import torch
    torch.cuda.synchronize()
    x = x.cpu().numpy()
    w = w.cpu().numpy()
    sw = (ctypes.c_int * 2)()
    tm = (ctypes.c_int * 2)()
    fgq._lib.check(fgq.lib().fgq_gaussjacobi_stats(ws.data_ptr(), sw, tm))
    ex = np.abs(x - xo) / EPS
    rw = np.abs(w - wo) / np.abs(wo)
    print(f"n={n} a={a} b={b}: node err max {ex.max():.2f} ulp(1) (boundary {max(ex[:10].max(), ex[-10:].max()):.2f}); "
          f"weight rel err max {rw.max():.2e}; sweeps gpu {tuple(sw)} oracle {info['sweeps']}; terms gpu {tuple(tm)} oracle {info['terms'][-1]}")
    assert tuple(sw) == tuple(info["sweeps"]), "device-side global Newton stopping differs from the reference rule"
    assert ex.max() <= 4, f"nodes differ by {ex.max()} ulp(1)"
    assert rw.max() <= 2e-13, f"weights differ by {rw.max()} relatively"
    assert np.all(np.diff(x) > 0)
    assert abs(w.sum() - 2 ** (a + b + 1) * beta(a + 1, b + 1)) < 2e-13 * max(1.0, w.sum())


def test_radau_lobatto_reference_goldens(fgq):
    # test/test_gaussradau.jl:6-29, test/test_gausslobatto.jl:13-36
    for rule, key in ((fgq.gaussradau, "radau"), (fgq.gausslobatto, "lobatto")):
        g = GOLD[key]
        x, w = rule(g["n"])
        assert len(x) == g["n"]
        for i, v in g["x"].items():
            assert abs(x[int(i) - 1] - v) < GOLD["tol"]
        for i, v in g["w"].items():
            assert abs(w[int(i) - 1] - v) < GOLD["tol"]
        assert x[0] == -1.0 and abs(np.dot(w, np.exp(x)) - (np.e - 1 / np.e)) < 1e-14
    x, w = fgq.gaussradau(1)
    assert x[0] == -1.0 and w[0] == 2.0
    x, w = fgq.gaussradau(2)
    assert np.allclose(x, [-1, 1 / 3], rtol=1e-15) and np.allclose(w, [0.5, 1.5], rtol=1e-15)
    x, w = fgq.gausslobatto(2)
    assert np.array_equal(x, [-1, 1]) and np.array_equal(w, [1, 1])
    x, w = fgq.gausslobatto(3)
    assert np.array_equal(x, [-1, 0, 1]) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3], rtol=1e-15)
    x, w = fgq.gausslobatto(42)
    assert x[-1] == 1.0 and abs(np.dot(w, x * x) - 2 / 3) < 1e-14


@pytest.mark.parametrize("n", list(range(2, 12)) + [20, 41, 42, 57, 99, 100])
@pytest.mark.parametrize("a,b", [(-0.1, 0.3), (0.9, -0.1), (1.0, 1.0), (0.0, 1.0), (4.5, -0.9), (0.3, -0.7)])
def test_recurrence_branch_parity(fgq, n, a, b):
    """n <= 100: jacobi_rec (gaussjacobi.jl:61-155), one block per rule."""
    x, w = fgq.gaussjacobi(n, a, b)
    xo, wo = jo.gaussjacobi(n, a, b)
    assert np.abs(x - xo).max() <= 4 * EPS
    # w = c / ((1 - x^2) P'(x)^2 sum): one ulp of x moves the weight by ~2|x|/(1 - x^2) ulps
    # (more precisely by ~(2 + |a| + |b|)/(1 - |x|): P'' / P' = (a - b + (a + b + 2) x)/(1 - x^2) at a root)
    bound = 64 * EPS + 2 * (2 + abs(a) + abs(b)) / (1 - np.abs(xo)) * (np.abs(x - xo) + 2 * EPS)
    rel = np.abs(w - wo) / wo
    assert np.all(rel <= bound), f"worst {np.max(rel / bound)} x bound"
    assert abs(w.sum() - 2 ** (a + b + 1) * beta(a + 1, b + 1)) < 2e-13 * max(1.0, w.sum())


def test_compatible_with_legendre_and_chebyshev(fgq):
    """test_gaussjacobi.jl:8-147: gaussjacobi at (0,0), (+-1/2,+-1/2) and their floating-point
    neighbours agrees with Gauss-Legendre / Gauss-Chebyshev: ||dx||_2 < 4e-15, ||dw||_2 < 1e-13."""
    cheb = {(-0.5, -0.5): fgq.gausschebyshevt, (0.5, 0.5): fgq.gausschebyshevu,
            (-0.5, 0.5): fgq.gausschebyshevv, (0.5, -0.5): fgq.gausschebyshevw}
    n_range = list(range(0, 11)) + list(range(20, 101, 10)) + [120, 300]
    for (a0, b0), ref in [((0.0, 0.0), fgq.gausslegendre)] + list(cheb.items()):
        for n in n_range:
            x, w = ref(n)
            for a in (a0, np.nextafter(a0, 1), np.nextafter(a0, -1)):
                for b in (b0, np.nextafter(b0, 1), np.nextafter(b0, -1)):
                    xj, wj = fgq.gaussjacobi(n, a, b)
                    assert np.linalg.norm(x - xj) < 4e-15, (n, a, b)
                    assert np.linalg.norm(w - wj) < 1e-13, (n, a, b)


@pytest.mark.parametrize("n", [1001, 2002, 100003])
def test_radau_lobatto(fgq, n):
    x, w = fgq.gaussradau(n)
    xo, wo = jo.gaussradau(n)
    assert x[0] == -1.0 and w[0] == 2 / n ** 2            # test_gaussradau.jl:27
    assert np.abs(x - xo).max() <= 4 * EPS and (np.abs(w - wo) / wo).max() <= 2e-13
    assert abs(w.sum() - 2) < 1e-13 and abs(np.dot(w, x * x) - 2 / 3) < 1e-13
    x, w = fgq.gausslobatto(n)
    xo, wo = jo.gausslobatto(n)
    assert x[0] == -1.0 and x[-1] == 1.0                   # test_gausslobatto.jl:33
    assert w[0] == 2 / (n * (n - 1)) and w[-1] == w[0]
    assert np.abs(x - xo).max() <= 4 * EPS and (np.abs(w - wo) / wo).max() <= 2e-13
    assert abs(w.sum() - 2) < 1e-13 and abs(np.dot(w, x * x) - 2 / 3) < 1e-13


def test_config_c3_full_size(fgq):
    """BASELINE config C3: gaussjacobi(1e7, 0.3, -0.7) on one GPU.  The reference itself raises
    BoundsError here (see test_oracle_jacobi.py); checked through invariants and ordering."""
    import torch
    n, a, b = 10 ** 7, 0.3, -0.7
    x, w, ws = fgq.gaussjacobi_device(n, a, b)
    torch.cuda.synchronize()
    assert bool((x[1:] > x[:-1]).all())
    assert abs(float(w.sum()) - 4.554443087962173) < 1e-12
    # first moment of the Jacobi weight: (b - a)/(a + b + 2) * mu0
    assert abs(float((w * x).sum()) - (b - a) / (a + b + 2) * 4.554443087962173) < 1e-12
    assert float(x[0]) > -1 and float(x[-1]) < 1


def test_config_c3_every_node_against_the_oracle(fgq):
    """BASELINE config C3 at FULL size: all 1e7 nodes and weights of gaussjacobi(1e7, 0.3, -0.7) against the oracle,
    evaluated in node chunks (oracle.jacobi_oracle.jacobi_asy_chunked: the reference's interior solve with its two
    global decisions — series truncation index, Newton stopping rule — reproduced exactly over the chunks;
    bit-identical to the dense form, tests/test_oracle_jacobi.py).  Same bounds as the smaller cases: nodes within
    4 ulp(1), weights within 2e-13 relative, same sweep counts and truncation indices."""
    import torch
    n, a, b = 10 ** 7, 0.3, -0.7
    info = {}
    xo, wo = jo.jacobi_asy_chunked(n, a, b, chunk=250_000, info=info)
    x, w, ws = fgq.gaussjacobi_device(n, a, b)
    torch.cuda.synchronize()
    x = x.cpu().numpy()
    w = w.cpu().numpy()
    sw = (ctypes.c_int * 2)()
    tm = (ctypes.c_int * 2)()
    fgq._lib.check(fgq.lib().fgq_gaussjacobi_stats(ws.data_ptr(), sw, tm))
    ex = np.abs(x - xo) / EPS
    rw = np.abs(w - wo) / np.abs(wo)
    print(f"C3 full size: node err max {ex.max():.2f} ulp(1), mean {ex.mean():.3f}; weight rel err max {rw.max():.2e}; "
          f"sweeps gpu {tuple(sw)} oracle {info['sweeps']}; terms gpu {tuple(tm)} oracle {info['terms']}")
    assert tuple(sw) == tuple(info["sweeps"])
    assert tm[0] == info["terms"][info["sweeps"][0]] and tm[1] == info["terms"][-1]
    assert ex.max() <= 4, f"nodes differ by {ex.max()} ulp(1)"
    assert rw.max() <= 2e-13, f"weights differ by {rw.max()} relatively"
    assert np.all(np.diff(x) > 0)


@pytest.mark.parametrize("n,a,b", [(2, 6.0, 0.3), (6, 5.0, 5.0), (7, 6.5, 6.5), (20, 11.0, 0.0), (100, 19.0, 21.0), (257, 5.0, -0.5), (1000, 7.5, 12.0), (4000, 5.5, 0.0)])
def test_golub_welsch_region(fgq, n, a, b):
    """max(a,b) >= 5 (gaussjacobi.jl:50-51, 570-599): the reference's dense eigen-solve against the
    per-node bisection + recurrence kernel.  LAPACK delivers eigenvalues to eps*||J|| absolutely."""
    x, w = fgq.gaussjacobi(n, a, b)
    xo, wo = jo.gaussjacobi(n, a, b)
    assert np.abs(x - xo).max() <= 16 * EPS
    # LAPACK returns the eigenvector matrix to eps ABSOLUTELY, so the oracle's weight mu0*V[1,k]^2 is only good
    # to ~n*eps*sqrt(w*mu0); the recurrence formula of the kernel keeps relative accuracy on tiny weights.
    mu0 = jo.jacobimoment(a, b)
    bound = 64 * n * EPS * np.sqrt(wo * mu0) + 1e-12 * wo
    print(f"GW n={n} a={a} b={b}: node err {np.abs(x - xo).max() / EPS:.1f} eps, weight err / bound {np.max(np.abs(w - wo) / bound):.3f}")
    assert np.all(np.abs(w - wo) <= bound)
    assert abs(w.sum() - jo.jacobimoment(a, b)) < 1e-13 * w.sum()
    assert np.all(np.diff(x) > 0)


def test_general_radau(fgq):
    """gaussradau(n, a, b) (gaussradau.jl:52-69; test_gaussradau.jl:31-50)."""
    for a, b in ((0.1, 0.2), (0.0, -0.5)):
        for n in list(range(1, 12)) + [40, 257, 1000]:
            x, w = fgq.gaussradau(n, a, b)
            xo, wo = jo.gaussradau_jacobi(n, a, b)
            assert x[0] == -1.0 and len(x) == n
            print(n, a, b, np.abs(x - xo).max() / EPS, np.abs(w - wo).max())
            assert np.abs(x - xo).max() <= 16 * EPS and np.abs(w - wo).max() <= 1e-13
            if n > 5:
                continue
            xj, wj = jo.gaussjacobi(n, a, b)
            assert abs(w.sum() - wj.sum()) < 1e-14
            for j in range(1, 2 * n - 1):
                assert abs(np.dot(w, x ** j) - np.dot(wj, xj ** j)) < 1e-14
    for n in (2, 3, 40):   # compare with Gauss-Radau-Legendre
        x, w = fgq.gaussradau(n, 0, 0)
        xr, wr = fgq.gaussradau(n)
        print(n, np.abs(x - xr).max(), np.abs(w - wr).max())
        assert np.allclose(x, xr, rtol=0, atol=1e-13) and np.allclose(w, wr, rtol=0, atol=1e-13)
    with pytest.raises(fgq.DomainError):
        fgq.gaussradau(0, 0, 0)
