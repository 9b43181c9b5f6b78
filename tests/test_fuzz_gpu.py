"""Seeded random sweep over (n, parameters) of every family against the oracles — corner-case hunting
beyond the reference's own test points.  Tolerance models as in the per-family parity tests."""
import numpy as np
import pytest
from scipy.special import beta

from oracle import hermite_oracle as ho
from oracle import jacobi_oracle as jo
from oracle import laguerre_oracle as lo

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def test_fuzz_legendre(fgq, oracle_mod):
    rng = np.random.default_rng(11)
    for n in np.concatenate([rng.integers(0, 62, 20), rng.integers(61, 5000, 25), rng.integers(5000, 3_000_000, 10)]):
        n = int(n)
        x, w = fgq.gausslegendre(n)
        xt, wt = oracle_mod.gausslegendre(n, twin=True)
        assert np.array_equal(x, xt) and np.array_equal(w, wt), n


def test_fuzz_jacobi(fgq):
    rng = np.random.default_rng(12)
    for _ in range(60):
        n = int(rng.choice([rng.integers(2, 101), rng.integers(101, 3000), rng.integers(3000, 40000)]))
        a, b = rng.uniform(-0.95, 4.9, 2)
        if rng.random() < 0.2:
            a = float(rng.integers(0, 5))
        if rng.random() < 0.2:
            b = a
        x, w = fgq.gaussjacobi(n, a, b)
        xo, wo = jo.gaussjacobi(n, a, b)
        assert np.abs(x - xo).max() <= 4 * EPS, (n, a, b, np.abs(x - xo).max() / EPS)
        bound = 3e-13 + 2 * (2 + abs(a) + abs(b)) / (1 - np.abs(xo)) * (np.abs(x - xo) + 2 * EPS)
        rel = np.abs(w - wo) / wo
        assert np.all(rel <= bound), (n, a, b, float(np.max(rel / bound)))
        mu0 = 2 ** (a + b + 1) * beta(a + 1, b + 1)
        assert abs(w.sum() - mu0) < 5e-13 * mu0, (n, a, b)


def test_fuzz_hermite(fgq):
    rng = np.random.default_rng(13)
    for n in np.concatenate([rng.integers(2, 21, 5), rng.integers(21, 201, 10), rng.integers(201, 60000, 15)]):
        n = int(n)
        x, w = fgq.gausshermite(n)
        xo, wo = ho.gausshermite(n)
        scale = np.sqrt(2 * n + 1)
        assert np.abs(x - xo).max() <= 8 * np.spacing(scale), n
        nz = wo > 1e-290
        assert np.all(np.abs(w[nz] - wo[nz]) / wo[nz] <= 1e-13 + 4 * np.abs(xo[nz]) * 8 * np.spacing(scale)), n
        assert abs(w.sum() - np.sqrt(np.pi)) < 1e-13


def test_fuzz_laguerre(fgq):
    from math import gamma
    rng = np.random.default_rng(14)
    for _ in range(40):
        n = int(rng.choice([rng.integers(1, 128), rng.integers(128, 3000), rng.integers(3000, 60000)]))
        al = float(rng.uniform(-0.95, 4.9))
        if rng.random() < 0.25:
            al = float(rng.integers(0, 5))
        x, w = fgq.gausslaguerre(n, al)
        xo, wo = lo.gausslaguerre(n, al)
        d = 1.0 / (4 * n + 2 * al + 2)
        tt = np.clip(xo * d, 1e-300, 1 - 1e-12)
        if n >= 128:
            rel_tol = (16 + 32 / np.sqrt(tt) + 32 / np.sqrt(1 - tt)) * EPS
        elif 3 <= n < 15:
            rel_tol = 8 * EPS * xo.max() / xo
        else:
            rel_tol = np.full(n, 256 * EPS)
        rx = np.abs(x - xo) / xo
        assert np.all(rx <= rel_tol), (n, al, float(np.max(rx / rel_tol)))
        nz = wo > 1e-290
        bound = 1e-11 + 2 * (xo[nz] + 1) * rel_tol[nz]
        rw = np.abs(w[nz] - wo[nz]) / wo[nz]
        assert np.all(rw <= bound), (n, al, float(np.max(rw / bound)))
        if abs(wo.sum() - gamma(al + 1)) < 1e-12 * gamma(al + 1):
            assert abs(w.sum() - gamma(al + 1)) < 3e-12 * gamma(al + 1), (n, al)
