"""CPU tests: the Jacobi/Radau/Lobatto oracle (numpy + AMOS besselj through scipy) against the
reference's golden values for the asymptotic branch (test/test_gaussjacobi.jl, test_besselroots.jl)."""
import json
import os

import numpy as np
import pytest
from scipy.special import beta, jv

from oracle import jacobi_oracle as jo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "jacobi_reference_tests.json")))


def _at(arr, i):
    i = int(i)
    return arr[i - 1] if i > 0 else arr[i]


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}_{c['a']}_{c['b']}")
def test_golden_values(case):
    x, w = jo.gaussjacobi(case["n"], case["a"], case["b"])
    assert len(x) == case["n"]
    for i, v in case["x"].items():
        assert abs(_at(x, i) - v) < GOLD["tol"]
    for i, v in case["w"].items():
        assert abs(_at(w, i) - v) < GOLD["tol"]
    if "exp_integral" in case:
        assert abs(np.dot(w, np.exp(x)) - case["exp_integral"]) < 1.5e-8 * case["exp_integral"]
    if "sum_w" in case:
        assert abs(w.sum() - 2 ** (case["a"] + 1) / (case["a"] + 1)) < GOLD["tol"]
    assert np.all(np.diff(x) > 0)


def test_besselroots():
    # test/test_besselroots.jl:7-12
    for nu in np.arange(0, 5.01, 0.1):
        assert np.max(np.abs(jv(nu, jo.approx_besselroots(float(nu), 10)))) < 1e-11


def test_radau_lobatto_goldens():
    # test/test_gaussradau.jl:21-29, test/test_gausslobatto.jl:27-36 (n = 42: recurrence branch)
    for rule, key in ((jo.gaussradau, "radau"), (jo.gausslobatto, "lobatto")):
        g = GOLD[key]
        x, w = rule(g["n"])
        for i, v in g["x"].items():
            assert abs(x[int(i) - 1] - v) < GOLD["tol"]
        for i, v in g["w"].items():
            assert abs(w[int(i) - 1] - v) < GOLD["tol"]
        assert x[0] == -1.0 and abs(np.dot(w, np.exp(x)) - (np.e - 1 / np.e)) < 1e-14
    x, w = jo.gaussradau(2)
    assert np.allclose(x, [-1, 1 / 3]) and np.allclose(w, [0.5, 1.5])
    x, w = jo.gausslobatto(3)
    assert np.allclose(x, [-1, 0, 1]) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3])


def test_radau_lobatto_structure():
    x, w = jo.gaussradau(1001)
    assert x[0] == -1.0 and abs(w.sum() - 2) < 1e-13 and abs(np.dot(w, x ** 3)) < 1e-13
    x, w = jo.gausslobatto(1002)
    assert x[0] == -1.0 and x[-1] == 1.0 and abs(w.sum() - 2) < 1e-13 and abs(np.dot(w, x * x) - 2 / 3) < 1e-13


def test_large_n_continues_past_the_reference_bounds_error():
    """For boundary angles below ~6e-6 (n > ~5.1e6) the reference's nck_mat(floor(1.25 kmax)) is
    too small for kmax <= 3 and gaussjacobi raises BoundsError (gaussjacobi.jl:641,656-667); the
    oracle (and the library) use the same binomials and simply continue.  Check the invariant."""
    x, w = jo.gaussjacobi(6 * 10 ** 6, 0.3, -0.7)
    assert abs(w.sum() - 2 ** 0.6 * beta(1.3, 0.3)) < 1e-13
    assert np.all(np.diff(x) > 0)


def test_chunked_interior_solve_equals_the_dense_one():
    """jacobi_asy_chunked (used for the full-size C3 comparison on the GPU box) is the same function as
    jacobi_asy: identical bits, sweep counts and truncation indices, whatever the chunk size."""
    from oracle import jacobi_oracle as jo
    for n, a, b, chunk in ((2000, 0.3, -0.7, 777), (10013, 0.9, -0.1, 1000), (10000, 0.1, 0.2, 4096), (60000, 0.3, -0.7, 25000)):
        i1, i2 = {}, {}
        x, w = jo.jacobi_asy(n, a, b, i1)
        xc, wc = jo.jacobi_asy_chunked(n, a, b, chunk=chunk, info=i2)
        assert np.array_equal(x, xc) and np.array_equal(w, wc), (n, a, b)
        assert i1 == i2
