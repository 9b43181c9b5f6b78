"""CPU tests: the Laguerre oracle (numpy + AMOS through scipy + the C recurrence) against the
reference's golden values for the asymptotic branch (test/test_gausslaguerre.jl:109-144)."""
import json
import os

import numpy as np
import pytest

from oracle import laguerre_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "laguerre_reference_tests.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}_a{c['alpha']}")
def test_golden_values(case):
    x, w = lo.gausslaguerre(case["n"], case["alpha"])
    assert len(x) == case["n"]
    for i, v in case["x"].items():
        assert abs(x[int(i) - 1] - v) < GOLD["tol"]
    for i, v in case["w"].items():
        assert abs(w[int(i) - 1] - v) < GOLD["tol"]
    if "cos_a" in case:
        assert abs(np.dot(w, np.cos(case["cos_a"] * x)) - case["cos_integral"]) < GOLD["tol"]
    if case["alpha"] == 0.0:
        assert abs(w.sum() - 1) < 1e-13 and abs(np.dot(w, x) - 1) < 1e-13
    assert np.all(np.diff(x) > 0)


def test_closed_forms_and_small_rules():
    from math import gamma
    for c in GOLD["closed_forms"]:   # test_gausslaguerre.jl:66-72
        x, w = lo.gausslaguerre(c["n"], c["alpha"])
        assert np.allclose(x, c["x"], rtol=1e-15, atol=0) and np.allclose(w, c["w"], rtol=1e-14, atol=0)
    # GW vs rec vs default (test_gausslaguerre.jl:91-99)
    xg, wg = lo.gausslaguerre_gw(42, 0.0)
    xr, wr = lo.gausslaguerre_rec(42, 0.0)
    assert abs(xg[36] - 98.388267163326702) < 1e-13 and abs(xr[36] - 98.388267163326702) < 1e-13
    assert abs(wg[6] - 0.055372813167092) < 1e-13 and abs(wr[6] - 0.055372813167092) < 1e-13
    for n, al in ((10, 0.5), (64, 2.5), (100, 6.0)):
        x, w = lo.gausslaguerre(n, al)
        assert abs(w.sum() - gamma(al + 1)) < 1e-12 * gamma(al + 1)


def test_reduced_rule():
    # test_gausslaguerre.jl:142-144
    x, w = lo.gausslaguerre(350000, 0.44, reduced=True)
    assert w[-1] < 100 * np.finfo(float).tiny and len(x) < 350000


def test_handover_indices_match_survey():
    info = {}
    lo.gausslaguerre(10 ** 6, 0.0, info=info)
    assert info["k_b"] == 1809 and info["k_a"] == 999981 and len(info["flagged"]) == 34
