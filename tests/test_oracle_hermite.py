"""CPU tests: the Hermite oracle (numpy + AMOS Airy through scipy) against the reference's golden
values for the asymptotic branch (test/test_gausshermite.jl:38-54)."""
import json
import os

import numpy as np
import pytest

from oracle import hermite_oracle as ho

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hermite_reference_tests.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}")
def test_golden_values(case):
    n = case["n"]
    x, w = ho.gausshermite(n)
    assert len(x) == n and len(w) == n
    for i, (v, tol) in case["x"].items():
        assert abs(x[int(i) - 1] - v) < tol
    for i, (v, tol) in case["w"].items():
        assert abs(w[int(i) - 1] - v) < tol
    assert np.dot(w, x) < case["dot_wx_lt"]
    assert abs(np.dot(w, x * x) - np.sqrt(np.pi) / 2) < case["dot_wx2_tol"]
    if "poly_integral" in case:
        assert abs(np.dot(w, 1 + x + x ** 2 + x ** 3) - case["poly_integral"]) < 1.5e-8 * case["poly_integral"]  # Julia isapprox, rtol = sqrt(eps)


def test_structure_and_sweeps():
    info = {}
    x, w = ho.gausshermite(100000, info=info)
    assert info["sweeps"] <= 4
    assert np.all(np.diff(x) > 0)
    assert abs(w.sum() - np.sqrt(np.pi)) < 1e-13
    xn, wn = ho.gausshermite(1001, normalize=True)
    assert abs(wn.sum() - 1) < 1e-13 and abs(np.dot(wn, xn * xn) - 1) < 1e-11
    with pytest.raises(ValueError):
        ho.gausshermite(-1)


@pytest.mark.parametrize("n", [201, 5000, 100_001, 1_000_000])
def test_truncated_rule_bound_is_safe(n):
    """The a-priori size of the truncated rule (fgq_gausshermite_reduced_bound: Sturm spacing of the zeros of H_n,
    DESIGN.md 8.8) against the oracle's full rule: every node outside the centre block lies beyond |x| = 27.5 and
    carries the weight 0.0; needs no device."""
    import fgq_b200
    from oracle import hermite_oracle as ho
    bound = int(fgq_b200.lib().fgq_gausshermite_reduced_bound(n))
    assert 0 < bound <= n and (n - bound) % 2 == 0
    x, w = ho.gausshermite(n)
    first = (n - bound) // 2
    if first:
        assert np.all(np.abs(x[:first]) > 27.5) and np.all(np.abs(x[first + bound:]) > 27.5)
        assert np.all(w[:first] == 0.0) and np.all(w[first + bound:] == 0.0)
        # and it is not wasteful: the block reaches at most a few nodes beyond the cut
        assert np.abs(x[first + 8]) > 27.3
    assert abs(w[first:first + bound].sum() - np.sqrt(np.pi)) < 1e-13
