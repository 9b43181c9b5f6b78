"""Consumer fusion for a user integrand: dot(w, f.(x)) of README.md:27-30 with f supplied by the
caller as an expression.  Checked against the ORACLE's dot(w, f.(x)) — the same nodes and weights summed on the host in
extended precision — not only against analytic integrals."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("exp(x)", np.exp, math.e - 1 / math.e),
         ("x * x", lambda x: x * x, 2 / 3),
         ("exp(x) * cos(3 * x)", lambda x: np.exp(x) * np.cos(3 * x), None),
         ("1 / (1 + 25 * x * x)", lambda x: 1 / (1 + 25 * x * x), 2 * math.atan(5) / 5),
         ("pow(fabs(x), 2.5) + sin(x)", lambda x: np.abs(x) ** 2.5 + np.sin(x), 2 / 3.5)]


def _oracle_dot(x, w, f):
    return float(np.sum((w * f(x)).astype(np.longdouble)))


@pytest.mark.parametrize("n", [61, 251, 1013, 10013, 100000, 1000003])
@pytest.mark.parametrize("expr,f,exact", CASES, ids=[c[0] for c in CASES])
def test_fused_user_integrand_against_oracle_dot(fgq, oracle_mod, n, expr, f, exact):
    s = fgq.gausslegendre_integrate(n, expr).cpu().numpy()
    xo, wo = oracle_mod.gausslegendre(n, twin="julia")
    ref = _oracle_dot(xo, wo, f)
    scale = _oracle_dot(xo, wo, lambda x: np.abs(f(x)))
    assert abs(s[0] - ref) <= 4e-15 * scale + 1e-300, (s[0], ref)
    # sum w of the rule itself (the reference's n > 850000 branch drops the W1 term: sum w = 2 + O(1e-13) there)
    assert abs(s[1] - float(np.sum(wo.astype(np.longdouble)))) < 1e-13
    if exact is not None and n >= 251:
        assert abs(s[0] - exact) < 1e-12 * max(1.0, abs(exact)) or "fabs" in expr  # |x|^2.5 converges algebraically
    # any rank split gives the same sum
    parts = sum(fgq.gausslegendre_integrate(n, expr, rank=r, world=3) for r in range(3)).cpu().numpy()
    assert abs(parts[0] - s[0]) <= 4e-15 * scale and abs(parts[1] - s[1]) < 1e-13


@pytest.mark.parametrize("n", [1 << 25, 10 ** 9, 10 ** 9 + 1])
def test_fused_user_integrand_at_full_size(fgq, n):
    """The headline size: nothing is stored, so the reduction is FP64-bound; exact integrals as the check."""
    s = fgq.gausslegendre_integrate(n, "exp(x) * cos(3 * x)").cpu().numpy()
    exact = (math.e * (math.cos(3) + 3 * math.sin(3)) - (math.cos(3) - 3 * math.sin(3)) / math.e) / 10
    assert abs(s[0] - exact) < 2e-13 and abs(s[1] - 2.0) < 2e-13


def test_stored_rules_of_the_other_families(fgq):
    """dot(w, f.(x)) over stored Jacobi / Hermite / Laguerre rules against the oracle's dot product."""
    import torch
    from oracle import hermite_oracle, jacobi_oracle, laguerre_oracle
    x, w, _ = fgq.gaussjacobi_device(2000, 0.3, -0.7)
    s = fgq.rule_integrate(x, w, "exp(x) * cos(2 * x)").cpu().numpy()
    xo, wo = jacobi_oracle.gaussjacobi(2000, 0.3, -0.7)
    ref = _oracle_dot(xo, wo, lambda t: np.exp(t) * np.cos(2 * t))
    assert abs(s[0] - ref) < 2e-13 * abs(ref) and abs(s[1] - wo.sum()) < 1e-13 * wo.sum()
    x, w, _ = fgq.gausshermite_device(5000)
    s = fgq.rule_integrate(x, w, "cos(x)").cpu().numpy()          # int cos(x) exp(-x^2) = sqrt(pi) exp(-1/4)
    xo, wo = hermite_oracle.gausshermite(5000)
    assert abs(s[0] - _oracle_dot(xo, wo, np.cos)) < 1e-13 and abs(s[0] - math.sqrt(math.pi) * math.exp(-0.25)) < 1e-13
    x, w, _ = fgq.gausslaguerre_device(5000, 0.0)
    s = fgq.rule_integrate(x, w, "cos(4 * x)").cpu().numpy()      # test_gausslaguerre.jl:120: 1/17
    xo, wo = laguerre_oracle.gausslaguerre(5000, 0.0)
    assert abs(s[0] - _oracle_dot(xo, wo, lambda t: np.cos(4 * t))) < 1e-13 and abs(s[0] - 1 / 17) < 1e-13
    torch.cuda.synchronize()


def test_bad_expressions_fail_loudly(fgq):
    for bad in ("", "exp(x", "x; x", "undefined_function(x)", "x }"):
        with pytest.raises(ValueError):
            fgq.gausslegendre_integrate(1000, bad)
    with pytest.raises(ValueError):
        fgq.gausslegendre_integrate(40, "x")          # n <= 60: no fused path, integrate the stored rule instead
