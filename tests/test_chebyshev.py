"""Gauss-Chebyshev closed forms: oracle identities on CPU, GPU parity through the C ABI
(reference test/test_gausschebyshev.jl)."""
import numpy as np
import pytest

from oracle import chebyshev_oracle as co


def _identities(rule):
    n = 10
    x, w = rule(1, n); assert abs(np.dot(x ** 2, w) - np.pi / 2) < 1e-14 and abs(np.dot(x ** 3, w)) < 1e-15
    x, w = rule(2, n); assert abs(np.dot(x ** 2, w) - np.pi / 8) < 1e-14 and abs(np.dot(x ** 3, w)) < 1e-15
    x, w = rule(3, n); assert abs(np.dot(x ** 2, w) - np.pi / 2) < 1e-14 and abs(np.dot(x ** 3, w) - 3 * np.pi / 8) < 1e-14
    x, w = rule(4, n); assert abs(np.dot(x ** 2, w) - np.pi / 2) < 1e-14 and abs(np.dot(x ** 3, w) + 3 * np.pi / 8) < 1e-14


def test_oracle_identities():
    _identities(co.gausschebyshev)
    with pytest.raises(ValueError):
        co.gausschebyshev(1, -1)


@pytest.mark.gpu
def test_gpu_identities_and_errors(fgq):
    rules = {1: fgq.gausschebyshevt, 2: fgq.gausschebyshevu, 3: fgq.gausschebyshevv, 4: fgq.gausschebyshevw}
    _identities(lambda kind, n: rules[kind](n))
    for f in rules.values():
        with pytest.raises(fgq.DomainError):   # test_gausschebyshev.jl:4-7
            f(-1)
        x, w = f(0)
        assert len(x) == 0
    assert fgq.lib().fgq_gausschebyshev_host(5, 0, 0, 0) == 3   # ArgumentError (:40)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [1, 2, 10, 42, 1001, 100000, 2000003])
def test_gpu_parity(fgq, oracle_mod, kind, n):
    rules = {1: fgq.gausschebyshevt, 2: fgq.gausschebyshevu, 3: fgq.gausschebyshevv, 4: fgq.gausschebyshevw}
    x, w = rules[kind](n)
    xo, wo = co.gausschebyshev(kind, n)
    nz = xo != 0
    assert np.abs(x[~nz]).max(initial=0.0) <= 1e-16
    assert oracle_mod.ulp_diff(x[nz], xo[nz]).max(initial=0.0) <= 2
    assert oracle_mod.ulp_diff(w, wo).max() <= 6   # the weight squares a sinpi/cospi (<= 2 ulp on the device)
    assert np.all(np.diff(x) > 0) or n > 10 ** 7
    # the Jacobi routes (gaussjacobi.jl:30-39)
    a, b = {1: (-0.5, -0.5), 2: (0.5, 0.5), 3: (-0.5, 0.5), 4: (0.5, -0.5)}[kind]
    xj, wj = fgq.gaussjacobi(n, a, b)
    assert np.array_equal(xj, x) and np.array_equal(wj, w)


@pytest.mark.gpu
def test_deprecated_aliases(fgq):
    """gausschebyshev(n, kind) (src/gausschebyshev.jl:117-144) and besselroots (src/besselroots.jl:247) keep working,
    with a deprecation warning, exactly as in the reference."""
    rules = {1: fgq.gausschebyshevt, 2: fgq.gausschebyshevu, 3: fgq.gausschebyshevv, 4: fgq.gausschebyshevw}
    for kind, f in rules.items():
        with pytest.warns(DeprecationWarning):
            x, w = fgq.gausschebyshev(17, kind)
        xs, ws = f(17)
        assert np.array_equal(x, xs) and np.array_equal(w, ws)
    with pytest.warns(DeprecationWarning):
        x, _ = fgq.gausschebyshev(5)          # kind defaults to 1
    assert np.array_equal(x, fgq.gausschebyshevt(5)[0])
    with pytest.raises(ValueError):           # ArgumentError in the reference
        fgq.gausschebyshev(5, 7)
    with pytest.raises(fgq.DomainError):      # checked before the kind, as in the reference
        fgq.gausschebyshev(-1, 7)
    with pytest.warns(DeprecationWarning):
        r = fgq.besselroots(0.3, 12)
    assert np.array_equal(r, fgq.approx_besselroots(0.3, 12))
