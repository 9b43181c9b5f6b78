"""chebyshev_oracle.py — CPU ORACLE (test infrastructure, NOT product code).

numpy restatement of src/gausschebyshev.jl:22-26, 51-55, 80-84, 109-113 (closed forms).  Julia's
cospi/sinpi reduce the argument exactly; the same reduction is done here before calling libm.
The reference pins these rules only through integral identities (test/test_gausschebyshev.jl): at the
ulp level the parity is UNPINNED (Julia cannot run in this image)."""
import numpy as np

PI = np.pi
_LD = np.longdouble
_PI_LD = _LD("3.14159265358979323846264338327950288")


def cospi(x):
    """cos(pi x) for 0 <= x <= 1: exact argument reduction, evaluation in x87 extended precision
    (64-bit mantissa), rounded once to double — i.e. correctly rounded up to rare ties, which is what
    Julia's cospi delivers (< 1 ulp)."""
    x = np.asarray(x, dtype=np.float64).astype(_LD)
    r = np.where(x <= 0.25, np.cos(_PI_LD * x),
                 np.where(x < 0.75, np.sin(_PI_LD * (_LD(0.5) - x)), -np.cos(_PI_LD * (1 - x))))
    return r.astype(np.float64)


def sinpi(x):
    x = np.asarray(x, dtype=np.float64).astype(_LD)
    r = np.where(x <= 0.25, np.sin(_PI_LD * x),
                 np.where(x < 0.75, np.cos(_PI_LD * (_LD(0.5) - x)), np.sin(_PI_LD * (1 - x))))
    return r.astype(np.float64)


def gausschebyshev(kind: int, n: int):
    if n < 0:
        raise ValueError("DomainError: Input n must be a non-negative integer")
    k = np.arange(n, 0, -1, dtype=np.float64)
    if kind == 1:
        return cospi((2 * k - 1) / (2 * n)), np.full(n, PI / n) if n else np.zeros(0)
    if kind == 2:
        return cospi(k / (n + 1)), PI / (n + 1) * sinpi(k / (n + 1)) ** 2
    if kind == 3:
        return cospi((2 * k - 1) / (2 * n + 1)), 4 * PI / (2 * n + 1) * cospi((2 * k - 1) / (4 * n + 2)) ** 2
    if kind == 4:
        return cospi(2 * k / (2 * n + 1)), 4 * PI / (2 * n + 1) * sinpi(k / (2 * n + 1)) ** 2
    raise ValueError("ArgumentError: Chebyshev kind should be 1, 2, 3, or 4")
