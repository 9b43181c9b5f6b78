"""hermite_oracle.py — CPU ORACLE (test infrastructure, NOT product code).

numpy restatement of the asymptotic Gauss-Hermite path of FastGaussQuadrature.jl v1.3.0:
  gausshermite / unweightedgausshermite   src/gausshermite.jl:28-65
  hermite_asy                             src/gausshermite.jl:68-89
  hermpoly_asy_airy                       src/gausshermite.jl:171-250
  hermite_initialguess                    src/gausshermite.jl:252-323
  AIRY_ROOTS                              src/constants.jl:121-135
  hermite_rec, hermpoly_rec               src/gausshermite.jl:91-137   (21 <= n <= 200)
  hermite_gw                              src/gausshermite.jl:325-341  (n <= 20; LAPACK through scipy)

Arithmetic outside /root/reference: `airyai`/`airyaiprime` are SpecialFunctions.jl -> AMOS `zairy`
(version unpinned in the reference).  scipy.special.airy on a COMPLEX argument calls the same AMOS
routine, so that is what is used here (for real dtype scipy switches to cephes inside [-10, 10]).
Broadcast expressions keep the reference's association; `x .^ (2/3)` etc. are numpy `**` (libm pow).

Pin: golden values of test/test_gausshermite.jl:38-54 (tests/test_oracle_hermite.py).  Below the
1e-14 level the result depends on the Airy implementation and libm: parity unpinned there.
"""
from __future__ import annotations

import numpy as np
from scipy.special import airy

AIRY_ROOTS = np.array([
    -2.338107410459767, -4.08794944413097, -5.520559828095551, -6.786708090071759,
    -7.944133587120853, -9.022650853340981, -10.04017434155809, -11.00852430373326,
    -11.93601556323626, -12.828776752865757, -13.69148903521072])
EPS = np.finfo(np.float64).eps
PI = np.pi


def airy_ai_aip(z: np.ndarray):
    """(Ai(z), Ai'(z)) through AMOS zairy, as SpecialFunctions.airyai/airyaiprime do."""
    ai, aip, _, _ = airy(np.asarray(z, dtype=np.complex128))
    return ai.real, aip.real


def hermite_initialguess(n: int) -> np.ndarray:
    """gausshermite.jl:252-323"""
    assert n >= 20
    if n % 2:
        m = (n - 1) >> 1
        a = 0.5
    else:
        m = n >> 1
        a = -0.5
    nu = 4 * m + 2 * a + 2
    k = np.arange(1, m + 1, dtype=np.float64)

    def T(t):
        return t ** (2 / 3) * (1 + 5 / 48 * t ** (-2.0) - 5 / 36 * t ** (-4.0) + (77125 / 82944) * t ** (-6.0)
                               - 108056875 / 6967296 * t ** (-8.0) + 162375596875 / 334430208 * t ** (-10.0))

    airyrts = -T(3 * PI / 8 * (4 * k - 1))
    airyrts[:10] = AIRY_ROOTS[:10]
    x_init = np.sqrt(np.abs(
        nu + (2 ** (2 / 3)) * airyrts * nu ** (1 / 3) + (1 / 5 * 2 ** (4 / 3)) * airyrts ** 2 * nu ** (-1 / 3)
        + (11 / 35 - a ** 2 - 12 / 175) * airyrts ** 3 / nu
        + ((16 / 1575) * airyrts + (92 / 7875) * airyrts ** 4) * 2 ** (2 / 3) * nu ** (-5 / 3)
        - ((15152 / 3031875) * airyrts ** 5 + (1088 / 121275) * airyrts ** 2) * 2 ** (1 / 3) * nu ** (-7 / 3)))
    x_init_airy = x_init[::-1]

    Tnk0 = np.full(m, PI / 2)
    rhs = ((4 * m + 3) - 4 * k) / nu * PI
    for _ in range(7):
        val = Tnk0 - np.sin(Tnk0) - rhs
        dval = 1 - np.cos(Tnk0)
        Tnk0 = Tnk0 - val / dval
    tnk = np.cos(Tnk0 / 2) ** 2
    x_init_sin = np.sqrt(nu * tnk - ((tnk + 1 / 4) / (tnk - 1) ** 2 + (3 * a ** 2 - 1)) / (3 * nu))

    p = 0.4985 + EPS
    lo = int(np.floor(p * n))
    hi = int(np.ceil(p * n))
    x0 = np.concatenate([x_init_sin[:lo], x_init_airy[hi - 1:]])
    if n % 2:
        x0 = np.concatenate([[0.0], x0])[:m + 1]
    else:
        x0 = x0[:m]
    return x0


def hermpoly_asy_airy(n: int, theta: np.ndarray):
    """gausshermite.jl:171-250"""
    musq = 2 * n + 1
    cosT = np.cos(theta)
    sinT = np.sin(theta)
    sin2T = 2 * cosT * sinT
    eta = 0.5 * theta - 0.25 * sin2T
    chi = -(3 * eta / 2) ** (2 / 3)
    phi = (-chi / sinT ** 2) ** (1 / 4)
    C = 2 * np.sqrt(PI) * musq ** (1 / 6) * phi
    Airy0, Airy1 = airy_ai_aip(musq ** (2 / 3) * chi)

    a0 = 1; b0 = 1
    a1 = 15 / 144; b1 = -7 / 5 * a1
    a2 = 5 * 7 * 9 * 11 / 2 / 144 ** 2; b2 = -13 / 11 * a2
    a3 = 7 * 9 * 11 * 13 * 15 * 17 / 6 / 144 ** 3; b3 = -19 / 17 * a3

    u0 = 1
    u1 = (cosT ** 3 - 6 * cosT) / 24
    u2 = (-9 * cosT ** 4 + 249 * cosT ** 2 + 145) / 1152
    u3 = (-4042 * cosT ** 9 + 18189 * cosT ** 7 - 28287 * cosT ** 5 - 151995 * cosT ** 3 - 259290 * cosT) / 414720

    val = 1 * Airy0
    B0 = -(a0 * u1 * phi ** 6 + a1 * u0) / chi ** 2
    val = val + B0 * Airy1 / musq ** (4 / 3)
    A1 = (b0 * u2 * phi ** 12 + b1 * u1 * phi ** 6 + b2 * u0) / chi ** 3
    val = val + A1 * Airy0 / musq ** 2
    B1 = -(u3 * phi ** 18 + a1 * u2 * phi ** 12 + a2 * u1 * phi ** 6 + a3 * u0) / chi ** 5
    val = val + B1 * Airy1 / musq ** (4 / 3 + 2)
    val = C * val

    C = np.sqrt(2 * PI) * musq ** (1 / 3) / phi
    v0 = 1
    v1 = (cosT ** 3 + 6 * cosT) / 24
    v2 = (15 * cosT ** 4 - 327 * cosT ** 2 - 143) / 1152
    v3 = (259290 * cosT + 238425 * cosT ** 3 - 36387 * cosT ** 5 + 18189 * cosT ** 7 - 4042 * cosT ** 9) / 414720
    C0 = -(b0 * phi ** 6 * v1 + b1 * v0) / chi
    dval = C0 * Airy0 / musq ** (2 / 3)
    D0 = a0 * v0
    dval = dval + D0 * Airy1
    C1 = -(v3 * phi ** 18 + b1 * v2 * phi ** 12 + b2 * v1 * phi ** 6 + b3 * v0) / chi ** 4
    dval = dval + C1 * Airy0 / musq ** (2 / 3 + 2)
    D1 = (a0 * v2 * phi ** 12 + a1 * v1 * phi ** 6 + a2 * v0) / chi ** 3
    dval = dval + D1 * Airy1 / musq ** 2
    dval = C * dval
    return val, dval


def hermite_asy(n: int, info: dict | None = None):
    """gausshermite.jl:68-89: half-range nodes (x >= 0) and unnormalised weights."""
    x0 = hermite_initialguess(n)
    t0 = x0 / np.sqrt(2 * n + 1)
    theta = np.arccos(t0)
    sweeps = 0
    for _ in range(20):
        val, dval = hermpoly_asy_airy(n, theta)
        dtheta = -val / (np.sqrt(2) * np.sqrt(2 * n + 1) * dval * np.sin(theta))
        theta = theta - dtheta
        sweeps += 1
        if np.max(np.abs(dtheta)) < np.sqrt(EPS) / 10:
            break
    t0 = np.cos(theta)
    x = np.sqrt(2 * n + 1) * t0
    w = x * val + np.sqrt(2) * dval
    w = 1 / w ** 2
    if info is not None:
        info["sweeps"] = sweeps
        info["theta"] = theta
    return x, w


def hermpoly_rec(n: int, x0: float):
    """gausshermite.jl:113-137 (scalar)."""
    import math
    w = math.exp(-x0 ** 2 / (4 * n))
    wc = 0
    Hold = 1.0
    H = x0
    for k in range(1, n):
        Hold, H = H, (x0 * H / math.sqrt(k + 1.0) - Hold / math.sqrt((k + 1.0) / k))
        while abs(H) >= 100 and wc < n:
            H *= w
            Hold *= w
            wc += 1
    for _ in range(wc + 1, n + 1):
        H *= w
        Hold *= w
    return H, -x0 * H + math.sqrt(float(n)) * Hold


def hermite_rec(n: int):
    """gausshermite.jl:91-111 (Float64: at most 10 Newton steps)."""
    x0 = hermite_initialguess(n) * np.sqrt(2.0)
    val = None
    for _ in range(10):
        val = np.array([hermpoly_rec(n, float(v)) for v in x0])
        with np.errstate(all="ignore"):
            dx = val[:, 0] / val[:, 1]
        dx[np.isnan(dx)] = 0
        x0 = x0 - dx
        if np.max(np.abs(dx)) < np.sqrt(EPS):
            break
    x0 = x0 / np.sqrt(2.0)
    return x0, 1 / val[:, 1] ** 2


def hermite_gw(n: int):
    """gausshermite.jl:325-341 (LAPACK symmetric-tridiagonal eigen-solve, as in the reference)."""
    from scipy.linalg import eigh_tridiagonal
    beta = np.sqrt(np.arange(1, n, dtype=np.float64) / 2)
    D, V = eigh_tridiagonal(np.zeros(n), beta)
    ind = np.argsort(D)
    x = D[ind]
    w = np.sqrt(PI) * V[0, ind] ** 2
    i0 = n // 2
    x = x[i0:]
    w = w[i0:]
    return x, np.exp(x ** 2) * w


def unweightedgausshermite(n: int, info: dict | None = None):
    """gausshermite.jl:35-65."""
    if n < 0:
        raise ValueError("DomainError: Input n must be a non-negative integer")
    if n == 0:
        return np.zeros(0), np.zeros(0)
    if n == 1:
        xh, wh = np.array([0.0]), np.array([np.sqrt(PI)])
    elif n <= 20:
        xh, wh = hermite_gw(n)
    elif n <= 200:
        xh, wh = hermite_rec(n)
    else:
        xh, wh = hermite_asy(n, info)
    if n % 2:
        w = np.concatenate([wh[::-1], wh[1:]])
        x = np.concatenate([-xh[::-1], xh[1:]])
    else:
        w = np.concatenate([wh[::-1], wh])
        x = np.concatenate([-xh[::-1], xh])
    w = w * (np.sqrt(PI) / np.sum(np.exp(-x ** 2) * w))
    return x, w


def gausshermite(n: int, normalize: bool = False, info: dict | None = None):
    """gausshermite.jl:28-33"""
    x, w = unweightedgausshermite(n, info)
    w = w * np.exp(-x ** 2)
    if normalize:
        return np.sqrt(2.0) * x, w / np.sqrt(PI)
    return x, w
