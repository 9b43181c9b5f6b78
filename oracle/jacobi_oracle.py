"""jacobi_oracle.py — CPU ORACLE (test infrastructure, NOT product code).

numpy restatement of the asymptotic Gauss-Jacobi path of FastGaussQuadrature.jl v1.3.0 and the
Radau / Lobatto epilogues built on it:
  gaussjacobi dispatch          src/gaussjacobi.jl:23-55
  weightsConstant, jacobi_asy   src/gaussjacobi.jl:157-194
  asy1, feval_asy1              src/gaussjacobi.jl:196-361   (interior, Hale-Townsend)
  asy2, feval_asy2              src/gaussjacobi.jl:363-468   (boundary, Bessel-type)
  asy2_higherterms              src/gaussjacobi.jl:470-568
  bary, bessel_taylor, nck_mat  src/gaussjacobi.jl:601-667
  approx_besselroots, McMahon, piessens   src/besselroots.jl:23-165
  gaussradau(n), gausslobatto(n)          src/gaussradau.jl:31-49, src/gausslobatto.jl:31-52
jacobi_rec / HalfRec / innerjacobi_rec! (n <= 100)   src/gaussjacobi.jl:61-155
The Golub-Welsch eigen-solve (max(a,b) >= 5) is out of scope.

Arithmetic outside /root/reference: `besselj(nu, x)` is SpecialFunctions.jl -> AMOS zbesj;
scipy.special.jv is the same AMOS routine.  The dense products in feval_asy1 are BLAS gemv in the
reference; numpy's `@` is used here (also BLAS).  Pin: the golden values of
test/test_gaussjacobi.jl:149-248, test_gaussradau.jl:21-29, test_gausslobatto.jl:27-36
(tests/test_oracle_jacobi.py).  Julia cannot run in this image: below the 1e-14 absolute level of those
tests (libm, AMOS and BLAS rounding) the parity is UNPINNED.
"""
from __future__ import annotations

import numpy as np
from scipy.special import jv

from .jacobi_tables import CUMSUMMAT_10, DIFFMAT_10, PIESSENS_C

EPS = np.finfo(np.float64).eps
PI = np.pi

BESSELJ0_ROOTS = np.array([
    2.4048255576957728, 5.5200781102863106, 8.6537279129110122, 11.791534439014281,
    14.930917708487785, 18.071063967910922, 21.211636629879258, 24.352471530749302,
    27.493479132040254, 30.634606468431975, 33.775820213573568, 36.917098353664044,
    40.058425764628239, 43.199791713176730, 46.341188371661814, 49.482609897397817,
    52.624051841114996, 55.765510755019979, 58.906983926080942, 62.048469190227170])

STIRLING_G = np.array([
    1, 1 / 12, 1 / 288, -139 / 51840, -571 / 2488320, 163879 / 209018880,
    5246819 / 75246796800, -534703531 / 902961561600,
    -4483131259 / 86684309913600, 432261921612371 / 514904800886784000])


def _stirling_f(z):
    return float(np.dot(STIRLING_G, np.concatenate([[1.0], np.cumprod(np.ones(9) / z)])))


# ---- besselroots.jl:23-165 ---------------------------------------------------------------------

def mcmahon(nu: float, k: int) -> float:
    mu = 4 * nu ** 2
    a1 = 1 / 8
    a3 = (7 * mu - 31) / 384
    a5 = 4 * (3779 + mu * (-982 + 83 * mu)) / 61440
    a7 = 6 * (-6277237 + mu * (1585743 + mu * (-153855 + 6949 * mu))) / 20643840
    a9 = 144 * (2092163573 + mu * (-512062548 + mu * (48010494 + mu * (-2479316 + 70197 * mu)))) / 11890851840
    a11 = 720 * (-8249725736393 + mu * (1982611456181 + mu * (-179289628602 + mu * (8903961290 + mu * (
        -287149133 + mu * 5592657))))) / 10463949619200
    a13 = 576 * (423748443625564327 + mu * (-100847472093088506 + mu * (8929489333108377 + mu * (
        -426353946885548 + mu * (13172003634537 + mu * (-291245357370 + mu * 4148944183)))))) / 13059009124761600
    b = 0.25 * (2 * nu + 4 * k - 1) * PI
    b2 = b ** 2
    return b - (mu - 1) * (((((((a13 / b2 + a11) / b2 + a9) / b2 + a7) / b2 + a5) / b2 + a3) / b2 + a1) / b)


def piessens(nu: float) -> np.ndarray:
    pt = (nu - 2) / 3
    T = np.empty(30)
    T[0] = 1.0
    T[1] = pt
    for i in range(2, 30):
        T[i] = 2 * pt * T[i - 1] - T[i - 2]
    y = PIESSENS_C.T @ T
    y[0] = y[0] * np.sqrt(nu + 1)
    return y


def approx_besselroots(nu: float, n: int) -> np.ndarray:
    x = np.zeros(n)
    if nu == 0:
        for k in range(1, min(n, 20) + 1):
            x[k - 1] = BESSELJ0_ROOTS[k - 1]
        for k in range(min(n, 20) + 1, n + 1):
            x[k - 1] = mcmahon(nu, k)
    elif -1 <= nu <= 5:
        c = piessens(nu)
        for k in range(1, min(n, 6) + 1):
            x[k - 1] = c[k - 1]
        for k in range(min(n, 6) + 1, n + 1):
            x[k - 1] = mcmahon(nu, k)
    elif 5 < nu:
        for k in range(1, n + 1):
            x[k - 1] = mcmahon(nu, k)
    return x


# ---- gaussjacobi.jl:157-168 --------------------------------------------------------------------

def weights_constant(n: int, a: float, b: float) -> float:
    M = min(20, n - 1)
    C = 1.0
    p = -a * b / n
    for m in range(1, M + 1):
        C += p
        p *= -(m + a) * (m + b) / (m + 1) / (n - m)
        if abs(p / C) < EPS / 100:
            break
    return 2 ** (a + b + 1) * C


# ---- gaussjacobi.jl:254-361 --------------------------------------------------------------------

def _phi_tables(n, a, b, M=20):
    """PHI, PHI2 of feval_asy1 (:274-294): parameter-only M x M tables."""
    def build(shift1, shift2):
        vec = np.array([(a + j - 0.5) * (-a + j - 0.5) / (2 * n + a + b + j + shift1) / j for j in range(1, M)])
        P1 = np.concatenate([[1.0], np.cumprod(vec)])
        P1[2::4] = -P1[2::4]
        P1[3::4] = -P1[3::4]
        P2 = np.eye(M)
        for l in range(1, M + 1):
            cnt = M - l - 2
            if cnt >= 1:
                v = np.array([(b + j - 0.5) * (-b + j - 0.5) / (2 * n + a + b + j + l + shift2) / j
                              for j in range(1, cnt + 1)])
                P2[l - 1, l:M - 2] = np.cumprod(v)
        return P1[:, None] * P2
    return build(1, 0), build(-1, -2)


def feval_asy1_constants(n, a, b):
    """C, C2 of feval_asy1 (:319-350)."""
    dsa = a ** 2 / (2 * n)
    dsb = b ** 2 / (2 * n)
    dsab = (a + b) ** 2 / (4 * n)
    ds = dsa + dsb - dsab
    s = ds
    i = 1
    dsold = ds
    def ratio(ds, s):  # Julia: abs(0/0) = NaN, and NaN comparisons are false (alpha = beta = 1)
        return abs(ds / s) if s != 0 else float("nan")
    while ratio(ds, s) + dsold > EPS / 10:
        dsold = ratio(ds, s)
        i += 1
        tmp = -(i - 1) / (i + 1) / n
        dsa = tmp * dsa * a
        dsb = tmp * dsb * b
        dsab = tmp * dsab * (a + b) / 2
        ds = dsa + dsb - dsab
        s = s + ds
    p2 = np.exp(s) * np.sqrt(2 * PI * (n + a) * (n + b) / (2 * n + a + b)) / (2 * n + a + b + 1)
    C = 2 * p2 * (_stirling_f(n + a) * _stirling_f(n + b) / _stirling_f(2 * n + a + b)) / PI
    C2 = C * (a + b + 2 * n) * (a + b + 1 + 2 * n) / (4 * (a + n) * (b + n))
    return C, C2


def feval_asy1(n, a, b, t, idx, info=None):
    """:254-361.  idx = slice of the interior index set (0-based)."""
    M = 20
    N = len(t)
    mm = np.arange(1, M + 1, dtype=np.float64)
    A = ((2 * n + a + b) + mm)[:, None] * t[None, :] / 2 - (a + 1 / 2) * PI / 2
    cosA = np.cos(A)
    sinA = np.sin(A)
    sinT_ = np.sin(t)
    cosT_ = np.cos(t)
    cosA2 = cosA * cosT_[None, :] + sinA * sinT_[None, :]
    sinA2 = sinA * cosT_[None, :] - cosA * sinT_[None, :]
    del A
    csc = (1.0 / np.sin(t / 2)) / 2
    sinT = np.concatenate([np.ones((N, 1)), np.cumprod(np.repeat(csc[:, None], M - 1, axis=1), axis=1)], axis=1)
    secT = (1.0 / np.cos(t / 2)) / 2
    PHI, PHI2 = _phi_tables(n, a, b, M)

    S = np.zeros(N)
    S2 = np.zeros(N)
    used = M
    for m in range(1, M + 1):
        lo = np.arange(0, m, 2)      # l = 1:2:m
        le = np.arange(1, m, 2)      # l = 2:2:m
        dS1 = (sinT[:, lo] @ PHI[lo, m - 1]) * cosA[m - 1, :]
        dS12 = (sinT[:, lo] @ PHI2[lo, m - 1]) * cosA2[m - 1, :]
        if len(le):
            dS2 = (sinT[:, le] @ PHI[le, m - 1]) * sinA[m - 1, :]
            dS22 = (sinT[:, le] @ PHI2[le, m - 1]) * sinA2[m - 1, :]
        else:
            dS2 = np.zeros(N)
            dS22 = np.zeros(N)
        if m - 1 > 10 and np.max(np.abs(dS1[idx] + dS2[idx])) < EPS / 100:
            used = m - 1
            break
        S += dS1
        S += dS2
        S2 += dS12
        S2 += dS22
        sinT[:, :m] *= secT[:, None]
    if info is not None:
        info.setdefault("terms", []).append(used)

    C, C2 = feval_asy1_constants(n, a, b)
    vals = C * S
    ders = (n * ((a - b) - (2 * n + a + b) * cosT_) * vals + (2 * (n + a) * (n + b) * C2) * S2) / (2 * n + a + b) / sinT_
    denom = 1.0 / (np.sin(np.abs(t) / 2) ** (a + 0.5) * np.cos(t / 2) ** (b + 0.5))
    vals = vals * denom
    ders = ders * denom
    return vals, ders


def asy1(n, a, b, nbdy, info=None):
    """:196-247"""
    j = np.arange(n, 0, -1, dtype=np.float64)
    K = PI * (2 * j + a - 0.5) / (2 * n + a + b + 1)
    tt = K + (1 / (2 * n + a + b + 1) ** 2) * ((0.25 - a ** 2) * (1.0 / np.tan(K / 2)) - (0.25 - b ** 2) * np.tan(K / 2))

    t = tt[tt <= PI / 2]
    mint = t[len(t) - nbdy]
    first = int(np.argmax(t < mint)) if np.any(t < mint) else None
    idx = slice(0, max(first, 1))           # 1:max(findfirst(t .< mint) - 1, 1)
    sweeps_r = 0
    for _ in range(10):
        vals, ders = feval_asy1(n, a, b, t, idx, info)
        dt = vals / ders
        t = t + dt
        sweeps_r += 1
        if np.max(np.abs(dt[idx])) < np.sqrt(EPS) / 100:
            break
    vals, ders = feval_asy1(n, a, b, t, idx, info)
    t = t + vals / ders
    x_right = np.cos(t)
    w_right = 1.0 / ders ** 2

    a, b = b, a
    t = PI - tt[: n - len(x_right)]
    mint = t[nbdy - 1]
    first = int(np.argmax(t > mint))
    idx = slice(max(first, 0), len(t))      # max(findfirst(t .> mint), 1):length(t)
    sweeps_l = 0
    for _ in range(10):
        vals, ders = feval_asy1(n, a, b, t, idx, info)
        dt = vals / ders
        t = t + dt
        sweeps_l += 1
        if np.max(np.abs(dt[idx])) < np.sqrt(EPS) / 100:
            break
    vals, ders = feval_asy1(n, a, b, t, idx, info)
    t = t + vals / ders
    x_left = np.cos(t)
    w_left = 1.0 / ders ** 2
    if info is not None:
        info["sweeps"] = (sweeps_r, sweeps_l)
        info["n_right"] = len(x_right)
    return np.concatenate([-x_left, x_right]), np.concatenate([w_left, w_right])


# ---- the same interior solve evaluated in node chunks (full-size parity at n = 1e7) ----------------
#
# feval_asy1 builds ~10 dense 20 x N arrays (8 GB per call at N = 5e6, as the reference does).  Every node is
# independent EXCEPT for two global, data-dependent decisions: the series truncation index (the first m > 11
# whose interior inf-norm is below eps/100, :309) and the Newton stopping rule (:209-218).  The chunked form
# reproduces both exactly with two passes per evaluation: pass 1 collects the per-term interior norms over all
# chunks (the norm of term m does not depend on where the sum is cut), pass 2 evaluates every chunk with the
# truncation index fixed.  tests/test_oracle_jacobi.py checks it against asy1 itself.

def _feval_asy1_chunk(n, a, b, t, interior, used=None):
    """One chunk of feval_asy1.  used=None: returns the 20 interior inf-norms |dS1 + dS2| of this chunk
    (0 where the chunk has no interior node); otherwise (vals, ders) with the series cut at `used` terms."""
    M = 20
    N = len(t)
    mm = np.arange(1, M + 1, dtype=np.float64)
    A = ((2 * n + a + b) + mm)[:, None] * t[None, :] / 2 - (a + 1 / 2) * PI / 2
    cosA = np.cos(A)
    sinA = np.sin(A)
    sinT_ = np.sin(t)
    cosT_ = np.cos(t)
    cosA2 = cosA * cosT_[None, :] + sinA * sinT_[None, :]
    sinA2 = sinA * cosT_[None, :] - cosA * sinT_[None, :]
    del A
    csc = (1.0 / np.sin(t / 2)) / 2
    sinT = np.concatenate([np.ones((N, 1)), np.cumprod(np.repeat(csc[:, None], M - 1, axis=1), axis=1)], axis=1)
    secT = (1.0 / np.cos(t / 2)) / 2
    PHI, PHI2 = _phi_tables(n, a, b, M)
    S = np.zeros(N)
    S2 = np.zeros(N)
    norms = np.zeros(M)
    for m in range(1, M + 1):
        if used is not None and m - 1 == used:
            break
        lo = np.arange(0, m, 2)
        le = np.arange(1, m, 2)
        dS1 = (sinT[:, lo] @ PHI[lo, m - 1]) * cosA[m - 1, :]
        dS12 = (sinT[:, lo] @ PHI2[lo, m - 1]) * cosA2[m - 1, :]
        if len(le):
            dS2 = (sinT[:, le] @ PHI[le, m - 1]) * sinA[m - 1, :]
            dS22 = (sinT[:, le] @ PHI2[le, m - 1]) * sinA2[m - 1, :]
        else:
            dS2 = np.zeros(N)
            dS22 = np.zeros(N)
        if used is None:
            if interior.any():
                norms[m - 1] = np.max(np.abs(dS1[interior] + dS2[interior]))
        else:
            S += dS1
            S += dS2
            S2 += dS12
            S2 += dS22
        sinT[:, :m] *= secT[:, None]
    if used is None:
        return norms
    C, C2 = feval_asy1_constants(n, a, b)
    vals = C * S
    ders = (n * ((a - b) - (2 * n + a + b) * cosT_) * vals + (2 * (n + a) * (n + b) * C2) * S2) / (2 * n + a + b) / sinT_
    denom = 1.0 / (np.sin(np.abs(t) / 2) ** (a + 0.5) * np.cos(t / 2) ** (b + 0.5))
    return vals * denom, ders * denom


def _feval_asy1_chunked(n, a, b, t, i0, i1, chunk, info=None):
    """feval_asy1 over all of t, interior = indices [i0, i1), in chunks."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    N = len(t)
    starts = list(range(0, N, chunk))

    def norm_job(c0):
        c1 = min(c0 + chunk, N)
        interior = np.zeros(c1 - c0, dtype=bool)
        interior[max(i0 - c0, 0):max(min(i1, c1) - c0, 0)] = True
        return _feval_asy1_chunk(n, a, b, t[c0:c1], interior)

    # chunks are independent (numpy releases the GIL inside its loops): spread them over the host cores
    pool = ThreadPoolExecutor(max_workers=min(len(starts), os.cpu_count() or 1))
    norms = np.zeros(20)
    for r in pool.map(norm_job, starts):
        norms = np.maximum(norms, r)
    used = 20
    for m in range(1, 21):
        if m - 1 > 10 and norms[m - 1] < EPS / 100:
            used = m - 1
            break
    if info is not None:
        info.setdefault("terms", []).append(used)
    vals = np.empty(N)
    ders = np.empty(N)

    def eval_job(c0):
        c1 = min(c0 + chunk, N)
        vals[c0:c1], ders[c0:c1] = _feval_asy1_chunk(n, a, b, t[c0:c1], None, used)

    list(pool.map(eval_job, starts))
    pool.shutdown()
    return vals, ders


def asy1_chunked(n, a, b, nbdy, chunk=250_000, info=None):
    """asy1 (:196-247) with every feval_asy1 evaluated in node chunks: same global decisions, bounded memory."""
    j = np.arange(n, 0, -1, dtype=np.float64)
    K = PI * (2 * j + a - 0.5) / (2 * n + a + b + 1)
    tt = K + (1 / (2 * n + a + b + 1) ** 2) * ((0.25 - a ** 2) * (1.0 / np.tan(K / 2)) - (0.25 - b ** 2) * np.tan(K / 2))
    del j, K

    def solve(a, b, t, i0, i1):
        sweeps = 0
        for _ in range(10):
            vals, ders = _feval_asy1_chunked(n, a, b, t, i0, i1, chunk, info)
            dt = vals / ders
            t = t + dt
            sweeps += 1
            if np.max(np.abs(dt[i0:i1])) < np.sqrt(EPS) / 100:
                break
        vals, ders = _feval_asy1_chunked(n, a, b, t, i0, i1, chunk, info)
        t = t + vals / ders
        return np.cos(t), 1.0 / ders ** 2, sweeps

    t = tt[tt <= PI / 2]
    mint = t[len(t) - nbdy]
    first = int(np.argmax(t < mint)) if np.any(t < mint) else None
    x_right, w_right, sweeps_r = solve(a, b, t, 0, max(first, 1))
    t = PI - tt[: n - len(x_right)]
    mint = t[nbdy - 1]
    first = int(np.argmax(t > mint))
    x_left, w_left, sweeps_l = solve(b, a, t, max(first, 0), len(t))
    if info is not None:
        info["sweeps"] = (sweeps_r, sweeps_l)
        info["n_right"] = len(x_right)
    return np.concatenate([-x_left, x_right]), np.concatenate([w_left, w_right])


def jacobi_asy_chunked(n, a, b, chunk=250_000, info=None):
    """jacobi_asy (:170-194) on top of asy1_chunked."""
    nbdy = 10
    x, w = asy1_chunked(n, a, b, nbdy, chunk, info)
    xb, wb = asy2(n, a, b, nbdy)
    x[n - nbdy:] = xb
    w[n - nbdy:] = wb
    if a != b:
        xb, wb = asy2(n, b, a, nbdy)
    x[:nbdy] = -xb[::-1]
    w[:nbdy] = wb[::-1]
    w = w * weights_constant(n, a, b)
    return x, w


# ---- boundary: gaussjacobi.jl:363-667 ------------------------------------------------------------

def nck_mat(n):
    M = np.zeros((n - 1, n), dtype=np.int64)
    M[:, 0] = 1
    M[0, 1] = 1
    for i in range(2, n):
        for j in range(2, i + 2):
            M[i - 1, j - 1] = M[i - 2, j - 2] + M[i - 2, j - 1]
    return M


def bessel_taylor(t, z, a):
    """:624-654: J_a(z + t) by Taylor expansion about z."""
    npts = len(t)
    kmax = min(int(np.ceil(abs(np.log(EPS) / np.log(np.max(np.abs(t)))))), 30)
    H = t[None, :] ** np.arange(0, kmax + 1)[:, None]
    nu = np.arange(-kmax, kmax + 1)
    Bjk = jv(a + nu[None, :], z[:, None] * np.ones((1, 2 * kmax + 1)))
    # nck_mat(floor(1.25 kmax)) holds binomials C(k, l); for kmax <= 3 (boundary angles below ~6e-6,
    # i.e. n > ~5e6) that matrix is too small and the reference raises BoundsError at :641 — the
    # same binomials are used here so that the formula simply continues.
    import math
    nck = np.array([[math.comb(k, l) for l in range(kmax + 2)] for k in range(1, kmax + 1)], dtype=np.int64)
    AA = np.zeros((npts, kmax + 1))
    AA[:, 0] = Bjk[:, kmax]
    fact = 1
    for k in range(1, kmax + 1):
        sgn = 1
        for l in range(0, k + 1):
            AA[:, k] = AA[:, k] + sgn * nck[k - 1, l] * Bjk[:, kmax + 2 * l - k]
            sgn = -sgn
        fact = k * fact
        AA[:, k] /= 2.0 ** k * fact
    return np.array([np.dot(AA[k, :], H[:, k]) for k in range(npts)])


def bary(x, fvals, xk, vk):
    fx = np.zeros(len(x))
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(len(x)):
            xx = vk / (x[j] - xk)
            fx[j] = np.dot(xx, fvals) / np.sum(xx)
    for k in np.nonzero(np.isnan(fx))[0]:
        hit = np.nonzero(x[k] == xk)[0]
        if len(hit):
            fx[k] = fvals[hit[0]]
    return fx


def asy2_higherterms(a, b, theta):
    """:470-568"""
    A = (0.25 - a ** 2)
    B = (0.25 - b ** 2)
    c = max(float(np.max(theta)), 0.5)
    N = 10
    t = 0.5 * c * (np.sin(PI * np.arange(-(N - 1), N, 2) / (2 * (N - 1))) + 1)
    v = np.concatenate([[0.5], np.ones(N - 1)])
    v[1::2] = -1
    v[-1] *= 0.5
    with np.errstate(divide="ignore", invalid="ignore"):
        g = A * (1 / np.tan(t / 2) - 2 / t) - B * np.tan(t / 2)
        gp = A * (2 / t ** 2 - 0.5 * (1 / np.sin(t / 2)) ** 2) - 0.5 * (0.25 - b ** 2) * (1 / np.cos(t / 2)) ** 2
        gpp = A * (-4 / t ** 3 + 0.25 * np.sin(t) * (1 / np.sin(t / 2)) ** 4) - 4 * B * np.sin(t / 2) ** 4 * (1 / np.sin(t)) ** 3
        g[0] = 0
        gp[0] = -A / 6 - 0.5 * B
        gpp[0] = 0
        B0 = 0.25 * g / t
        B0p = 0.25 * (gp / t - g / t ** 2)
        B0[0] = 0.25 * (-A / 6 - 0.5 * B)
        B0p[0] = 0
        A10 = a * (A + 3 * B) / 24
        A1 = 0.125 * gp - (1 + 2 * a) / 2 * B0 - g ** 2 / 32 - A10
        A1p = 0.125 * gpp - (1 + 2 * a) / 2 * B0p - gp * g / 16
        A1p_t = A1p / t
        A1p_t[0] = -A / 720 - A ** 2 / 576 - A * B / 96 - B ** 2 / 64 - B / 48 + a * (A / 720 + B / 48)
        fcos = B / (2 * np.cos(t / 2)) ** 2
        f = -A * (1 / 12 + t ** 2 / 240 + t ** 4 / 6048 + t ** 6 / 172800 + t ** 8 / 5322240
                  + 691 * t ** 10 / 118879488000 + t ** 12 / 5748019200
                  + 3617 * t ** 14 / 711374856192000 + 43867 * t ** 16 / 300534953951232000)
        big = t > 0.5
        f[big] = A * (1 / t[big] ** 2 - 1 / (2 * np.sin(t[big] / 2)) ** 2)
        f = f - fcos
        C = (0.5 * c) * CUMSUMMAT_10
        D = (2 / c) * DIFFMAT_10
        I = C @ A1p_t
        J = C @ (f * A1)
        tB1 = -0.5 * A1p - (0.5 + a) * I + 0.5 * J
        tB1[0] = 0
        B1 = tB1 / t
        B1[0] = (A / 720 + A ** 2 / 576 + A * B / 96 + B ** 2 / 64 + B / 48
                 + a * (A ** 2 / 576 + B ** 2 / 64 + A * B / 96) - a ** 2 * (A / 720 + B / 48))
        K = C @ (f * tB1)
        A2 = 0.5 * (D @ tB1) - (0.5 + a) * B1 - 0.5 * K
        A2 = A2 - A2[0]
        A2p = D @ A2
        A2p = A2p - A2p[0]
        A2p_t = A2p / t
        w = PI / 2 - t[1:]
        w[1::2] = -w[1::2]
        w[-1] = 0.5 * w[-1]
        A2p_t[0] = np.sum(w * A2p_t[1:]) / np.sum(w)
        tB2 = -0.5 * A2p - (0.5 + a) * (C @ A2p_t) + 0.5 * (C @ (f * A2))
        B2 = tB2 / t
        B2[0] = np.sum(w * B2[1:]) / np.sum(w)
        K = C @ (f * tB2)
        A3 = 0.5 * (D @ tB2) - (0.5 + a) * B2 - 0.5 * K
        A3 = A3 - A3[0]
    return (lambda th: bary(th, tB1, t, v), lambda th: bary(th, A2, t, v),
            lambda th: bary(th, tB2, t, v), lambda th: bary(th, A3, t, v))


def feval_asy2(n, a, b, t, higherterms):
    """:405-468"""
    rho = n + 0.5 * (a + b + 1)
    rho2 = n + 0.5 * (a + b - 1)
    A = (0.25 - a ** 2)
    B = (0.25 - b ** 2)
    Ja = jv(a, rho * t)
    Jb = jv(a + 1, rho * t)
    Jbb = jv(a + 1, rho2 * t)
    Jab = bessel_taylor(-t, rho * t, a)
    gt = A * (1 / np.tan(t / 2) - (2 / t)) - B * np.tan(t / 2)
    gtdx = A * (2 / t ** 2 - 0.5 * (1 / np.sin(t / 2)) ** 2) - 0.5 * B * (1 / np.cos(t / 2)) ** 2
    tB0 = 0.25 * gt
    A10 = a * (A + 3 * B) / 24
    A1 = gtdx / 8 - (1 + 2 * a) / 8 * gt / t - gt ** 2 / 32 - A10
    tB1, A2, tB2, A3 = higherterms
    tB1t = tB1(t)
    A2t = A2(t)
    vals = Ja + Jb * tB0 / rho + Ja * A1 / rho ** 2 + Jb * tB1t / rho ** 3 + Ja * A2t / rho ** 4
    vals2 = Jab + Jbb * tB0 / rho2 + Jab * A1 / rho2 ** 2 + Jbb * tB1t / rho2 ** 3 + Jab * A2t / rho2 ** 4
    tB2t = tB2(t)
    A3t = A3(t)
    vals = vals + (Jb * tB2t / rho ** 5 + Ja * A3t / rho ** 6)
    vals2 = vals2 + (Jbb * tB2t / rho2 ** 5 + Jab * A3t / rho2 ** 6)
    ds = 0.5 * (a ** 2) / n
    s = ds
    jj = 1
    while s != 0 and abs(ds / s) > EPS / 10:
        jj += 1
        ds = -(jj - 1) / (jj + 1) / n * (ds * a)
        s = s + ds
    p2 = np.exp(s) * np.sqrt((n + a) / n) * (n / rho) ** a
    C = p2 * (_stirling_f(n + a) / _stirling_f(n)) / np.sqrt(2)
    valstmp = C * vals
    denom = np.sin(t / 2) ** (a + 0.5) * np.cos(t / 2) ** (b + 0.5)
    vals = np.sqrt(t) * valstmp / denom
    C2 = C * n / (n + a) * (rho / rho2) ** a
    ders = (n * (a - b - (2 * n + a + b) * np.cos(t)) * valstmp + 2 * (n + a) * (n + b) * C2 * vals2) / (2 * n + a + b)
    ders = ders * (np.sqrt(t) / (denom * np.sin(t)))
    return vals, ders


def asy2(n, a, b, npts):
    """:363-398"""
    jk = approx_besselroots(a, npts)
    phik = jk / (n + 0.5 * (a + b + 1))
    t = phik + ((a ** 2 - 0.25) * (1 - phik * (1 / np.tan(phik))) / (2 * phik) - 0.25 * (a ** 2 - b ** 2) * np.tan(0.5 * phik)) / (n + 0.5 * (a + b + 1)) ** 2
    higher = asy2_higherterms(a, b, t)
    for _ in range(10):
        vals, ders = feval_asy2(n, a, b, t, higher)
        dt = vals / ders
        t = t + dt
        if np.max(np.abs(dt)) < np.sqrt(EPS) / 200:
            break
    vals, ders = feval_asy2(n, a, b, t, higher)
    dt = vals / ders
    t = t + dt
    t = t[::-1]
    ders = ders[::-1]
    return np.cos(t), 1.0 / ders ** 2


def jacobi_asy(n, a, b, info=None):
    """:170-194"""
    nbdy = 10
    x, w = asy1(n, a, b, nbdy, info)
    xb, wb = asy2(n, a, b, nbdy)
    x[n - nbdy:] = xb
    w[n - nbdy:] = wb
    if a != b:
        xb, wb = asy2(n, b, a, nbdy)
    x[:nbdy] = -xb[::-1]
    w[:nbdy] = wb[::-1]
    w = w * weights_constant(n, a, b)
    return x, w


# ---- n <= 100: gaussjacobi.jl:61-155 --------------------------------------------------------------

def innerjacobi_rec(n, x, a, b):
    """:129-155 — in C (oracle/jacobi_rec_oracle.c) so that muladd is a true fma."""
    from . import innerjacobi_rec as _c
    return _c(n, x, a, b)


def half_rec(n, a, b, flag):
    """:94-127"""
    m = int(np.ceil(n / 2)) if flag == 1 else int(np.floor(n / 2))
    r = np.arange(m, 0, -1, dtype=np.float64)
    c1 = 1 / (2 * n + a + b + 1)
    a1 = 0.25 - a ** 2
    b1 = 0.25 - b ** 2
    c12 = c1 ** 2
    C = (2.0 * r + (a - 0.5)) * (PI * c1)
    C_2 = C / 2
    x = np.cos(c12 * (-b1 * np.tan(C_2) + a1 * (1 / np.tan(C_2))) + C)
    for _ in range(10):
        P1, P2 = innerjacobi_rec(n, x, a, b)
        dx = P1 / P2
        dx2 = 0.0
        for v in dx * dx:            # ifelse(_dx2 > dx2, _dx2, dx2): NaN never wins
            if v > dx2:
                dx2 = v
        x = x - dx
        if not (dx2 > EPS / 1.0e6):
            break
    P1, P2 = innerjacobi_rec(n, x, a, b)
    return x, P2


def jacobi_rec(n, a, b):
    """:61-92"""
    from math import gamma
    x11, x12 = half_rec(n, a, b, 1)
    x21, x22 = half_rec(n, b, a, 0)
    x = np.zeros(n)
    w = np.zeros(n)
    m1, m2 = len(x11), len(x21)
    sum_w = 0.0
    for i in range(1, m2 + 1):
        idx = m2 + 1 - i
        xi = -x21[i - 1]
        wi = 1 / ((1 - xi ** 2) * x22[i - 1] ** 2)
        w[idx - 1] = wi
        x[idx - 1] = xi
        sum_w += wi
    for i in range(1, m1 + 1):
        idx = m2 + i
        xi = x11[i - 1]
        wi = 1 / ((1 - xi ** 2) * x12[i - 1] ** 2)
        w[idx - 1] = wi
        x[idx - 1] = xi
        sum_w += wi
    c = (2 ** (a + b + 1) * gamma(2 + a) * gamma(2 + b) / (gamma(2 + a + b) * (a + 1) * (b + 1)))
    return x, w * (c / sum_w)


def jacobi_jacobimatrix(n, a, b):
    """:570-584 -> (diag, offdiag)"""
    s_ = a + b
    ii = np.arange(2, n, dtype=np.float64)
    si = 2 * ii + s_
    aa = np.concatenate([[(b - a) / (2 + s_)], (b ** 2 - a ** 2) / ((si - 2) * si),
                         [(b ** 2 - a ** 2) / ((2 * n - 2 + s_) * (2 * n + s_))]])
    bb = np.concatenate([[2 * np.sqrt((1 + a) * (1 + b) / (s_ + 3)) / (s_ + 2)],
                         2 * np.sqrt(ii * (ii + a) * (ii + b) * (ii + s_) / (si ** 2 - 1)) / si])
    return aa, bb


def jacobimoment(a, b):
    """:586-591"""
    from math import exp, lgamma, log
    s_ = a + b
    return exp((s_ + 1) * log(2.0) + lgamma(a + 1) + lgamma(b + 1) - lgamma(2 + s_))


def jacobi_gw(n, a, b):
    """:593-599 (LAPACK symmetric-tridiagonal eigen-solve, as in the reference)."""
    from scipy.linalg import eigh_tridiagonal
    aa, bb = jacobi_jacobimatrix(n, a, b)
    x, V = eigh_tridiagonal(aa, bb)
    return x, V[0, :] ** 2 * jacobimoment(a, b)


def gaussradau_jacobi(n, a, b):
    """gaussradau.jl:52-69"""
    from scipy.linalg import eigh_tridiagonal
    if n <= 0:
        raise ValueError("DomainError: Input N must be a positive integer")
    m = n - 1
    s_ = a + b
    mu = jacobimoment(a, b)
    if n == 1:
        return np.array([-1.0]), np.array([mu])
    aa, bb = jacobi_jacobimatrix(n, a, b)
    aa[-1] = -1 + 2 * m * float(m + a) / ((2 * m + s_) * (2 * m + s_ + 1))
    x, V = eigh_tridiagonal(aa, bb)
    w = V[0, :] ** 2 * mu
    x[0] = -1
    return x, w


def gaussjacobi(n, a, b, info=None):
    """:23-55 for general (a, b): closed form n = 1, recurrence n <= 100, asymptotics n > 100 (max(a,b) < 5).
    (0,0) -> Legendre and (+-1/2, +-1/2) -> Chebyshev live in their own oracles."""
    from math import gamma
    a = float(a)
    b = float(b)
    if n < 0:
        raise ValueError("DomainError: n must be non-negative")
    if n == 0:
        return np.zeros(0), np.zeros(0)
    if n == 1:
        return np.array([(b - a) / (a + b + 2)]), np.array([2 ** (a + b + 1) * gamma(a + 1) * gamma(b + 1) / gamma(a + b + 2)])
    if min(a, b) <= -1.0:
        raise ValueError("DomainError: nonintegrable weight function")
    if n <= 100 and max(a, b) < 5.0:
        return jacobi_rec(n, a, b)
    if max(a, b) >= 5.0:
        if n > 4000:
            raise NotImplementedError("ErrorException: n must be <= 4000 for max(a,b) >= 5")
        return jacobi_gw(n, a, b)
    if n <= 100:
        return jacobi_rec(n, a, b)
    return jacobi_asy(n, a, b, info)


def gaussradau(n):
    """gaussradau.jl:31-49"""
    if n == 1:
        return np.array([-1.0]), np.array([2.0])
    if n == 2:
        return np.array([-1.0, 1 / 3]), np.array([0.5, 1.5])
    x, w = gaussjacobi(n - 1, 0.0, 1.0)
    w = w / (1 + x)
    return np.concatenate([[-1.0], x]), np.concatenate([[2 / n ** 2], w])


def gausslobatto(n):
    """gausslobatto.jl:31-52"""
    if n <= 1:
        raise ValueError("DomainError: Lobatto undefined for n <= 1.")
    if n == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    if n == 3:
        return np.array([-1.0, 0.0, 1.0]), np.array([1 / 3, 4 / 3, 1 / 3])
    x, w = gaussjacobi(n - 2, 1.0, 1.0)
    w = w / (1 - x ** 2)
    e = 2 / (n * (n - 1))
    return np.concatenate([[-1.0], x, [1.0]]), np.concatenate([[e], w, [e]])
