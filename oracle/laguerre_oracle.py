"""laguerre_oracle.py — CPU ORACLE (test infrastructure, NOT product code).

numpy restatement of the explicit-asymptotics Gauss-Laguerre path of FastGaussQuadrature.jl v1.3.0:
  gausslaguerre dispatch           src/gausslaguerre.jl:1,23,59-96
  gausslaguerre_asy                src/gausslaguerre.jl:110-261   (region hand-over, recompute hook)
  gl_bulk / gl_bessel / gl_airy    src/gausslaguerre.jl:273-382
  sum_adaptive                     src/gausslaguerre.jl:400-409   (T = -1: adaptive order)
  gl_bulk_solve_t                  src/gausslaguerre.jl:411-427
  region evaluators                src/gausslaguerre.jl:429-529
  compute_airy_root                src/gausslaguerre.jl:508-517
  gl_rec_newton, evalLaguerreRec   src/gausslaguerre.jl:547-579,625-641 (C: laguerre_rec_oracle.c)
  gausslaguerre_rec                src/gausslaguerre.jl:582-618   (15 <= n < 128, or alpha >= 5)
  gausslaguerre_GW                 src/gausslaguerre.jl:535-542   (n < 15; LAPACK through scipy)

The reference walks k = 1..n sequentially and hands over Bessel -> bulk -> Airy at the first k where
the next region's error estimate wins.  Here the three expansions are evaluated for whole index
ranges (numpy) and the same hand-over rule is applied to the arrays; the result is identical
because every node's value depends on (n, alpha, k) and on which region is selected, nothing else.

Arithmetic outside /root/reference: besselj (AMOS zbesj; scipy.special.jv is the same routine),
airyaiprime (AMOS zairy; scipy.special.airy on complex input).  evalpoly's muladd chain is plain
multiply-add here.  Pin: golden values of test/test_gausslaguerre.jl:66-144.  Julia cannot run in this
image: below the 1e-13 absolute level of those tests the parity is UNPINNED.
"""
from __future__ import annotations

import numpy as np
from scipy.special import airy, jv

from . import gl_rec_newton
from .hermite_oracle import AIRY_ROOTS
from .jacobi_oracle import approx_besselroots, mcmahon

EPS = np.finfo(np.float64).eps
PI = np.pi
UNDERFLOW = 10 * np.finfo(np.float64).tiny


def evalpoly(x, coeffs):
    r = coeffs[-1] + 0 * x
    for c in coeffs[-2::-1]:
        r = r * x + c
    return r


def gl_bulk(t, d, al):
    """:273-322 (general alpha) / :296-314 (alpha = 0)."""
    tinv = 1.0 / (1 - t)
    _t = (1 - t) / t
    _t2 = _t * _t
    _t3 = _t2 * _t
    d2 = d * d
    d4 = d2 * d2
    d6 = d4 * d2
    if al == 0:
        x3 = -d / 12 * evalpoly(tinv, (-4, -4, 5))
        x5 = d2 * d * _t / 720 * evalpoly(tinv, (224, -336, -16, -576, 2814, -3815, 1600))
        x7 = -d4 * d / 181440 * _t2 * evalpoly(tinv, (-285696, 714240, -516864, -13760, -17680, -1727136, 16131880,
                                                    -48469876, 66424575, -43122800, 10797500))
        x9 = d6 * d / 10886400 * _t3 * evalpoly(tinv, (210677760, -737372160, 916264704, -426423552, -10338048, 2863616,
                                                     -26285280, -3062050944, 43651480992, -212307298152, 518401904799,
                                                     -714465642135, 566519158800, -241928673000, 43222750000))
        w3 = d2 / 6 * (2 * t + 3) / (t - 1) ** 3
        w5 = d4 / 720 * _t2 * evalpoly(tinv, (0, 0, 112, 32, 1712, -12408, 27517, -24860, 8000))
        w7 = -d6 * _t3 / 90720 * evalpoly(tinv, (0, 0, -71424, 159744, 20640, 28480, 4300160, -50986344, 201908326,
                                               -386872990, 393326325, -204917300, 43190000))
        return (x3, x5, x7, x9), (w3, w5, w7)
    a2 = al * al
    c1 = evalpoly(a2, (-4, 12))
    x3 = -evalpoly(tinv, (c1, -4, 5)) * d / 12
    c1, c2 = evalpoly(a2, (224, -960, 480)), evalpoly(a2, (7, -30, 15))
    x5 = d2 * d * _t / 720 * evalpoly(tinv, (c1, -48 * c2, -16, -576, 2814, -3815, 1600))
    c1, c2, c3 = evalpoly(a2, (-285696, 1354752, -967680, 193536)), evalpoly(a2, (-31, 147, -105, 21)), evalpoly(a2, (-1346, 6405, -4620, 945))
    c4, c5 = evalpoly(a2, (43, -126, 63)), evalpoly(a2, (-221, -630, 315))
    x7 = -d4 * d / 181440 * _t2 * evalpoly(tinv, (c1, -23040 * c2, 384 * c3, -320 * c4, 80 * c5, -1727136, 16131880,
                                                -48469876, 66424575, -43122800, 10797500))
    c1, c2, c3 = (evalpoly(a2, (210677760, -1028505600, 812851200, -232243200, 24883200)),
                  evalpoly(a2, (127, -620, 490, -140, 15)), evalpoly(a2, (1193053, -5826660, 4613070, -1324260, 143325)))
    c4, c5, c6 = (evalpoly(a2, (555239, -2716980, 2163630, -631260, 70875)), evalpoly(a2, (-641, 2960, -2155, 450)),
                  evalpoly(a2, (-1598, 17685, -13905, 3375)))
    c7, c8, c9 = evalpoly(a2, (-7823, -9042, 4521)), evalpoly(a2, (15948182, -206850, 103425)), evalpoly(a2, (64957561, -24000, 12000))
    x9 = d6 * d / 10886400 * _t3 * evalpoly(tinv, (c1, -5806080 * c2, 768 * c3, -768 * c4, 16128 * c5, -1792 * c6, 3360 * c7,
                                                 -192 * c8, 672 * c9, -212307298152, 518401904799, -714465642135,
                                                 566519158800, -241928673000, 43222750000))
    w3 = d2 / 6 * (2 * t + 3) / (t - 1) ** 3
    c1 = evalpoly(a2, (7, -30, 15))
    w5 = d4 * _t2 / 720 * evalpoly(tinv, (0, 0, 16 * c1, 32, 1712, -12408, 27517, -24860, 8000))
    c1, c2, c3 = evalpoly(a2, (-31, 147, -105, 21)), evalpoly(a2, (-416, 1995, -1470, 315)), evalpoly(a2, (43, -126, 63))
    c4, c5 = evalpoly(a2, (-89, -378, 189)), evalpoly(a2, (53752, -630, 315))
    w7 = -d6 * _t3 / 90720 * evalpoly(tinv, (0, 0, 2304 * c1, -384 * c2, 480 * c3, -320 * c4, 80 * c5, -50986344,
                                           201908326, -386872990, 393326325, -204917300, 43190000))
    return (x3, x5, x7, x9), (w3, w5, w7)


def gl_bessel(jak, d, al):
    """:325-367"""
    j2 = jak * jak
    d2 = d * d
    d4 = d2 * d2
    d8 = d4 * d4
    if al == 0:
        x3 = d2 * evalpoly(j2, (-2, 1)) / 3
        x5 = d4 * evalpoly(j2, (94, -57, 11)) / 45
        x7 = d4 * d2 * evalpoly(j2, (-48308, 28102, -6516, 657)) / 2835
        x9 = d8 * evalpoly(j2, (12059918, -6605817, 1456807, -172740, 10644)) / 42525
        x11 = d8 * d2 * evalpoly(j2, (-11427291076, 6028914206, -1248722004, 138902061, -9908262, 410649)) / 1403325
        w3 = 2 * d2 * evalpoly(j2, (-1, 1)) / 3
        w5 = d4 * evalpoly(j2, (94, -114, 33)) / 45
        w7 = 4 * d4 * d2 * evalpoly(j2, (-12077, 14051, -4887, 657)) / 2835
        w9 = d8 * evalpoly(j2, (12059918, -13211634, 4370421, -690960, 53220)) / 42525
        w11 = 2 * d8 * d2 * evalpoly(j2, (-5713645538, 6028914206, -1873083006, 277804122, -24770655, 1231947)) / 1403325
        return (x3, x5, x7, x9, x11), (w3, w5, w7, w9, w11)
    a2 = al * al
    c1 = evalpoly(a2, (-2, 2))
    x3 = d2 * evalpoly(j2, (c1, 1)) / 3
    w3 = d2 * evalpoly(j2, (c1, 2)) / 3
    c1, c2 = evalpoly(a2, (94, -140, 46)), evalpoly(a2, (-19, 11))
    x5 = d4 * evalpoly(j2, (c1, 3 * c2, 11)) / 45
    w5 = d4 * evalpoly(j2, (c1, 6 * c2, 33)) / 45
    c1, c2, c3 = evalpoly(a2, (-12077, 19887, -9303, 1493)), evalpoly(a2, (14051, -10750, 2459)), evalpoly(a2, (-181, 73))
    x7 = d4 * d2 * evalpoly(j2, (4 * c1, 2 * c2, 36 * c3, 657)) / 2835
    w7 = 4 * d4 * d2 * evalpoly(j2, (c1, c2, 27 * c3, 657)) / 2835
    c1, c2 = evalpoly(a2, (6029959, -10087180, 5095482, -1146220, 107959)), evalpoly(a2, (-2201939, 1678761, -507801, 63299))
    c3, c4 = evalpoly(a2, (1456807, -729422, 125671)), evalpoly(a2, (-2879, 887))
    x9 = d8 * evalpoly(j2, (2 * c1, 3 * c2, c3, 60 * c4, 10644)) / 42525
    w9 = d8 * evalpoly(j2, (2 * c1, 6 * c2, 3 * c3, 240 * c4, 53220)) / 42525
    return (x3, x5, x7, x9), (w3, w5, w7, w9)


def gl_airy(ak, d, al):
    """:370-382"""
    cbrt_d = np.cbrt(d)
    cbrt_d2 = cbrt_d * cbrt_d
    ak2 = ak * ak
    ak3 = ak2 * ak
    ak4 = ak2 * ak2
    x1 = 1 / d + ak * np.cbrt(4.0) / cbrt_d
    x3 = ak2 * np.cbrt(16.0) / 5 * cbrt_d + (11.0 / 35 - al ** 2 - 12.0 / 175 * ak3) * d + (16.0 / 1575 * ak + 92.0 / 7875 * ak4) * np.cbrt(4.0) * d * cbrt_d2
    x5 = -(15152.0 / 3031875 * ak4 * ak + 1088.0 / 121275 * ak2) * np.cbrt(2.0) * d * d * cbrt_d
    return x1, x3, x5


def sum_adaptive(vals):
    """:400-409, vectorised: add terms while they decrease in magnitude."""
    z = vals[0].copy()
    alive = np.ones_like(z, dtype=bool)
    i_last = np.zeros(z.shape, dtype=np.int64)       # index i (0-based) of the last term added
    for i in range(1, len(vals)):
        alive = alive & (np.abs(vals[i]) < np.abs(vals[i - 1]))
        z = np.where(alive, z + vals[i], z)
        i_last = np.where(alive, i, i_last)
    nxt = np.minimum(i_last + 1, len(vals) - 1)
    stack = np.stack([np.abs(v) for v in vals])
    delta = np.take_along_axis(stack, nxt[None, :], axis=0)[0]
    return z, delta


def gl_bulk_solve_t(n, k, d):
    """:411-427, vectorised over k with each node's own stopping rule."""
    pt = (4 * n - 4 * k + 3) * d
    t = PI ** 2 / 16 * (pt - 1) ** 2
    diff = np.full(t.shape, 100.0)
    it = 0
    active = np.ones(t.shape, dtype=bool)
    iters = np.zeros(t.shape, dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        while it < 20:
            active = active & (np.abs(diff) > 100 * EPS)
            if not active.any():
                break
            it += 1
            newdiff = (pt * PI + 2 * np.sqrt(t - t ** 2) - np.arccos(2 * t - 1)) * np.sqrt(t / (1 - t)) / 2
            diff = np.where(active, newdiff, diff)
            t = np.where(active, t - newdiff, t)
            iters = np.where(active, iters + 1, iters)
    return t, iters


def asy_bulk(n, al, k, d):
    """:429-465 with T = -1"""
    t, iters = gl_bulk_solve_t(n, k, d)
    with np.errstate(all="ignore"):
        xs, ws = gl_bulk(t, d, al)
        xk, xdelta = sum_adaptive(xs)
        wk, wdelta = sum_adaptive(ws)
        xk = xk + t / d
        if al == 0:
            wfactor = np.exp(-xk) * 2 * PI * np.sqrt(t / (1 - t))
        else:
            wfactor = xk ** al * np.exp(-xk) * 2 * PI * np.sqrt(t / (1 - t))
        wk = wfactor * (1 + wk)
        wdelta = wdelta * wfactor
    return xk, wk, np.maximum(xdelta, wdelta), iters


def asy_bessel(n, al, jak, d):
    """:468-506 with T = -1"""
    with np.errstate(all="ignore"):
        xs, ws = gl_bessel(jak, d, al)
        xk, xdelta = sum_adaptive(xs)
        wk, wdelta = sum_adaptive(ws)
        xfactor = jak ** 2 * d
        xk = xfactor * (1 + xk)
        xdelta = xdelta * xfactor
        if al == 0:
            wfactor = 4 * d * np.exp(-xk) / jv(-1, jak) ** 2
        else:
            wfactor = 4 * d * xk ** al * np.exp(-xk) / jv(al - 1, jak) ** 2
        wk = wfactor * (1 + wk)
        wdelta = wdelta * wfactor
    return xk, wk, np.maximum(xdelta, wdelta)


def compute_airy_root(n, k):
    """:508-517"""
    index = n - k + 1
    t = 3 * PI / 2 * (index - 0.25)
    ak = -t ** (2 / 3) * evalpoly(1.0 / (t * t), (6967296, 725760, -967680, 6478500, -10856875)) / 6967296
    small = index <= 11
    ak = np.where(small, AIRY_ROOTS[np.minimum(index, 11) - 1], ak)
    return ak


def asy_airy(n, al, k, d):
    """:519-529 with T = -1"""
    ak = compute_airy_root(n, k)
    with np.errstate(all="ignore"):
        xs = gl_airy(ak, d, al)
        xk, xdelta = sum_adaptive([np.asarray(v, dtype=np.float64) for v in xs])
        aip = airy(ak.astype(np.complex128))[1].real
        wk = 4 ** (1 / 3) * xk ** (al + 1 / 3) * np.exp(-xk) / aip ** 2
        wdelta = np.abs(wk)
    return xk, wk, np.maximum(xdelta, wdelta)


def gausslaguerre_asy(n, al, reduced=False, recompute=True, info=None):
    """:110-261 with T = -1 (the call made by the dispatch at :91)."""
    d = 1.0 / (4 * n + 2 * al + 2)
    k_bessel = max(int(np.ceil(np.sqrt(n))), 7)
    k_airy = int(np.floor(0.9 * n))
    jak_vector = approx_besselroots(al, k_bessel)
    kk = np.arange(1, n + 1, dtype=np.float64)
    ki = np.arange(1, n + 1)

    # Bessel vs bulk, k = 1.. until bulk wins (:138-172).  Evaluate on a growing prefix.
    K = min(n, max(4 * k_bessel, 64))
    while True:
        jak = np.array([jak_vector[k - 1] if k < k_bessel else mcmahon(al, k) for k in range(1, K + 1)])
        xb, wb, db = asy_bessel(n, al, jak, d)
        xu, wu, du, _ = asy_bulk(n, al, kk[:K], d)
        wins = du < db
        if wins.any() or K == n:
            break
        K = min(n, 4 * K)
    k_b = int(np.argmax(wins)) + 1 if wins.any() else n      # first k where bulk wins (that k uses bulk)
    x = np.zeros(n)
    w = np.zeros(n)
    delta = np.zeros(n)
    x[:k_b - 1] = xb[:k_b - 1]; w[:k_b - 1] = wb[:k_b - 1]
    delta[:k_b - 1] = np.minimum(db[:k_b - 1], du[:k_b - 1])
    if wins.any():
        x[k_b - 1] = xu[k_b - 1]; w[k_b - 1] = wu[k_b - 1]; delta[k_b - 1] = min(db[k_b - 1], du[k_b - 1])
    else:
        x[k_b - 1] = xb[k_b - 1]; w[k_b - 1] = wb[k_b - 1]; delta[k_b - 1] = min(db[k_b - 1], du[k_b - 1])
    k = k_b
    k_a = None
    if k < n:
        # bulk only up to k_airy - 1 (:176-197)
        hi = max(k, k_airy - 1)
        if hi > k:
            sl = slice(k, hi)
            xs_, ws_, ds_, _ = asy_bulk(n, al, kk[sl], d)
            x[sl] = xs_; w[sl] = ws_; delta[sl] = ds_
            k = hi
        # bulk vs Airy until Airy wins (:201-229), then Airy only (:232-253)
        sl = slice(k, n)
        xu2, wu2, du2, _ = asy_bulk(n, al, kk[sl], d)
        xa, wa, da = asy_airy(n, al, ki[sl], d)
        awins = da < du2
        first = int(np.argmax(awins)) if awins.any() else (n - k)
        # positions < first: bulk; position == first: Airy (the winner); after: Airy only
        x[k:k + first] = xu2[:first]; w[k:k + first] = wu2[:first]
        delta[k:k + first] = np.minimum(da[:first], du2[:first])
        if awins.any():
            p = k + first
            x[p] = xa[first]; w[p] = wa[first]; delta[p] = min(da[first], du2[first])
            x[p + 1:] = xa[first + 1:]; w[p + 1:] = wa[first + 1:]; delta[p + 1:] = da[first + 1:]
            k_a = p + 1
    flagged = []
    if recompute:
        for i in np.nonzero(delta > 1.0e-13)[0]:
            xr, wr = gl_rec_newton(x[i], n, al)
            flagged.append(int(i) + 1)
            if abs(xr - x[i]) < 100 * delta[i]:
                x[i] = xr
                w[i] = wr
    if info is not None:
        info.update(k_b=k_b, k_a=k_a, flagged=flagged)
    if reduced:
        small = np.nonzero(np.abs(w) < UNDERFLOW)[0]
        if len(small):
            return x[:small[0]], w[:small[0]]
    return x, w


def gausslaguerre_rec(n, al, reduced=False):
    """:582-618"""
    x = np.zeros(n)
    w = np.zeros(n)
    n_pre = min(n, 7)
    nu = 4 * n + 2 * al + 2
    x_pre = approx_besselroots(al, n_pre) ** 2 / nu
    no_underflow = True
    for k in range(1, n + 1):
        if k <= n_pre:
            xk = x_pre[k - 1]
        else:
            xk = (7 * x[k - 2] - 21 * x[k - 3] + 35 * x[k - 4] - 35 * x[k - 5] + 21 * x[k - 6] - 7 * x[k - 7] + x[k - 8])
        xk, wk = gl_rec_newton(xk, n, al, computeweight=no_underflow)
        if no_underflow and abs(wk) < UNDERFLOW:
            no_underflow = False
        if reduced and not no_underflow:
            return x[:k - 1], w[:k - 1]
        x[k - 1] = xk
        w[k - 1] = wk
    return x, w


def gausslaguerre_gw(n, al):
    """:535-542 (LAPACK symmetric-tridiagonal eigen-solve, as in the reference)."""
    from math import gamma
    from scipy.linalg import eigh_tridiagonal
    k = np.arange(1, n + 1, dtype=np.float64)
    diag = 2 * k + (al - 1)
    off = np.sqrt(k[:-1] * (al + k[:-1]))
    x, V = eigh_tridiagonal(diag, off)
    return x, gamma(al + 1) * V[0, :] ** 2


def gausslaguerre(n, al=0.0, reduced=False, info=None):
    """:59-96"""
    from math import gamma, sqrt
    al = float(al)
    if al <= -1:
        raise ValueError("DomainError: alpha <= -1")
    if n < 0:
        raise ValueError("DomainError: n must be positive")
    if n == 0:
        return np.zeros(0), np.zeros(0)
    if n == 1:
        return np.array([1 + al]), np.array([gamma(1 + al)])
    if n == 2:
        s = sqrt(al + 2)
        return (np.array([al + 2 - s, al + 2 + s]),
                np.array([((al - s + 2) * gamma(al + 2)) / (2 * (al + 2) * (s - 1) ** 2),
                          ((al + s + 2) * gamma(al + 2)) / (2 * (al + 2) * (s + 1) ** 2)]))
    if n < 15:
        return gausslaguerre_gw(n, al)
    if n < 128 or al >= 5:
        return gausslaguerre_rec(n, al, reduced=reduced)
    return gausslaguerre_asy(n, al, reduced=reduced, recompute=True, info=info)
