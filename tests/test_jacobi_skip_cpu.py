"""The Jacobi truncation-norm passes skip the interior nodes whose angle lies beyond `tskip` (DESIGN.md 8.1,
`jacobi_norm_skip_angle`): the claim is that for theta >= tskip EVERY candidate term m = 12..20 of the interior
series (gaussjacobi.jl:296-317) is below the reference's threshold eps/100 whatever its phase, so such a node cannot
change the outcome of the comparison the norms feed.  Checked here, without a device, against the oracle's own
evaluation of those terms."""
import ctypes

import numpy as np
import pytest

from oracle import jacobi_oracle as jo

EPS = np.finfo(float).eps
PI = np.pi


def _terms(n, a, b, t):
    """|dS1 + dS2| of terms m = 12..20 at the angles t, as feval_asy1 forms them (oracle/jacobi_oracle.py)."""
    M = 20
    PHI, _ = jo._phi_tables(n, a, b, M)
    csc = (1.0 / np.sin(t / 2)) / 2
    sec = (1.0 / np.cos(t / 2)) / 2
    out = {}
    for m in range(12, M + 1):
        A = ((2 * n + a + b) + m) * t / 2 - (a + 0.5) * PI / 2
        pw = np.stack([csc ** l * sec ** (m - 1 - l) for l in range(m)], axis=1)
        lo, le = np.arange(0, m, 2), np.arange(1, m, 2)
        d1 = (pw[:, lo] @ PHI[lo, m - 1]) * np.cos(A)
        d2 = (pw[:, le] @ PHI[le, m - 1]) * np.sin(A)
        out[m] = np.abs(d1 + d2)
    return out


@pytest.mark.parametrize("n,a,b", [(101, 0.3, -0.7), (300, -0.9, 2.5), (1000, 4.9, 4.9), (12345, 0.0, 1.0),
                                    (10 ** 6, 0.3, -0.7), (10 ** 7, 0.3, -0.7), (5000, -0.99, -0.99)])
def test_skipped_nodes_are_below_the_truncation_threshold(n, a, b):
    import fgq_b200
    L = fgq_b200.lib()
    tr, tl = ctypes.c_double(0), ctypes.c_double(0)
    assert L.fgq_debug_jacobi_skip_angles(n, a, b, ctypes.byref(tr), ctypes.byref(tl)) == 0
    for (aa, bb, tskip) in ((a, b, tr.value), (b, a, tl.value)):
        if not np.isfinite(tskip):
            continue   # nothing is skipped for this half
        assert 0 < tskip <= PI / 2
        # angles from the skip angle up to the end of the half (pi/2 + 1e-3, the limit the kernel applies),
        # dense near tskip where the terms are largest
        t = np.unique(np.concatenate([tskip * (1 + np.logspace(-12, 0, 400)), np.linspace(tskip, PI / 2 + 1e-3, 2000)]))
        t = t[t <= PI / 2 + 1e-3]
        terms = _terms(n, aa, bb, t)
        worst = max(float(v.max()) for v in terms.values())
        assert worst < EPS / 100, (n, aa, bb, tskip, worst)
        # and the bound is not vacuous: it leaves a margin of at most ~4-40x at the skip angle itself
        assert worst < EPS / 100 / 3.5
    # the skip angle scales like 1/n: n * tskip is a constant of the series (~77), i.e. ~25 nodes per end stay
    if n >= 10 ** 6 and np.isfinite(tr.value):
        assert 40 < n * tr.value < 200


@pytest.mark.parametrize("n,a,b", [(101, 0.3, -0.7), (9000, 0.3, -0.7), (20000, -0.9, 2.5), (123457, 4.9, 4.9),
                                    (10 ** 6, 0.3, -0.7), (3 * 10 ** 6 + 1, 0.0, 1.0), (10 ** 7, 0.3, -0.7),
                                    (10 ** 6, -0.99, -0.99)])
def test_norm_pass_tiles_hold_every_node_below_the_skip_angle(n, a, b):
    """The norm passes do not even read the tiles the host rules out (`jacobi_norm_tiles`): every node whose initial
    theta (gaussjacobi.jl:202-204, the formula of jacobi_init_kernel) is below the skip angle — with a margin of 50 %
    for the Newton updates, which move an interior theta by a 1e-6 fraction at most — must lie in a tile that is
    walked; for large rules the walked range is a few dozen tiles out of tens of thousands."""
    import fgq_b200
    L = fgq_b200.lib()
    tr, tl = ctypes.c_double(0), ctypes.c_double(0)
    assert L.fgq_debug_jacobi_skip_angles(n, a, b, ctypes.byref(tr), ctypes.byref(tl)) == 0
    out = (ctypes.c_int64 * 5)()
    assert L.fgq_debug_jacobi_norm_tiles(n, a, b, out) == 0
    n_left, lo_r, hi_r, lo_l, hi_l = (int(v) for v in out)
    i = np.arange(n, dtype=np.float64)
    j = n - i
    d = 2 * n + a + b + 1
    K = PI * (2 * j + a - 0.5) / d
    tt = K + (1 / (d * d)) * ((0.25 - a * a) * (1.0 / np.tan(K / 2)) - (0.25 - b * b) * np.tan(K / 2))
    assert 0 < n_left < n and tt[n_left - 1] > PI / 2 >= tt[n_left]
    T = 128
    # right half: nodes n_left .. n-1, theta = tt; r = index - n_left
    th = tt[n_left:]
    need = np.nonzero(th < 1.5 * tr.value)[0] if np.isfinite(tr.value) else np.arange(len(th))
    if len(need):
        assert lo_r * T <= need.min() and need.max() < hi_r * T
    assert 0 <= lo_r < hi_r == (len(th) + T - 1) // T
    # left half: nodes 0 .. n_left-1, theta = pi - tt; r = index
    th = PI - tt[:n_left]
    need = np.nonzero(th < 1.5 * tl.value)[0] if np.isfinite(tl.value) else np.arange(len(th))
    if len(need):
        assert lo_l * T <= need.min() and need.max() < hi_l * T
    assert 0 == lo_l < hi_l <= (len(th) + T - 1) // T
    if n >= 10 ** 6:
        assert hi_r - lo_r <= 64 and hi_l - lo_l <= 64   # the bound is not vacuous
    if n <= 9000:
        assert lo_r == 0 and hi_l == (n_left + T - 1) // T   # small halves are walked whole
