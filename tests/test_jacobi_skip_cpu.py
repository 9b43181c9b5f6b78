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
