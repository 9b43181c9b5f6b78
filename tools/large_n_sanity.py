"""One-off sanity run of the non-Legendre families at n = 1e8 (no oracle at this size: invariants only)."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


(x, w, ws), ms = timed(lambda: fgq_b200.gaussjacobi_device(n, 0.3, -0.7))
print(f"jacobi n={n}: {ms:.1f} ms, sum w err {float(w.sum()) - 4.554443087962173:.2e}, sorted {bool((x[1:] > x[:-1]).all())}, "
      f"first moment err {float((w * x).sum()) - (-1.0 / 1.6) * 4.554443087962173:.2e}")
del x, w, ws
(x, w, ws), ms = timed(lambda: fgq_b200.gausshermite_device(n))
print(f"hermite n={n}: {ms:.1f} ms, sum w err {float(w.sum()) - math.sqrt(math.pi):.2e}, sorted {bool((x[1:] > x[:-1]).all())}, "
      f"second moment err {float((w * x * x).sum()) - math.sqrt(math.pi) / 2:.2e}, nonzero weights {int((w > 0).sum())}")
del x, w, ws
(x, w, ws), ms = timed(lambda: fgq_b200.gausslaguerre_device(n, 0.0))
print(f"laguerre n={n}: {ms:.1f} ms, sum w err {float(w.sum()) - 1:.2e}, sorted {bool((x[1:] > x[:-1]).all())}, "
      f"first moment err {float((w * x).sum()) - 1:.2e}, nonzero weights {int((w > 0).sum())}")
