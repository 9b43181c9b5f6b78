"""Small invocation of every entry point (used under compute-sanitizer memcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

for n in (0, 1, 5, 42, 61, 1013, 26385, 100001):
    fgq_b200.gausslegendre(n)
fgq_b200.gausslegendre_batch(np.arange(0, 61))
for w in (1, 3):
    for r in range(w):
        fgq_b200.gausslegendre_device(30011, rank=r, world=w, mode="pair")
        fgq_b200.gausslegendre_device(30011, rank=r, world=w, mode="contiguous")
fgq_b200.gausslegendre_moments_fused(30011, with_exp=True)
for n, a, b in ((1, 1.0, 2.0), (42, -0.1, 0.3), (100, 4.5, -0.9), (251, -0.7, 1.3), (3000, 0.3, -0.7), (300, 0.5, -0.5)):
    fgq_b200.gaussjacobi(n, a, b)
for n in (1, 2, 3, 42, 500):
    fgq_b200.gaussradau(n)
    if n > 1:
        fgq_b200.gausslobatto(n)
for n in (1, 2, 18, 42, 200, 251, 5001):
    fgq_b200.gausshermite(n)
    fgq_b200.gausshermite(n, normalize=True)
for n, al in ((1, 0.4), (2, -0.37), (10, 0.5), (42, 0.0), (100, 6.0), (251, 0.0), (3000, 2.5)):
    fgq_b200.gausslaguerre(n, al)
    fgq_b200.gausslaguerre(n, al, reduced=True)
for k in (1, 2, 3, 4):
    fgq_b200._chebyshev(k, 1001) if hasattr(fgq_b200, "_chebyshev") else fgq_b200.api._chebyshev(k, 1001)
print("run_small ok, launches:", fgq_b200.launch_count())
