"""SHA-256 digests of the small-rule branches (Hermite n <= 200, Jacobi n <= 100, Laguerre n < 128) over a
fixed parameter set: a bit-level regression check for optimisations that must not change results."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

rng = np.random.default_rng(5)


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


n_h = np.arange(0, 201, dtype=np.int32)
print("hermite      ", digest(fgq_b200.gausshermite_batch(n_h)[1:]))
print("hermite norm ", digest(fgq_b200.gausshermite_batch(n_h, normalize=True)[1:]))
n_j = np.tile(np.arange(0, 101, dtype=np.int32), 4)
a = rng.uniform(-0.95, 4.9, len(n_j))
b = rng.uniform(-0.95, 4.9, len(n_j))
a[:101] = b[:101]
print("jacobi       ", digest(fgq_b200.gaussjacobi_batch(n_j, a, b)[1:]))
n_l = np.tile(np.arange(0, 128, dtype=np.int32), 3)
al = rng.uniform(-0.95, 6.0, len(n_l))
al[:128] = 0.0
print("laguerre     ", digest(fgq_b200.gausslaguerre_batch(n_l, al)[1:]))
x, w = fgq_b200.gausshermite(137)
print("hermite single 137 ", digest([x, w]))
x, w = fgq_b200.gaussjacobi(77, 0.3, -0.7)
print("jacobi single 77   ", digest([x, w]))
_, xb, wb = fgq_b200.gausslegendre_batch(np.arange(0, 61))
print("legendre batch     ", digest([xb, wb]))
