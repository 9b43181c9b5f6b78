"""SHA-256 digests of the small-rule branches (Hermite n <= 200, Jacobi n <= 100, Laguerre n < 128) over a
fixed parameter set: a bit-level regression check for optimisations that must not change results."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

if os.environ.get("FGQ_DIGEST_LIB"):  # compare against another build of the library
    sys.modules[fgq_b200.lib.__module__].LIB_PATH = os.environ["FGQ_DIGEST_LIB"]

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
# the asymptotic Jacobi path (+ Radau / Lobatto epilogues), Hermite and Laguerre asymptotics
for n, a, b in ((101, 0.3, -0.7), (1000, -0.9, 2.5), (12345, 4.9, 4.9), (1_000_000, 0.3, -0.7), (200_001, 1.0, 0.0)):
    print(f"jacobi asy {n} {a} {b}".ljust(34), digest(fgq_b200.gaussjacobi(n, a, b)))
print("radau 5000".ljust(34), digest(fgq_b200.gaussradau(5000)))
print("lobatto 5000".ljust(34), digest(fgq_b200.gausslobatto(5000)))
print("hermite 100000".ljust(34), digest(fgq_b200.gausshermite(100_000)))
print("laguerre 100000".ljust(34), digest(fgq_b200.gausslaguerre(100_000, 0.5)))
print("legendre 1000001".ljust(34), digest(fgq_b200.gausslegendre(1_000_001)))
for n in (201, 5001, 1_000_000):
    print(f"hermite asy {n}".ljust(34), digest(fgq_b200.gausshermite(n)))
print("hermite asy 100001 normalize".ljust(34), digest(fgq_b200.gausshermite(100_001, normalize=True)))
print("hermite reduced 1000000".ljust(34), digest(fgq_b200.gausshermite_reduced(1_000_000)[1:]))
