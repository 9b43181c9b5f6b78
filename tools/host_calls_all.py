"""Wall time of the host-returning (drop-in) call of every family at the BASELINE.json sizes, fresh
pageable numpy outputs each call (what a Julia caller gets), best of 5 after a warm-up."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402


def wall(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3


n_i = np.random.default_rng(0).integers(2, 61, size=1_000_000).astype(np.int32)
for label, fn in (("gausslegendre(1e5)", lambda: fgq_b200.gausslegendre(100_000)),
                  ("gausslegendre(1e6)", lambda: fgq_b200.gausslegendre(1_000_000)),
                  ("gausslegendre(1e7)", lambda: fgq_b200.gausslegendre(10_000_000)),
                  ("gausslegendre(1e8)", lambda: fgq_b200.gausslegendre(100_000_000)),
                  ("gaussjacobi(1e7,.3,-.7)", lambda: fgq_b200.gaussjacobi(10_000_000, 0.3, -0.7)),
                  ("gaussjacobi(1e6,.3,-.7)", lambda: fgq_b200.gaussjacobi(1_000_000, 0.3, -0.7)),
                  ("gausshermite(1e6)", lambda: fgq_b200.gausshermite(1_000_000)),
                  ("gausslaguerre(1e6)", lambda: fgq_b200.gausslaguerre(1_000_000)),
                  ("gausschebyshevt(1e7)", lambda: fgq_b200.gausschebyshevt(10_000_000)),
                  ("batch 1e6 rules", lambda: fgq_b200.gausslegendre_batch(n_i))):
    best, med = wall(fn)
    print(f"{label:28s} best {best:9.3f} ms   median {med:9.3f} ms", flush=True)
