"""Host-returning gausslegendre(n): pinned vs pageable destination, first call vs steady state."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
fgq_b200.gausslegendre(1 << 22)
for label, mk in (("pageable (np.empty, untouched)", lambda: (np.empty(n), np.empty(n))),
                  ("pageable (pre-faulted)", lambda: (np.zeros(n) + 1, np.zeros(n) + 1)),
                  ("pinned", lambda: tuple(torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)))):
    out = mk()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        fgq_b200.gausslegendre(n, out=out)
        ts.append(time.perf_counter() - t0)
    print(f"{label:32s} first {ts[0]:.3f} s, then {min(ts[1:]):.3f} s  -> {n / min(ts[1:]) / 1e9:.2f} Gnodes/s; sum w - 2 = {out[1].sum() - 2:.1e}")
    del out
