"""Device-resident gausslegendre(n) timing sweep (CUDA events, best of 5): Gnodes/s and GB/s written."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

for n in [10 ** 5, 10 ** 6, 850000, 850001, 10 ** 7, 3 * 10 ** 7, (1 << 25) - 2, 1 << 25, 10 ** 8, 10 ** 9]:
    sh = fgq_b200.gausslegendre_device(n)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fgq_b200.gausslegendre_device(n, out=sh)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"n={n:>11d}  {best * 1e3:9.1f} us  {n / best / 1e6:8.1f} Gnodes/s  {16 * n / best / 1e6:8.1f} GB/s")
    del sh
