"""Times the headline kernel (n = 1e9 rule, device-resident): 20 tail launches alone, then a sustained loop."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sh = fgq_b200.gausslegendre_device(n)
torch.cuda.synchronize()
for _ in range(5):
    fgq_b200.gausslegendre_device(n, out=sh)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    fgq_b200.gausslegendre_device(n, out=sh)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
ms_sorted = sorted(ms)
print(f"{os.environ.get('FGQ_LIB_PATH', 'default')}: n={n} steps={steps} mean {sum(ms) / steps:.4f} ms  min {ms_sorted[0]:.4f}  "
      f"median {ms_sorted[steps // 2]:.4f}  first10 {sum(ms[:10]) / 10:.4f}  last10 {sum(ms[-10:]) / 10:.4f}  "
      f"-> {n / (sum(ms) / steps) / 1e6:.1f} Gnodes/s")
