"""Times the headline kernel (n = 1e9 rule, device-resident, one GPU) of ANY build of the library: the .so named by
FGQ_LIB_PATH (default: the package's) is loaded directly, so older / experimental builds with a different symbol set
can be compared on the same box.  Reports the first and last ten steps of a sustained loop (power / clock effects)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.environ.get("FGQ_LIB_PATH") or os.path.join(ROOT, "fastgaussquadrature.jl_b200", "libfgq_cuda.so")
L = ctypes.CDLL(path)
L.fgq_gausslegendre_device_pair.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 5
L.fgq_gausslegendre_device_pair.restype = ctypes.c_int
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = (n + 1) // 2
pad = m + (m & 1)
xb = torch.empty(2 * pad, dtype=torch.float64, device="cuda")
wb = torch.empty(2 * pad, dtype=torch.float64, device="cuda")
sp = torch.cuda.current_stream().cuda_stream


def step():
    rc = L.fgq_gausslegendre_device_pair(n, 1, m + 1, xb.data_ptr(), wb.data_ptr(), xb.data_ptr() + 8 * pad,
                                         wb.data_ptr() + 8 * pad, sp)
    assert rc == 0, rc


for _ in range(5):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    step()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
srt = sorted(ms)
k = min(10, steps)
print(f"{os.path.basename(path):28s} n={n} steps={steps} mean {sum(ms) / steps:.4f} ms  min {srt[0]:.4f}  median {srt[steps // 2]:.4f}  "
      f"first{k} {sum(ms[:k]) / k:.4f}  last{k} {sum(ms[-k:]) / k:.4f}  -> {n / (sum(ms) / steps) / 1e6:.1f} Gnodes/s  "
      f"sum w = {float(wb[:m].sum() + wb[pad:pad + n // 2].sum()):.15f}", flush=True)
