"""gausslaguerre(n, alpha) on the device, three times (the command profiled by ncu for the Laguerre kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 6
x = torch.empty(n, dtype=torch.float64, device="cuda")
w = torch.empty(n, dtype=torch.float64, device="cuda")
ws = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
for _ in range(3):
    fgq_b200._lib.check(L.fgq_gausslaguerre_device(n, 0.0, x.data_ptr(), w.data_ptr(), ws.data_ptr(), 0))
torch.cuda.synchronize()
print("sum w", float(w.sum()))
