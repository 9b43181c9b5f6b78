"""Runs each BASELINE.json configuration other than the headline once or twice on cuda:0 (device-
resident entry points).  Used as the command under ncu for profiles/r01_configs_*."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
chk = fgq_b200._lib.check
sp = 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    n = 100_000
    x = torch.empty(n, dtype=torch.float64, device="cuda"); w = torch.empty_like(x)
    chk(L.fgq_gausslegendre_device(n, 0, n, x.data_ptr(), w.data_ptr(), sp))
    n = 10_000_000
    x = torch.empty(n, dtype=torch.float64, device="cuda"); w = torch.empty_like(x)
    ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    chk(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
    n = 1_000_000
    x = torch.empty(n, dtype=torch.float64, device="cuda"); w = torch.empty_like(x)
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    chk(L.fgq_gausshermite_device(n, 0, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
    ws = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    chk(L.fgq_gausslaguerre_device(n, 0.0, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
    n_i = np.random.default_rng(0).integers(2, 61, size=1_000_000).astype(np.int32)
    off = np.zeros(len(n_i) + 1, dtype=np.int64); np.cumsum(n_i, out=off[1:])
    dn = torch.from_numpy(n_i).cuda(); do = torch.from_numpy(off).cuda()
    x = torch.empty(int(off[-1]), dtype=torch.float64, device="cuda"); w = torch.empty_like(x)
    chk(L.fgq_gausslegendre_batch_device(len(n_i), dn.data_ptr(), do.data_ptr(), x.data_ptr(), w.data_ptr(), sp))
    torch.cuda.synchronize()
print("ok")
