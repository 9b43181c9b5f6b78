"""One batch of 1e5 small rules per family (Jacobi / Hermite / Laguerre) and one truncated gausshermite(1e6)
on cuda:0 through the device entry points.  Used as the command under ncu for profiles/r01_small_batch_*."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
chk = fgq_b200._lib.check
sp = 0
NR = 100_000
rng = np.random.default_rng(0)


def layout(n_i):
    off = np.zeros(len(n_i) + 1, dtype=np.int64)
    np.cumsum(n_i, out=off[1:])
    tot = int(off[-1])
    return off, torch.empty(tot, dtype=torch.float64, device="cuda"), torch.empty(tot, dtype=torch.float64, device="cuda")


n_j = rng.integers(2, 101, NR).astype(np.int32)
a_j, b_j = rng.uniform(-0.9, 4.9, NR), rng.uniform(-0.9, 4.9, NR)
off, x, w = layout(n_j)
ws = torch.empty(int(L.fgq_gaussjacobi_batch_workspace_bytes(NR)), dtype=torch.uint8, device="cuda")
chk(L.fgq_gaussjacobi_batch_device(NR, n_j.ctypes.data, a_j.ctypes.data, b_j.ctypes.data, off.ctypes.data,
                                   x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
n_h = rng.integers(2, 201, NR).astype(np.int32)
off, x, w = layout(n_h)
ws = torch.empty(int(L.fgq_gausshermite_batch_workspace_bytes(NR)), dtype=torch.uint8, device="cuda")
chk(L.fgq_gausshermite_batch_device(NR, n_h.ctypes.data, off.ctypes.data, 0, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
n_l = rng.integers(2, 128, NR).astype(np.int32)
a_l = rng.uniform(-0.9, 4.9, NR)
off, x, w = layout(n_l)
ws = torch.empty(int(L.fgq_gausslaguerre_batch_workspace_bytes(NR)), dtype=torch.uint8, device="cuda")
chk(L.fgq_gausslaguerre_batch_device(NR, n_l.ctypes.data, a_l.ctypes.data, off.ctypes.data, x.data_ptr(), w.data_ptr(),
                                     ws.data_ptr(), sp))
n = 1_000_000
nb = int(L.fgq_gausshermite_reduced_bound(n))
x = torch.empty(nb, dtype=torch.float64, device="cuda")
w = torch.empty(nb, dtype=torch.float64, device="cuda")
ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
chk(L.fgq_gausshermite_reduced_device(n, 0, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp, None, None))
torch.cuda.synchronize()
print("ok")
