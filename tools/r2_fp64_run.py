"""One call of every bench configuration on cuda:0, device-resident entry points only, recording which of this
library's launches belong to which configuration (fgq_launch_count before / after).  Run under
  ncu --metrics <FP64 instruction counters>,gpu__time_duration.sum --csv --log-file gpurun_out/r2_fp64_ncu.csv
so that tools/r2_fp64_parse.py can sum the executed FP64 instructions per configuration
(profiles/r02_fp64_counts.json, read by bench.py for the per-config rooflines)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
chk = fgq_b200._lib.check
sp = 0
ranges = {}


def section(key, nodes, fn):
    torch.cuda.synchronize()
    a = fgq_b200.launch_count()
    fn()
    torch.cuda.synchronize()
    ranges[key] = {"first": a, "last": fgq_b200.launch_count(), "nodes": nodes}


def bufs(n):
    return torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")


n = 100_000
x, w = bufs(n)
section("C1_gausslegendre_1e5", n, lambda: chk(L.fgq_gausslegendre_device(n, 0, n, x.data_ptr(), w.data_ptr(), sp)))
n = 10 ** 9
sh = fgq_b200.gausslegendre_device(n)          # allocation + first fill (counted apart)
section("C2_gausslegendre_1e9", n, lambda: fgq_b200.gausslegendre_device(n, out=sh))
del sh
torch.cuda.empty_cache()
sums = torch.zeros(3, dtype=torch.float64, device="cuda")
m = (n + 1) // 2
section("fused_moments_gausslegendre_1e9", n,
        lambda: chk(L.fgq_gausslegendre_moments_fused_device(n, 1, m + 1, 0, sums.data_ptr(), sp)))
n = 10_000_000
x, w = bufs(n)
ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
section("C3_gaussjacobi_1e7", n,
        lambda: chk(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp)))
n = 1_000_000
x, w = bufs(n)
ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
section("C4_gausshermite_1e6", n,
        lambda: chk(L.fgq_gausshermite_device(n, 0, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp)))
nb = int(L.fgq_gausshermite_reduced_bound(n))
xr, wr = bufs(nb)
section("reduced_gausshermite_1e6", nb,
        lambda: chk(L.fgq_gausshermite_reduced_device(n, 0, xr.data_ptr(), wr.data_ptr(), ws.data_ptr(), sp, None, None)))
ws2 = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
section("C4_gausslaguerre_1e6", n,
        lambda: chk(L.fgq_gausslaguerre_device(n, 0.0, x.data_ptr(), w.data_ptr(), ws2.data_ptr(), sp)))
n_i = np.random.default_rng(0).integers(2, 61, size=1_000_000).astype(np.int32)
off = np.zeros(len(n_i) + 1, dtype=np.int64)
np.cumsum(n_i, out=off[1:])
dn, do = torch.from_numpy(n_i).cuda(), torch.from_numpy(off).cuda()
xb, wb = bufs(int(off[-1]))
section("C5_batch_1e6_rules", int(off[-1]),
        lambda: chk(L.fgq_gausslegendre_batch_device(len(n_i), dn.data_ptr(), do.data_ptr(), xb.data_ptr(), wb.data_ptr(), sp)))
nr = 100_000
rng = np.random.default_rng(0)


def batch(key, n_k, wsbytes, launch):
    o = np.zeros(len(n_k) + 1, dtype=np.int64)
    np.cumsum(n_k, out=o[1:])
    xb, wb = bufs(int(o[-1]))
    wsb = torch.empty(int(wsbytes), dtype=torch.uint8, device="cuda")
    section(key, int(o[-1]), lambda: chk(launch(o, xb, wb, wsb)))


n_j = rng.integers(2, 101, nr).astype(np.int32)
a_j, b_j = rng.uniform(-0.9, 4.9, nr), rng.uniform(-0.9, 4.9, nr)
batch("batch_gaussjacobi_1e5_rules", n_j, L.fgq_gaussjacobi_batch_workspace_bytes(nr),
      lambda o, xb, wb, wsb: L.fgq_gaussjacobi_batch_device(nr, n_j.ctypes.data, a_j.ctypes.data, b_j.ctypes.data, o.ctypes.data,
                                                            xb.data_ptr(), wb.data_ptr(), wsb.data_ptr(), sp))
n_h = rng.integers(2, 201, nr).astype(np.int32)
batch("batch_gausshermite_1e5_rules", n_h, L.fgq_gausshermite_batch_workspace_bytes(nr),
      lambda o, xb, wb, wsb: L.fgq_gausshermite_batch_device(nr, n_h.ctypes.data, o.ctypes.data, 0, xb.data_ptr(), wb.data_ptr(),
                                                             wsb.data_ptr(), sp))
n_l = rng.integers(2, 128, nr).astype(np.int32)
a_l = rng.uniform(-0.9, 4.9, nr)
batch("batch_gausslaguerre_1e5_rules", n_l, L.fgq_gausslaguerre_batch_workspace_bytes(nr),
      lambda o, xb, wb, wsb: L.fgq_gausslaguerre_batch_device(nr, n_l.ctypes.data, a_l.ctypes.data, o.ctypes.data, xb.data_ptr(),
                                                              wb.data_ptr(), wsb.data_ptr(), sp))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(ranges, open("gpurun_out/r2_fp64_ranges.json", "w"), indent=1)
print("ok", fgq_b200.launch_count())
