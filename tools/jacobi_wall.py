"""Wall-clock (host + device) of the device-resident gaussjacobi call, to expose host-side planning cost."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200
L = fgq_b200.lib(); chk = fgq_b200._lib.check
for n in (300, 1013, 10013, 100000, 1000000, 10000000):
    x = torch.empty(n, dtype=torch.float64, device="cuda"); w = torch.empty_like(x)
    ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    call = lambda: chk(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), 0))
    call(); torch.cuda.synchronize()
    best = 1e9; best_enq = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); call(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        best = min(best, t2 - t0); best_enq = min(best_enq, t1 - t0)
    print(f"n={n:>9d}: wall {best*1e3:8.2f} ms (host enqueue+planning {best_enq*1e3:8.2f} ms)")
