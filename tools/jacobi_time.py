"""Device time of gaussjacobi(n, 0.3, -0.7) (CUDA events, best of 5 after a warm-up) for a few sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
for n in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7):
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fgq_b200._lib.check(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"FGQ_JACOBI_PDL={os.environ.get('FGQ_JACOBI_PDL', '1')} n={n}: {best:.4f} ms  sum w err {float(w.sum()) - 4.554443087962173:.1e}")
