"""Device time of gausshermite(n) (CUDA events, best of 7 after a warm-up) and a digest of the result."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
for n, normalize in ((201, 0), (5001, 0), (10 ** 5, 0), (10 ** 6, 0), (10 ** 6 + 1, 1), (10 ** 7, 0)):
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fgq_b200._lib.check(L.fgq_gausshermite_device(n, normalize, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h = hashlib.sha256(x.cpu().numpy().tobytes() + w.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"gausshermite n={n} normalize={normalize}: {best * 1e3:.1f} us  digest {h}", flush=True)
