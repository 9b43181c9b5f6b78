"""Device time of gausslaguerre(n, alpha) (CUDA events, best of 6 after a warm-up), hand-over statistics and a digest
of the result.  FGQ_LAGUERRE_FAST=0 selects the sequential launch chain, FGQ_LAGUERRE_WINDOW=<count> forces the
fall-back of the large-rule path: the digests of the three must be equal."""
import ctypes
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
mode = f"FAST={os.environ.get('FGQ_LAGUERRE_FAST', '1')} WINDOW={os.environ.get('FGQ_LAGUERRE_WINDOW', '-')}"
for n, alpha in ((128, 0.0), (200, 0.0), (1000, -0.7), (1000, 4.5), (5000, 1.0), (8192, 0.0), (10 ** 4, 0.5), (10 ** 5, 0.0), (350000, 0.44), (10 ** 6, 0.0), (10 ** 6 + 1, 2.5), (10 ** 7, -0.5)):
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fgq_b200._lib.check(L.fgq_gausslaguerre_device(n, alpha, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    kb, ka, nf, nr = (ctypes.c_int64() for _ in range(4))
    fgq_b200._lib.check(L.fgq_gausslaguerre_stats(ws.data_ptr(), ctypes.byref(kb), ctypes.byref(ka), ctypes.byref(nf), ctypes.byref(nr)))
    h = hashlib.sha256(x.cpu().numpy().tobytes() + w.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"n={n} alpha={alpha}: k_b {kb.value} k_a {ka.value} flagged {nf.value} reduced {nr.value} warn {fgq_b200.lib().fgq_warnings()} digest {h}", flush=True)
    print(f"   [{mode}] {best * 1e3:.1f} us", file=sys.stderr, flush=True)
