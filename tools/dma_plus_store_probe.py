"""Does a D2H DMA stream take bandwidth away from the host threads' non-temporal stores?  (Whether letting the copy
engine write part of a host-returning result could lift the e2e rate above the host threads' own store bandwidth.)
Measures: the host store peak alone (8 / 12 / 16 threads), the DMA rate alone, and both at once."""
import ctypes
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
n = 1 << 29  # 4 GiB of doubles
a = fgq_b200.empty_pinned(n)
a[:] = 0.0
b = torch.empty(n, dtype=torch.float64, pin_memory=True)
d = torch.zeros(n, dtype=torch.float64, device="cuda")
g = ctypes.c_double()


def store(threads):
    best = 0.0
    for _ in range(3):
        L.fgq_measure_host_store_peak(a.ctypes.data, n, threads, ctypes.byref(g))
        best = max(best, g.value)
    return best


for t in (8, 12, 16):
    print(f"host stores alone, {t} threads: {store(t):.1f} GB/s", flush=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    b.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print(f"DMA alone: {4 * n * 8 / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
stop = False
copied = [0, 0.0]


def dma_loop():
    t0 = time.perf_counter()
    while not stop:
        b.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        copied[0] += 1
    copied[1] = time.perf_counter() - t0


for t in (12, 16):
    stop = False
    copied = [0, 0.0]
    th = threading.Thread(target=dma_loop)
    th.start()
    time.sleep(0.2)
    s = store(t)
    stop = True
    th.join()
    print(f"both at once, {t} store threads: host stores {s:.1f} GB/s + DMA {copied[0] * n * 8 / copied[1] / 1e9:.1f} GB/s", flush=True)
