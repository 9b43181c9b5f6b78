"""Timing of the batched small rules of the Jacobi / Hermite / Laguerre families:
device time of the batch kernel (CUDA events on the launch stream, host planning + upload excluded and
reported separately) and the wall time of the whole host-returning call."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402

L = fgq_b200.lib()
NR = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rng = np.random.default_rng(0)


def run(label, n_i, wsbytes, launch, host_call):
    offsets = np.zeros(len(n_i) + 1, dtype=np.int64)
    np.cumsum(n_i, out=offsets[1:])
    total = int(offsets[-1])
    out = torch.empty((2, total), dtype=torch.float64, device="cuda")
    ws = torch.empty(wsbytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    best_dev, best_call = 1e30, 1e30
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        # the upload of the rule descriptors is enqueued before e0's successor kernel: time the span
        # from the first enqueued operation to the end of the kernel, then the call's host time
        e0.record(st)
        rc = launch(offsets, out, ws, st.cuda_stream)
        t1 = time.perf_counter()
        e1.record(st)
        torch.cuda.synchronize()
        assert rc == 0, L.fgq_last_error()
        best_dev = min(best_dev, e0.elapsed_time(e1))
        best_call = min(best_call, (t1 - t0) * 1e3)
    wall = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        host_call()
        wall = min(wall, (time.perf_counter() - t0) * 1e3)
    print(f"{label:34s} rules {len(n_i):8d} nodes {total:10d}  stream span {best_dev:8.3f} ms "
          f"(host planning+enqueue {best_call:8.3f} ms; kernel ~ {best_dev - best_call:8.3f} ms)  "
          f"{len(n_i) / best_dev / 1e3:8.2f} Mrules/s {total / best_dev / 1e3:9.1f} Mnodes/s   "
          f"host-returning call {wall:8.2f} ms", flush=True)


def ptr(a):
    return a.ctypes.data


# Jacobi: n uniform on [2, 100], random (alpha, beta)
n_j = rng.integers(2, 101, NR).astype(np.int32)
a_j, b_j = rng.uniform(-0.9, 4.9, NR), rng.uniform(-0.9, 4.9, NR)
run("gaussjacobi_batch n in [2,100]", n_j, L.fgq_gaussjacobi_batch_workspace_bytes(NR),
    lambda off, out, ws, s: L.fgq_gaussjacobi_batch_device(NR, ptr(n_j), ptr(a_j), ptr(b_j), ptr(off),
                                                           out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(), s),
    lambda: fgq_b200.gaussjacobi_batch(n_j, a_j, b_j))

# Hermite: n uniform on [2, 200]
n_h = rng.integers(2, 201, NR).astype(np.int32)
run("gausshermite_batch n in [2,200]", n_h, L.fgq_gausshermite_batch_workspace_bytes(NR),
    lambda off, out, ws, s: L.fgq_gausshermite_batch_device(NR, ptr(n_h), ptr(off), 0, out[0].data_ptr(),
                                                            out[1].data_ptr(), ws.data_ptr(), s),
    lambda: fgq_b200.gausshermite_batch(n_h))

# Laguerre: n uniform on [2, 127], alpha uniform
n_l = rng.integers(2, 128, NR).astype(np.int32)
a_l = rng.uniform(-0.9, 4.9, NR)
run("gausslaguerre_batch n in [2,127]", n_l, L.fgq_gausslaguerre_batch_workspace_bytes(NR),
    lambda off, out, ws, s: L.fgq_gausslaguerre_batch_device(NR, ptr(n_l), ptr(a_l), ptr(off), out[0].data_ptr(),
                                                             out[1].data_ptr(), ws.data_ptr(), s),
    lambda: fgq_b200.gausslaguerre_batch(n_l, a_l))
