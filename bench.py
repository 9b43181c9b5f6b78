#!/usr/bin/env python
"""bench.py — headline benchmark: gausslegendre Gnodes/s at n = 1e9, device-resident, sharded
over N B200s (BASELINE.json `metric`, config[1]).

  python bench.py --gpus 1 --steps K --warmup W
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU arm: the C oracle port on the host cores

One step = one full generation of the n-point rule into HBM (no inputs: every node is a
function of its index; 16 bytes written per node).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gausslegendre_gnodes_per_s"
UNIT = "Gnodes/s"
N_DEFAULT = 10 ** 9
CPU_SAMPLE_N = 5 * 10 ** 8     # bounded single-thread CPU sample: a full gausslegendre(5e8), same branch as 1e9 (~8 s)
CPU_FULL_N_MIN_FREE_GB = 40    # the oracle at n = 1e9 holds x, w (16 GB) + two half-range temporaries (8 GB)
PUBLISHED_JULIA_GNODES = 0.0428  # docs/src/benchmark.md:21-22 (n = 1e9+1, hardware unstated)
NVLINK_GBS_PER_DIR = 900.0     # NVLink 5 per GPU and direction (B200_PROFILING.md)


def make_config(n: int) -> dict:
    """The workload both arms (--impl b200 / reference) are timed on: BASELINE.json configs[1]."""
    return {"workload": f"gausslegendre(n={n}) Float64 nodes+weights, Bogaert asymptotic path (BASELINE.json configs[1])",
            "n": n,
            "l2": "each step writes 16*n bytes (>> 126 MB L2) and reads nothing; no flush needed"}


def free_ram_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def ncu_traffic(n: int, world: int, mode: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the tail kernel, per launch, read from the committed
    `ncu --set full` raw page of the same command (profiles/): never a number measured under this run."""
    import csv
    path = os.path.join(ROOT, "profiles", "r02_legendre_1e9_ncu_raw.csv")
    if not (n == 10 ** 9 and world == 1 and mode == "pair" and os.path.exists(path)):
        return None, "no committed ncu capture for this configuration"
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        kcol = hdr.index("Kernel Name")
        rcol, wcol = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        best = None
        for r in rows[2:]:
            if "legendre_asy_kernel" in r[kcol]:
                tot = float(r[rcol].replace(",", "")) * scale.get(units[rcol], 1.0) + \
                      float(r[wcol].replace(",", "")) * scale.get(units[wcol], 1.0)
                best = tot if best is None else max(best, tot)   # the tail launch is the big one
        return best, "profiles/r02_legendre_1e9_ncu_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
    except Exception as e:  # noqa: BLE001
        return None, f"unreadable ncu capture: {e!r}"


def fp64_counts():
    """Per-config executed FP64 instruction counts (ncu smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on,
    summed over the launches of one call), committed as profiles/r02_fp64_counts.json by tools/r2_fp64_counts.py."""
    path = os.path.join(ROOT, "profiles", "r02_fp64_counts.json")
    try:
        return json.load(open(path))
    except Exception:  # noqa: BLE001
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json, torch copy_ read+write)"
        except Exception:  # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_oracle_rate(n_sample: int, threads: int, reps: int = 1):
    """Gnodes/s of the C oracle (oracle/legendre_oracle.c) on a full gausslegendre(n_sample)."""
    import numpy as np
    import oracle
    used = oracle.set_threads(threads)
    x = np.zeros(n_sample)  # pre-faulted outputs
    w = np.zeros(n_sample)
    L = oracle.lib()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = L.oracle_gausslegendre(n_sample, oracle._dp(x), oracle._dp(w))
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    return n_sample / best / 1e9, used, best


def run_reference(args, rank, world):
    """CPU arm: the oracle port (Julia is not installed on the box, so the reference itself cannot run), all host
    threads.  Each step is one full gausslegendre(n) at the metric's own n = 1e9 when the box has the memory for
    it, otherwise the stated bounded sample (same branch: 1-term nodes, 0-term weights)."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    free = free_ram_gb()
    n_step = args.n if (args.cpu_n is None and free >= CPU_FULL_N_MIN_FREE_GB) else (args.cpu_n or CPU_SAMPLE_N)
    n_step = min(n_step, args.n)
    warm = min(args.warmup, 3)   # OpenMP pool + first-touch page faults; a step takes seconds
    for _ in range(warm):
        cpu_oracle_rate(n_step, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, used, _ = cpu_oracle_rate(n_step, 0)
    dt = time.perf_counter() - t0
    value = n_step * args.steps / dt / 1e9
    rate1, _, _ = cpu_oracle_rate(min(CPU_SAMPLE_N, n_step), 1)
    sample = (f"each step = full oracle gausslegendre(n={n_step}) on the host"
              + ("" if n_step == args.n else f" (bounded sample of n={args.n}: same branch, 1-term nodes / 0-term weights; "
                                             f"{free:.0f} GB free < {CPU_FULL_N_MIN_FREE_GB} GB needed for the full n)")
              + f", OpenMP over {used} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (none read: nodes are generated from their index)",
        "config": make_config(args.n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                         "sample_n": n_step, "single_thread_value": rate1, "host_cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "published_reference": {"value": PUBLISHED_JULIA_GNODES, "unit": UNIT,
                                "what": "Julia @btime gausslegendre(1000000001), docs/src/benchmark.md:21-22, hardware unstated"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_other_configs(fgq_b200, cpu: bool):
    """The other BASELINE.json configs (parity-test cases, not the bench line): device-resident time
    per call with CUDA events (best of 3 after a warm-up), Mnodes/s, and the CPU oracle on a bounded
    sample beside it.  Reported under "other_configs"; the headline metric is untouched."""
    import numpy as np
    import torch
    out = {}

    def dev_time(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best

    def cpu_time(fn):
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    L = fgq_b200.lib()
    chk = fgq_b200._lib.check
    # C1: gausslegendre(1e5), the README example
    n = 100_000
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    ms = dev_time(lambda: chk(L.fgq_gausslegendre_device(n, 0, n, x.data_ptr(), w.data_ptr(), sp)), reps=10)
    t0 = time.perf_counter()
    for _ in range(20):
        fgq_b200.gausslegendre(n)
    host_ms = (time.perf_counter() - t0) / 20 * 1e3
    out["C1_gausslegendre_1e5"] = {"device_ms": ms, "mnodes_per_s": n / ms / 1e3, "host_call_ms": host_ms,
                                   "published_julia_ms": 2.192}
    # consumer fusion on the headline config: generate-and-reduce, nothing stored (FP64-bound)
    nf = 10 ** 9
    ms = dev_time(lambda: fgq_b200.gausslegendre_moments_fused(nf))
    sf = fgq_b200.gausslegendre_moments_fused(nf).cpu().numpy()
    out["fused_moments_gausslegendre_1e9"] = {"device_ms": ms, "gnodes_per_s": nf / ms / 1e6,
                                              "sum_w_minus_2": float(sf[0] - 2), "sum_wx2_minus_2_3": float(sf[1] - 2 / 3)}
    # C2 with odd n (docs/src/benchmark.md:21 times gausslegendre(1000000001)): the mirrored pair of a
    # thread straddles a 16-byte boundary, handled with one warp shuffle
    no = 10 ** 9 + 1
    sh_odd = fgq_b200.gausslegendre_device(no, mode="pair")
    ms = dev_time(lambda: fgq_b200.gausslegendre_device(no, mode="pair", out=sh_odd), reps=5)
    out["C2_odd_gausslegendre_1e9p1"] = {"device_ms": ms, "gnodes_per_s": no / ms / 1e6}
    del sh_odd
    torch.cuda.empty_cache()

    def host_wall(fn, reps=3):
        fn()
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best * 1e3

    # C3: gaussjacobi(1e7, 0.3, -0.7)
    n = 10_000_000
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    ms = dev_time(lambda: chk(L.fgq_gaussjacobi_device(n, 0.3, -0.7, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp)))
    out["C3_gaussjacobi_1e7"] = {"device_ms": ms, "mnodes_per_s": n / ms / 1e3,
                                 "sum_w_err": float(w.sum().item() - 4.554443087962173),
                                 "host_call_ms": host_wall(lambda: fgq_b200.gaussjacobi(10_000_000, 0.3, -0.7))}
    # C4: gausshermite(1e6), gausslaguerre(1e6)
    n = 1_000_000
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    ms = dev_time(lambda: chk(L.fgq_gausshermite_device(n, 0, 0, n, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp)))
    out["C4_gausshermite_1e6"] = {"device_ms": ms, "mnodes_per_s": n / ms / 1e3,
                                  "sum_w_err": float(w.sum().item() - np.sqrt(np.pi)),
                                  "host_call_ms": host_wall(lambda: fgq_b200.gausshermite(1_000_000))}
    ws = torch.empty(int(L.fgq_gausslaguerre_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    ms = dev_time(lambda: chk(L.fgq_gausslaguerre_device(n, 0.0, x.data_ptr(), w.data_ptr(), ws.data_ptr(), sp)))
    out["C4_gausslaguerre_1e6"] = {"device_ms": ms, "mnodes_per_s": n / ms / 1e3, "sum_w_err": float(w.sum().item() - 1.0),
                                   "host_call_ms": host_wall(lambda: fgq_b200.gausslaguerre(1_000_000))}
    # C5: 1e6 small rules, n in [2, 60]
    n_i = np.random.default_rng(0).integers(2, 61, size=1_000_000).astype(np.int32)
    off = np.zeros(len(n_i) + 1, dtype=np.int64)
    np.cumsum(n_i, out=off[1:])
    tot = int(off[-1])
    dn = torch.from_numpy(n_i).cuda()
    do = torch.from_numpy(off).cuda()
    x = torch.empty(tot, dtype=torch.float64, device="cuda")
    w = torch.empty(tot, dtype=torch.float64, device="cuda")
    ms = dev_time(lambda: chk(L.fgq_gausslegendre_batch_device(len(n_i), dn.data_ptr(), do.data_ptr(), x.data_ptr(), w.data_ptr(), sp)))
    out["C5_batch_1e6_rules"] = {"device_ms": ms, "mrules_per_s": len(n_i) / ms / 1e3, "mnodes_per_s": tot / ms / 1e3, "nodes": tot}
    # truncated Gauss-Hermite rule (only the centre block of non-zero weights is computed)
    n = 1_000_000
    nb = int(L.fgq_gausshermite_reduced_bound(n))
    xr = torch.empty(nb, dtype=torch.float64, device="cuda")
    wr = torch.empty(nb, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n)) + 16, dtype=torch.uint8, device="cuda")
    ms = dev_time(lambda: chk(L.fgq_gausshermite_reduced_device(n, 0, xr.data_ptr(), wr.data_ptr(), ws.data_ptr(), sp, None, None)))
    out["reduced_gausshermite_1e6"] = {"device_ms": ms, "entries": nb, "sum_w_err": float(wr.sum().item() - np.sqrt(np.pi)),
                                       "host_call_ms": host_wall(lambda: fgq_b200.gausshermite_reduced(1_000_000))}
    # batches of small rules of the other families (1e5 rules each, host-side planning included
    # in host_call_ms; device_ms spans descriptor upload + kernel on the stream)
    nr = 100_000
    rng = np.random.default_rng(0)

    def batch_entry(n_i, wsbytes, launch, host_call):
        off = np.zeros(len(n_i) + 1, dtype=np.int64)
        np.cumsum(n_i, out=off[1:])
        tot = int(off[-1])
        xb = torch.empty(tot, dtype=torch.float64, device="cuda")
        wb = torch.empty(tot, dtype=torch.float64, device="cuda")
        wsb = torch.empty(int(wsbytes), dtype=torch.uint8, device="cuda")
        ms = dev_time(lambda: chk(launch(off, xb, wb, wsb)))
        return {"rules": len(n_i), "nodes": tot, "stream_ms_incl_host_planning": ms, "mrules_per_s": len(n_i) / ms / 1e3,
                "host_call_ms": host_wall(host_call)}

    n_j = rng.integers(2, 101, nr).astype(np.int32)
    a_j, b_j = rng.uniform(-0.9, 4.9, nr), rng.uniform(-0.9, 4.9, nr)
    out["batch_gaussjacobi_1e5_rules"] = batch_entry(
        n_j, L.fgq_gaussjacobi_batch_workspace_bytes(nr),
        lambda off, xb, wb, wsb: L.fgq_gaussjacobi_batch_device(nr, n_j.ctypes.data, a_j.ctypes.data, b_j.ctypes.data,
                                                                off.ctypes.data, xb.data_ptr(), wb.data_ptr(),
                                                                wsb.data_ptr(), sp),
        lambda: fgq_b200.gaussjacobi_batch(n_j, a_j, b_j))
    n_h = rng.integers(2, 201, nr).astype(np.int32)
    out["batch_gausshermite_1e5_rules"] = batch_entry(
        n_h, L.fgq_gausshermite_batch_workspace_bytes(nr),
        lambda off, xb, wb, wsb: L.fgq_gausshermite_batch_device(nr, n_h.ctypes.data, off.ctypes.data, 0, xb.data_ptr(),
                                                                 wb.data_ptr(), wsb.data_ptr(), sp),
        lambda: fgq_b200.gausshermite_batch(n_h))
    n_l = rng.integers(2, 128, nr).astype(np.int32)
    a_l = rng.uniform(-0.9, 4.9, nr)
    out["batch_gausslaguerre_1e5_rules"] = batch_entry(
        n_l, L.fgq_gausslaguerre_batch_workspace_bytes(nr),
        lambda off, xb, wb, wsb: L.fgq_gausslaguerre_batch_device(nr, n_l.ctypes.data, a_l.ctypes.data, off.ctypes.data,
                                                                  xb.data_ptr(), wb.data_ptr(), wsb.data_ptr(), sp),
        lambda: fgq_b200.gausslaguerre_batch(n_l, a_l))
    if cpu:
        import oracle
        from oracle import hermite_oracle, jacobi_oracle, laguerre_oracle
        oracle.set_threads(1)
        t = cpu_time(lambda: oracle.gausslegendre(100_000))
        out["C1_gausslegendre_1e5"]["cpu_oracle_ms_1thread"] = t * 1e3
        t = cpu_time(lambda: jacobi_oracle.gaussjacobi(200_000, 0.3, -0.7))
        out["C3_gaussjacobi_1e7"]["cpu_oracle"] = {"sample_n": 200_000, "seconds": t, "mnodes_per_s": 0.2 / t, "kind": "port (numpy)"}
        t = cpu_time(lambda: hermite_oracle.gausshermite(1_000_000))
        out["C4_gausshermite_1e6"]["cpu_oracle"] = {"sample_n": 1_000_000, "seconds": t, "mnodes_per_s": 1.0 / t, "kind": "port (numpy)"}
        t = cpu_time(lambda: laguerre_oracle.gausslaguerre(1_000_000, 0.0))
        out["C4_gausslaguerre_1e6"]["cpu_oracle"] = {"sample_n": 1_000_000, "seconds": t, "mnodes_per_s": 1.0 / t, "kind": "port (numpy)"}
        t = cpu_time(lambda: oracle.gausslegendre_batch(n_i))
        out["C5_batch_1e6_rules"]["cpu_oracle_1thread"] = {"seconds": t, "mrules_per_s": 1.0 / t, "kind": "port (C)"}
        oracle.set_threads(0)
    return out


def time_e2e(args, fgq_b200, rank, world, barrier, max_over_ranks):
    """The metric through the host-returning C-ABI call, host buffers, copies inside the timed region.
    world == 1: the drop-in call gausslegendre(n) itself, into (a) pinned, pre-faulted, reused buffers — the headline
    `value` —, (b) pre-faulted pageable buffers, (c) a FRESH untouched allocation every step (what
    `Vector{Float64}(undef, n)` in julia/FGQCuda.jl is); beside them the host's own store-bandwidth roofline.
    world > 1: each rank fetches its mirror-pair shard into pinned buffers."""
    import ctypes
    import numpy as np
    import torch
    n = args.n
    L = fgq_b200.lib()
    hw = None
    if world == 1 and args.e2e_mode == "pair":
        hx = torch.empty(n, dtype=torch.float64, pin_memory=True)
        hw = torch.empty(n, dtype=torch.float64, pin_memory=True)
        out = (hx.numpy(), hw.numpy())
        call = lambda: fgq_b200.gausslegendre(n, out=out)  # noqa: E731
    elif args.e2e_mode == "pair":
        k_lo, k_hi = fgq_b200.half_index_partition(n, world)[rank]
        bufs = [torch.empty(k_hi - k_lo, dtype=torch.float64, pin_memory=True) for _ in range(4)]
        out4 = tuple(b.numpy() for b in bufs)
        call = lambda: fgq_b200.gausslegendre_host_pair(n, rank, world, out=out4)  # noqa: E731
    else:
        # both segments of the rank's mirror-pair shard by DMA (16 B/node over PCIe, no host threads)
        k_lo, k_hi = fgq_b200.half_index_partition(n, world)[rank]
        segs = [(k_lo - 1, min(k_hi - 1, (n + 1) // 2)), (n - min(k_hi - 1, n // 2), n - k_lo + 1)]
        outs = [tuple(torch.empty(b - a, dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)) for a, b in segs]

        def call():
            for (a, b), o in zip(segs, outs):
                fgq_b200.gausslegendre(n, index_range=(a, b), out=o)

    def timed(fn, steps, fresh=None):
        per = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            fn()
            per.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0), per

    for _ in range(2):  # warm-up (scratch allocation, host worker threads)
        call()
    barrier()
    st_p, st_d, st_l, st_s = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
    L.fgq_debug_drain_stats(ctypes.byref(st_p), ctypes.byref(st_d), ctypes.byref(st_l), ctypes.byref(st_s))
    cd_b, cd_v = ctypes.c_int64(0), ctypes.c_int64(0)
    L.fgq_debug_codec_stats(ctypes.byref(cd_b), ctypes.byref(cd_v))  # reset
    dt, per_call = timed(call, args.e2e_steps)
    L.fgq_debug_drain_stats(ctypes.byref(st_p), ctypes.byref(st_d), ctypes.byref(st_l), ctypes.byref(st_s))
    L.fgq_debug_codec_stats(ctypes.byref(cd_b), ctypes.byref(cd_v))
    coded = torch.tensor([float(cd_b.value), float(cd_v.value)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(coded)
    coded_bytes, coded_values = float(coded[0].item()), float(coded[1].item())
    # bytes that really crossed PCIe per step: the coded streams when the transport coding is active
    # (FGQ_HOST_CODEC, default on), otherwise the plain left half
    d2h = int(coded_bytes / args.e2e_steps) if coded_values > 0 else 8 * (n + n % 2)
    e2e = {"value": n * args.e2e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 0,
           "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
           "destination": "pinned host buffers, pre-faulted, reused every step",
           "transport": ({"coding": "second differences of the bit patterns, lossless (csrc/delta_codec.h)",
                          "bytes_per_value": coded_bytes / coded_values} if coded_values > 0 else
                         {"coding": "none"}),
           "call_s_min": min(per_call), "call_s_median": float(np.median(per_call)),
           "what": "host-returning C-ABI call: kernel + device-side packing + PCIe D2H of the coded left half of each "
                   "shard, chunk-pipelined; host threads unpack it into the left half and write the mirrored half "
                   "(+-x symmetry)",
           "mode": args.e2e_mode,
           "drain_rank0": {"pieces": st_p.value, "mirror_by_dma": st_d.value, "landed_before_pickup": st_l.value,
                           "last_dma_share": st_s.value / 1000.0},
           "check_sum_w_minus_2": float(np.float64(hw.sum().item()) - 2.0) if hw is not None else None}
    if world == 1 and args.e2e_mode == "pair":
        # the host's own roofline: every result byte is written once by a host thread (non-temporal stores)
        g = ctypes.c_double(0.0)
        if L.fgq_measure_host_store_peak(out[0].ctypes.data, n, 0, ctypes.byref(g)) == 0 and g.value > 0:
            ach = 16.0 * n / min(per_call) / 1e9
            e2e["host_roofline"] = {"bound": "host memory stores", "achieved_gbs": ach, "peak_gbs": g.value,
                                    "frac": ach / g.value, "algorithmic_bytes_per_node": 16,
                                    "peak_source": "fgq_measure_host_store_peak: the library's host threads streaming "
                                                   "non-temporal stores into the same pinned buffer",
                                    "host_threads": int(os.environ.get("FGQ_HOST_THREADS", 0)) or min(16, os.cpu_count() or 16)}
        del hx, hw, out
        k = max(1, min(args.e2e_steps, 10))
        px, pw = np.zeros(n), np.zeros(n)        # pageable, pre-faulted
        fgq_b200.gausslegendre(n, out=(px, pw))
        dtp, per_p = timed(lambda: fgq_b200.gausslegendre(n, out=(px, pw)), k)
        e2e["e2e_pageable"] = {"value": n * k / dtp / 1e9, "unit": UNIT, "steps": k, "call_s_min": min(per_p),
                               "destination": "pageable numpy arrays, pre-faulted, reused"}
        del px, pw
        k = max(1, min(args.e2e_steps, 5))
        dtf, per_f = timed(lambda: fgq_b200.gausslegendre(n), k)   # np.empty inside: untouched pages every step
        e2e["e2e_fresh"] = {"value": n * k / dtf / 1e9, "unit": UNIT, "steps": k, "call_s_min": min(per_f),
                            "destination": "a fresh untouched allocation per step (np.empty = Vector{Float64}(undef, n)): "
                                           "first-touch page faults inside the timed region"}
    return e2e


def attach_rooflines(other, counts, dfma_tf, hbm_gbs):
    """Every other_configs entry gets its own roofline: executed FP64 work per node from the committed ncu
    instruction counts (profiles/r02_fp64_counts.json) against the measured DFMA peak, and 16 B per node of
    algorithmic store traffic against the measured HBM peak."""
    if not other:
        return
    for key, ent in other.items():
        ms = ent.get("device_ms") or ent.get("stream_ms_incl_host_planning")
        nodes = ent.get("nodes") or {"C1_gausslegendre_1e5": 1e5, "fused_moments_gausslegendre_1e9": 1e9,
                                     "C2_odd_gausslegendre_1e9p1": 1e9 + 1, "C3_gaussjacobi_1e7": 1e7,
                                     "C4_gausshermite_1e6": 1e6, "C4_gausslaguerre_1e6": 1e6,
                                     "reduced_gausshermite_1e6": ent.get("entries")}.get(key)
        if not ms or not nodes:
            continue
        c = counts.get(key, {})
        r = {"bytes_per_node": 0 if key.startswith("fused") else 16,
             "achieved_gbs": (0 if key.startswith("fused") else 16.0) * nodes / (ms * 1e-3) / 1e9}
        r["frac_of_hbm"] = r["achieved_gbs"] / hbm_gbs
        if "flops_per_node" in c:
            tf = c["flops_per_node"] * nodes / (ms * 1e-3) / 1e12
            r.update({"flops_per_node": c["flops_per_node"], "flops_source": c.get("source"),
                      "achieved_tflops": tf, "frac_of_dfma_peak": tf / dfma_tf if dfma_tf else None})
        r["bound"] = "hbm" if r["frac_of_hbm"] >= (r.get("frac_of_dfma_peak") or 0) else "fp64"
        ent["roofline"] = r


def run_capi(args):
    """ONE process, --gpus devices, nothing but `ccall`-able symbols of libfgq_cuda.so (ctypes here): fgq_init,
    fgq_gausslegendre_device_sharded, fgq_timer_*, fgq_allgather, fgq_gausslegendre_device_replicated,
    fgq_gausslegendre_host_multi.  Same metric, same kernels, same partition as the torchrun arm."""
    import ctypes
    import numpy as np
    import torch
    import fgq_b200
    m = fgq_b200.multi
    n = args.n
    nd = m.init(ndev=args.gpus)
    rule = m.gausslegendre_sharded(n)            # allocates, fills, reduces sum w / sum w x^2
    sums = rule.sums
    for _ in range(max(args.warmup, 3)):
        m.gausslegendre_sharded(n, out=rule, sums=False)
    m.synchronize()
    samplers = [ClockSampler(d) for d in range(nd)]
    for sp in samplers:
        sp.start()
    launches0 = fgq_b200.launch_count()
    m.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.gausslegendre_sharded(n, out=rule, sums=False)
    ms, per_dev = m.timer_stop()
    wall = time.perf_counter() - t0
    launches = fgq_b200.launch_count() - launches0
    clk = [sp.stop() for sp in samplers]
    clocks = {"sm_mhz": min(c["sm_mhz"] for c in clk if c["sm_mhz"]) if any(c["sm_mhz"] for c in clk) else None,
              "sm_max_mhz": clk[0]["sm_max_mhz"], "reasons": sorted(set(r for c in clk for r in c["reasons"])),
              "samples": sum(c["samples"] for c in clk), "per_device_sm_mhz": [c["sm_mhz"] for c in clk]}
    value = n * args.steps / (ms * 1e-3) / 1e9
    allgather = None
    if nd > 1 and not args.no_allgather:
        inbound = 16.0 * ((n + 1) // 2) * (nd - 1) / nd
        allgather = {"inbound_bytes_per_gpu": inbound}
        rep = None
        for transport in ("push", "ce", "nccl"):
            os.environ["FGQ_ALLGATHER"] = transport
            try:
                rep = m.allgather(rule, out=rep)   # warm-up / allocation
                m.synchronize()
                m.timer_start()
                for _ in range(3):
                    m.allgather(rule, out=rep)
                t_ms, _ = m.timer_stop()
                t_ms /= 3
                allgather[transport] = {"ms": t_ms, "inbound_gbs_per_gpu": inbound / (t_ms * 1e-3) / 1e9,
                                        "frac_of_nvlink_900": inbound / (t_ms * 1e-3) / 1e9 / NVLINK_GBS_PER_DIR}
            except Exception as e:  # noqa: BLE001
                allgather[transport] = {"error": repr(e)}
        os.environ.pop("FGQ_ALLGATHER", None)
        if rep is not None:
            xs, ws = rep.tensors(nd - 1)
            allgather["replica_sum_w_minus_2"] = float(ws.sum().item() - 2.0)
            allgather["replica_symmetric"] = bool(torch.equal(xs[: n // 2], -xs[n - n // 2:].flip(0)))
            m.gausslegendre_replicated(n, out=rep)
            m.synchronize()
            m.timer_start()
            for _ in range(3):
                m.gausslegendre_replicated(n, out=rep)
            t_ms, _ = m.timer_stop()
            allgather["replicate_by_generation_ms"] = t_ms / 3
            torch.cuda.set_device(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fgq_b200._lib.check(fgq_b200.lib().fgq_gausslegendre_mirror_device(n, rep.x[0], rep.w[0],
                                                                                   torch.cuda.current_stream().cuda_stream))
            e1.record()
            torch.cuda.synchronize()
            allgather["mirror_ms"] = e0.elapsed_time(e1) / 3
            for tr in ("push", "ce", "nccl"):
                if "ms" in allgather.get(tr, {}):
                    tm = allgather[tr]["ms"] - allgather["mirror_ms"]
                    allgather[tr]["transfer_ms"] = tm
                    allgather[tr]["transfer_gbs_per_gpu"] = inbound / (tm * 1e-3) / 1e9
            allgather["what"] = ("left segments shipped (push: one kernel per device storing into every peer replica "
                                 "over NVLink; ce: cudaMemcpyPeerAsync; nccl: grouped send/recv), right halves mirrored locally")
            rep.free()
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(n, dtype=torch.float64, pin_memory=True)
        hw = torch.empty(n, dtype=torch.float64, pin_memory=True)
        out = (hx.numpy(), hw.numpy())
        for _ in range(2):
            m.gausslegendre_host_multi(n, out=out)
        cd_b, cd_v = ctypes.c_int64(0), ctypes.c_int64(0)
        fgq_b200.lib().fgq_debug_codec_stats(ctypes.byref(cd_b), ctypes.byref(cd_v))
        per = []
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            t1 = time.perf_counter()
            m.gausslegendre_host_multi(n, out=out)
            per.append(time.perf_counter() - t1)
        dt = time.perf_counter() - t0
        fgq_b200.lib().fgq_debug_codec_stats(ctypes.byref(cd_b), ctypes.byref(cd_v))
        g = ctypes.c_double(0.0)
        fgq_b200.lib().fgq_measure_host_store_peak(out[0].ctypes.data, n, 0, ctypes.byref(g))
        e2e = {"value": n * args.e2e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(cd_b.value / args.e2e_steps) if cd_v.value else 8 * (n + n % 2),
               "steps": args.e2e_steps, "call_s_min": min(per), "call_s_median": float(np.median(per)),
               "what": "fgq_gausslegendre_host_multi: every device generates, packs and ships its half-index range over its "
                       "own PCIe link; the host's threads are shared out between the devices",
               "destination": "pinned host buffers, pre-faulted, reused every step",
               "check_sum_w_minus_2": float(np.float64(hw.sum().item()) - 2.0),
               "host_roofline": {"achieved_gbs": 16.0 * n / min(per) / 1e9, "peak_gbs": g.value,
                                 "frac": (16.0 * n / min(per) / 1e9) / g.value if g.value else None}}
    cfg = make_config(n)
    detail = {"shard_mode": "pair", "driver": "capi: one process, multi-GPU C ABI (include/fgq.h)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nd, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (none read: nodes are generated from their index)",
            "config": cfg, "config_detail": detail, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "device_ms_per_device": [v / args.steps for v in per_dev], "host_wall_ms_per_step": wall / args.steps * 1e3,
            "allgather": allgather,
            "checks": {"sum_w_minus_2": sums[0] - 2.0, "sum_wx2_minus_2_3": sums[1] - 2.0 / 3.0}}
    print(json.dumps(line), flush=True)
    rule.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--mode", default="pair", choices=["pair", "contiguous"])
    ap.add_argument("--cpu-n", type=int, default=None,
                    help="nodes per reference-arm step (default: n itself when the box has >= 40 GB free, else 5e8)")
    ap.add_argument("--driver", default="torch", choices=["torch", "capi"],
                    help="torch: one process per GPU (torchrun); capi: ONE process drives --gpus devices through the "
                         "multi-GPU C ABI (fgq_init / fgq_gausslegendre_device_sharded / fgq_allgather)")
    ap.add_argument("--no-allgather", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None, help="default: --steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", choices=["pair", "dma"], default="pair",
                    help="pair: left half over PCIe, mirrored by host threads; dma: both halves over PCIe, no host work")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.e2e_steps is None:
        args.e2e_steps = args.steps
    if args.driver == "capi":
        run_capi(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if world > 1:
        # one process per GPU on a shared host: split its cores between the ranks' mirroring threads
        os.environ.setdefault("FGQ_HOST_THREADS", str(max(2, min(16, (os.cpu_count() or 16) // world))))
    import fgq_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    L = fgq_b200.lib()
    shard = fgq_b200.gausslegendre_device(n, rank=rank, world=world, mode=args.mode)  # allocates
    torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fgq_b200.gausslegendre_device(n, rank=rank, world=world, mode=args.mode, out=shard)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = fgq_b200.launch_count()
    barrier()
    ev[0].record()
    for i in range(args.steps):
        fgq_b200.gausslegendre_device(n, rank=rank, world=world, mode=args.mode, out=shard)
        ev[i + 1].record()
    barrier()
    launches = fgq_b200.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    total_ms_max = max_over_ranks(total_ms)
    value = n * args.steps / (total_ms_max * 1e-3) / 1e9

    # invariant check of the last step's output: sum w = 2, sum w x^2 = 2/3 (all-reduced: 2 doubles over NCCL)
    sums = fgq_b200.rule_moments_device(shard)
    if world > 1:
        dist.all_reduce(sums)
    sums = sums.cpu().numpy()

    # ---- optional collective, reported separately (north star: "NCCL over NVLink is used only for an optional
    # allgather into a replicated array"): one all_gather_into_tensor per array of the LEFT segments + a local
    # mirror kernel; beside it the same replicated result by generation on every rank (no transfer at all).
    allgather = None
    if world > 1 and not args.no_allgather:
        from fgq_b200 import distributed as fd
        xr = torch.empty(n, dtype=torch.float64, device="cuda")
        wr = torch.empty(n, dtype=torch.float64, device="cuda")
        st = {}
        fd.allgather_rule(shard, out=(xr, wr), stats=st)   # warm-up (NCCL channel set-up)
        barrier()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fd.allgather_rule(shard, out=(xr, wr))
        e1.record()
        barrier()
        ag_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        # checksum of the replica against the sharded rule: same sum w, sum w x^2 (to rounding), symmetric, ordered
        full = fgq_b200.LegendreShard(n, 0, 1, "contiguous", [(0, n, xr, wr)])
        rs = fgq_b200.rule_moments_device(full).cpu().numpy()
        k_lo, k_hi = fgq_b200.half_index_partition(n, world)[rank]
        own_ok = bool(torch.equal(xr[k_lo - 1:k_hi - 1], shard.segments[0][2])) and \
            bool(torch.equal(wr[n - (k_hi - 1):n - (k_lo - 1)].flip(0)[:shard.segments[1][3].numel()],
                             shard.segments[1][3].flip(0)))
        inbound = 16.0 * ((n + 1) // 2) * (world - 1) / world   # bytes arriving at each GPU: left halves of the peers
        # the alternative: every rank generates the whole rule itself
        e0.record()
        for _ in range(reps):
            m_all = (n + 1) // 2
            fgq_b200._lib.check(L.fgq_gausslegendre_device_pair(n, 1, m_all + 1, xr.data_ptr(), wr.data_ptr(),
                                                                xr.data_ptr() + 8 * (n - m_all), wr.data_ptr() + 8 * (n - m_all),
                                                                torch.cuda.current_stream().cuda_stream))
        e1.record()
        barrier()
        regen_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        e0.record()
        for _ in range(reps):   # the local +-x mirror alone (part of `ms`)
            fgq_b200._lib.check(L.fgq_gausslegendre_mirror_device(n, xr.data_ptr(), wr.data_ptr(),
                                                                  torch.cuda.current_stream().cuda_stream))
        e1.record()
        barrier()
        mirror_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        allgather = {"ms": ag_ms, "mirror_ms": mirror_ms, "transfer_ms": ag_ms - mirror_ms,
                     "transfer_gbs_per_gpu": inbound / ((ag_ms - mirror_ms) * 1e-3) / 1e9,
                     "collectives_per_call": st.get("collectives"),
                     "inbound_bytes_per_gpu": inbound, "inbound_gbs_per_gpu": inbound / (ag_ms * 1e-3) / 1e9,
                     "frac_of_nvlink_900": inbound / (ag_ms * 1e-3) / 1e9 / NVLINK_GBS_PER_DIR,
                     "what": "all_gather_into_tensor (NCCL) of the left segments, x and w, + local +-x mirror kernel; "
                             "the replica ends up complete on every rank",
                     "replica_sum_w_minus_2": float(rs[0] - 2.0), "replica_sum_wx2_minus_2_3": float(rs[1] - 2.0 / 3.0),
                     "replica_matches_own_shard": own_ok,
                     "replicate_by_generation_ms": regen_ms,
                     "note": "generating the whole rule on every rank is cheaper than gathering it"}
        del xr, wr, full
        torch.cuda.empty_cache()

    # ---- end-to-end through the host-returning C-ABI call: device generation + D2H into HOST buffers every step
    # (h2d = 0: the only input is the integer n).
    e2e = None
    if not args.no_e2e:
        e2e = time_e2e(args, fgq_b200, rank, world, barrier, max_over_ranks)

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        # dominant kernel: legendre_asy_kernel<0,0,false> (the McMahon-tail launch); time it alone
        k_lo, k_hi = fgq_b200.half_index_partition(n, world)[rank]
        k_tail = max(k_lo, 13193)  # odd, like every shard boundary: keeps 16-byte alignment
        tail_nodes = 2 * (k_hi - k_tail)
        reps = 20
        xl, wl = shard.segments[0][2], shard.segments[0][3]
        if args.mode == "pair" and k_hi - k_tail > 0:
            xr, wr = shard._right_full
            off = k_tail - k_lo
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sp = torch.cuda.current_stream().cuda_stream
            e0.record()
            for _ in range(reps):
                # right segment of [k_tail, k_hi) ends where the full one ends: same base pointers
                fgq_b200._lib.check(L.fgq_gausslegendre_device_pair(
                    n, k_tail, k_hi, xl.data_ptr() + 8 * off, wl.data_ptr() + 8 * off,
                    xr.data_ptr(), wr.data_ptr(), sp))
            e1.record()
            torch.cuda.synchronize()
            kern_ms = e0.elapsed_time(e1) / reps
        else:
            kern_ms = statistics.mean(step_ms)
            tail_nodes = shard.count
        achieved = 16.0 * tail_nodes / (kern_ms * 1e-3) / 1e9
        dfma_tf, store_gbs = fgq_b200.measure_peaks()
        traffic, traffic_src = ncu_traffic(n, world, args.mode)
        counts = fp64_counts()
        c2 = counts.get("C2_gausslegendre_1e9", {})
        flops_node = c2.get("flops_per_node", 2 * 109 / 2.0)   # executed FP64 instr/node x (1 + fma share); fallback: SASS count
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "legendre_asy_kernel<NT=0,WT=0,HEAD=false>", "kernel_ms": kern_ms,
            "algorithmic_bytes_per_node": 16, "nodes_per_launch": tail_nodes,
            "peak_source": peaks_src,
            "store_only_peak_gbs": store_gbs, "frac_of_store_peak": achieved / store_gbs if store_gbs else None,
            "fp64": {"algorithmic_flops_per_node": 22, "executed_flops_per_node": flops_node,
                     "executed_source": c2.get("source", "static SASS count of the two bulk paths (100 / 118 FP64 instructions per 4 nodes)"),
                     "achieved_tflops_algorithmic": 22.0 * tail_nodes / (kern_ms * 1e-3) / 1e12,
                     "achieved_tflops_executed": flops_node * tail_nodes / (kern_ms * 1e-3) / 1e12,
                     "dfma_peak_tflops": dfma_tf,
                     "frac_executed": (flops_node * tail_nodes / (kern_ms * 1e-3) / 1e12) / dfma_tf if dfma_tf else None},
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            cn = args.cpu_n or CPU_SAMPLE_N
            r_all, used, t_all = cpu_oracle_rate(cn, 0)
            r_one, _, t_one = cpu_oracle_rate(cn, 1)
            cpu_baseline = {
                "value": r_one, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"full oracle gausslegendre(n={cn}) (same branch as n=1e9), one thread like the reference; {t_one:.2f} s",
                "all_cores_value": r_all, "all_cores": used,
                "published_julia": {"value": PUBLISHED_JULIA_GNODES, "what": "docs/src/benchmark.md:21-22, hardware unstated"},
            }
        other = None
        if world == 1 and not args.no_other_configs:
            del shard
            torch.cuda.empty_cache()
            other = time_other_configs(fgq_b200, cpu=not args.no_cpu_baseline)
            attach_rooflines(other, counts, dfma_tf, peaks["hbm_gbs"])
        cfg = make_config(n)   # identical to the reference arm's
        detail = {"shard_mode": args.mode, "driver": "torch.distributed, one process per GPU",
                  "sharding": f"{args.mode}-sharded over {world} GPU(s), device-resident, no collective on the path"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (none read: nodes are generated from their index)",
            "config": cfg, "config_detail": detail,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "allgather": allgather,
            "checks": {"sum_w_minus_2": float(sums[0] - 2.0), "sum_wx2_minus_2_3": float(sums[1] - 2.0 / 3.0)},
            "step_ms_median": statistics.median(step_ms), "step_ms_min": min(step_ms),
            "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
