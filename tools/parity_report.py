"""Full-range parity report for the headline config: every half-index of gausslegendre(n) on the GPU
against the C oracle (twin mode = bit-for-bit expectation; libm mode = glibc sin/cos/tan, an
independent < 1 ulp libm; julia mode = oracle/julia_libm.h, the restated fdlibm-derived algorithms
Julia Base uses, i.e. the functions the reference calls).  Writes a markdown table of ulp histograms.
usage: python tools/parity_report.py [n] [out.md]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fgq_b200  # noqa: E402
import oracle  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"parity_legendre_{n}.md")
oracle.set_threads(0)
sh = fgq_b200.gausslegendre_device(n, mode="pair")
torch.cuda.synchronize()
(l0, l1, xl, wl), (r0, r1, xr, wr) = sh.segments
m = (n + 1) // 2
mr = n // 2
assert bool(torch.equal(xl[:mr], -xr.flip(0))) and bool(torch.equal(wl[:mr], wr.flip(0))), "mirror symmetry"
CH = 25_000_000
hx = {md: np.zeros(8, dtype=np.int64) for md in ("libm", "julia")}
hw = {md: np.zeros(8, dtype=np.int64) for md in ("libm", "julia")}
twin_bad = 0
t0 = time.time()
for k0 in range(1, m + 1, CH):
    k1 = min(k0 + CH - 1, m)
    gx = (-xl[k0 - 1:k1]).cpu().numpy()
    gw = wl[k0 - 1:k1].cpu().numpy()
    if n % 2 and k1 == m:
        gx[-1] = 0.0  # centre node is forced to 0.0 on both sides
    xt, wt = oracle.legendre_asy_half(n, k0, k1, twin=True)
    if n % 2 and k1 == m:
        xt[-1] = 0.0
    twin_bad += int(np.count_nonzero(gx != xt) + np.count_nonzero(gw != wt))
    for md in ("libm", "julia"):
        xo, wo = oracle.legendre_asy_half(n, k0, k1, twin=md)
        if n % 2 and k1 == m:
            xo[-1] = 0.0
        nz = xo != 0
        ux = oracle.ulp_diff(gx[nz], xo[nz]).astype(np.int64)
        uw = oracle.ulp_diff(gw, wo).astype(np.int64)
        hx[md] += np.bincount(np.minimum(ux, 7), minlength=8)
        hw[md] += np.bincount(np.minimum(uw, 7), minlength=8)
dt = time.time() - t0
lines = [f"# Parity of gausslegendre({n}) on the B200 against the C oracle — every half-index ({m} of them)", "",
         f"* mirror symmetry x[i] == -x[n-1-i], w[i] == w[n-1-i]: exact over the whole rule",
         f"* twin oracle (library's own sin/cos compiled for the host): **{twin_bad} mismatching values** out of {2 * m} (bit-for-bit)",
         "* ulp distance |got-ref|/spacing(ref) against the two independent oracles — `libm`: glibc sin/cos/tan; "
         "`julia`: oracle/julia_libm.h, the restated Julia Base (fdlibm-derived, FMA Horner) sin/cos/tan:", "",
         "| ulp | nodes vs libm | weights vs libm | nodes vs julia | weights vs julia |", "|---|---|---|---|---|"]
for u in range(8):
    lines.append(f"| {u}{'+' if u == 7 else ''} | {hx['libm'][u]} | {hw['libm'][u]} | {hx['julia'][u]} | {hw['julia'][u]} |")
mx = {md: (int(np.max(np.nonzero(hx[md])[0])), int(np.max(np.nonzero(hw[md])[0]))) for md in hx}
lines += ["", f"max node / weight ulp: vs libm {mx['libm'][0]} / {mx['libm'][1]}, vs julia {mx['julia'][0]} / {mx['julia'][1]}; "
              f"oracle + compare time {dt:.1f} s on {os.cpu_count()} host threads."]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
