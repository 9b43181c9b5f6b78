"""Summarise the end-of-round captures (tools/final_gpu.sh) into profiles/r01_final_launches.md and
profiles/r01_final_ncu.md:
  gpurun_out/cfg_launches_final2.csv, gpurun_out/small_batch_launches.csv  (gpu__time_duration per launch)
  gpurun_out/final_{rec_tile,jacobi_pass,small_batch,laguerre}.csv        (ncu --set full, --page raw --csv)"""
import csv
import os


def load(f):
    return [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]


def launch_table(f, title, cmd):
    rows = load(f)
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    per = {}
    order = []
    for r in rows[1:]:
        k = r[ki].split('(')[0]
        if k not in per:
            order.append(k)
        per.setdefault(k, []).append(float(r[vi].replace(',', '')))
    lines = [f"## {title}", "", f"`ncu --metrics gpu__time_duration.sum --clock-control none` of `{cmd}` (cold, serialised: compare "
             "shares, not absolutes).", "", "| kernel | launches | total us | max us | first launches, us |", "|---|---|---|---|---|"]
    for k in order:
        v = per[k]
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {max(v) / 1e3:.1f} | {', '.join(f'{x / 1e3:.1f}' for x in v[:6])} |")
    lines.append(f"| total | {sum(len(v) for v in per.values())} | {sum(sum(v) for v in per.values()) / 1e3:.1f} | | |")
    return lines


WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def full_table(f, title):
    r = load(f)
    hh = r[0]
    k2 = hh.index('Kernel Name')
    cols = r[2:]
    lines = ["", f"## {title}", "", "| metric | unit | " + " | ".join(f"`{x[k2].split('(')[0][-34:]}`" for x in cols) + " |",
             "|---|---|" + "---|" * len(cols)]
    for k in WANT:
        if k in hh:
            i = hh.index(k)
            lines.append(f"| {k} | {r[1][i]} | " + " | ".join(x[i] for x in cols) + " |")
    return lines


out = ["# Round 1, end of round — launch lists (after the staged Jacobi norm pass, the staged Laguerre hand-over, the",
       "table-driven small-rule recurrences and the small-rule batches)", ""]
out += launch_table('gpurun_out/cfg_launches_final2.csv', "C1, C3, C4, C5 once each", "python tools/run_configs.py 1")
out += [""] + launch_table('gpurun_out/small_batch_launches.csv', "1e5-rule batches of the other families + truncated gausshermite(1e6)",
                           "python tools/run_small_batches.py")
open('profiles/r01_final_launches.md', 'w').write("\n".join(out) + "\n")
out2 = ["# Round 1, end of round — `ncu --set full --clock-control none` of the kernels changed late in the round"]
for f, t in (("final_rec_tile", "legendre_rec_tile_kernel (C5: 1e6 rules)"),
             ("final_jacobi_pass", "jacobi_pass_kernel, first six launches of C3 (norm stage 1, norm stage 2, evaluation; right half then left half)"),
             ("final_small_batch", "small-rule batch kernels (1e5 rules each)"),
             ("final_laguerre", "laguerre kernels (C4: n = 1e6)")):
    p = f"gpurun_out/{f}.csv"
    if os.path.exists(p):
        out2 += full_table(p, t)
open('profiles/r01_final_ncu.md', 'w').write("\n".join(out2) + "\n")
print("\n".join(out + out2))
