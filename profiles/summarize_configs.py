"""Summarise gpurun_out/cfg_launches2.csv + gpurun_out/cfg2_full_*.csv into profiles/r01_configs_final.{md,csv}."""
import csv
import glob
import os


def load(f):
    return [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]


rows = load('gpurun_out/cfg_launches2.csv')
h = rows[0]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
per = {}
out = ["kernel,duration_ns"]
for r in rows[1:]:
    out.append(f"\"{r[ki]}\",{r[vi]}")
    per.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')))
open('profiles/r01_configs_launches_final.csv', 'w').write("\n".join(out) + "\n")
lines = ["# Round 1, final state — launch list of `python tools/run_configs.py 1` (C1, C3, C4, C5 once each)", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold, serialised: compare shares); full list in "
         "`r01_configs_launches_final.csv`.", "", "| kernel | launches | total us | max us |", "|---|---|---|---|"]
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {max(v) / 1e3:.1f} |")
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for f in sorted(glob.glob('gpurun_out/cfg2_full_*.csv')):
    r = load(f)
    hh = r[0]
    k2 = hh.index('Kernel Name')
    lines += ["", f"## ncu --set full: {os.path.basename(f)[10:-4]}", "",
              "| metric | unit | " + " | ".join(x[k2][:40] for x in r[2:]) + " |", "|---|---|" + "---|" * (len(r) - 2)]
    for k in want:
        if k in hh:
            i = hh.index(k)
            lines.append(f"| {k} | {r[1][i]} | " + " | ".join(x[i] for x in r[2:]) + " |")
open('profiles/r01_configs_final.md', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
