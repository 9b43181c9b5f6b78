"""Turn gpurun_out/launches.csv + a .ncu-rep into the markdown/CSV summaries kept under profiles/.
usage: python profiles/summarize_ncu.py <launches.csv> <report.ncu-rep> <out_prefix>"""
import csv
import io
import subprocess
import sys

launches, rep, out = sys.argv[1:4]
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("=="))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
per = {}
lines = ["kernel,duration_" + rows[1][ui]]
for r in rows[1:]:
    lines.append(f"\"{r[ki]}\",{r[vi]}")
    per.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
open(out + "_launches.csv", "w").write("\n".join(lines) + "\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hh = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
md = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rr) - 2)) + " |", "|---|---|" + "---|" * (len(rr) - 2)]
for k in want:
    if k in hh:
        i = hh.index(k)
        md.append(f"| {k} | {rr[1][i]} | " + " | ".join(r[i] for r in rr[2:]) + " |")
tot = sum(sum(v) for v in per.values())
md.append("")
md.append("| kernel | launches | total ns | share |")
md.append("|---|---|---|---|")
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    md.append(f"| {k} | {len(v)} | {sum(v):.0f} | {sum(v) / tot:.4f} |")
open(out + "_ncu.md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
