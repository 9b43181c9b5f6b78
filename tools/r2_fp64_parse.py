"""gpurun_out/r2_fp64_ncu.csv (+ r2_fp64_ranges.json) -> profiles/r02_fp64_counts.json: executed FP64 work per node
of every bench configuration, flops = DADD + DMUL + 2 DFMA thread-instructions (predicated-on)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_fp64_ncu.csv")
ranges = json.load(open(os.path.join(ROOT, "gpurun_out", "r2_fp64_ranges.json")))
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
# long format: one row per (launch, metric)
per = {}
order = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    lid = int(r[col["ID"]])
    if lid not in per:
        per[lid] = {"name": r[col["Kernel Name"]]}
        order.append(lid)
    per[lid][r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
# only this library's kernels are counted by fgq_launch_count: keep those (namespace fgq / known names)
ours = [lid for lid in order if "fgq" in per[lid]["name"] or "legendre" in per[lid]["name"] or "jacobi" in per[lid]["name"]
        or "hermite" in per[lid]["name"] or "laguerre" in per[lid]["name"] or "moments" in per[lid]["name"]
        or "codec" in per[lid]["name"] or "besselroots" in per[lid]["name"] or "chebyshev" in per[lid]["name"]
        or "endpoint" in per[lid]["name"] or "tridiag" in per[lid]["name"] or "allgather" in per[lid]["name"]]
out = {}
for key, rg in ranges.items():
    sel = ours[rg["first"]:rg["last"]]
    da = sum(per[i].get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0) for i in sel)
    dm = sum(per[i].get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0) for i in sel)
    df = sum(per[i].get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0) for i in sel)
    t_ns = sum(per[i].get("gpu__time_duration.sum", 0) for i in sel)
    flops = da + dm + 2 * df
    out[key] = {"launches": len(sel), "nodes": rg["nodes"], "dadd": da, "dmul": dm, "dfma": df,
                "fp64_instr_per_node": (da + dm + df) / rg["nodes"], "flops_per_node": flops / rg["nodes"],
                "ncu_time_ms_serialised": t_ns / 1e6,
                "kernels": sorted(set(per[i]["name"].split("(")[0][:60] for i in sel)),
                "source": "ncu smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum over the launches of one "
                          "call (tools/r2_fp64_run.py); flops = dadd + dmul + 2 dfma"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_fp64_counts.json"), "w"), indent=1)
for k, v in out.items():
    print(f"{k:36s} launches {v['launches']:3d}  instr/node {v['fp64_instr_per_node']:9.1f}  flops/node {v['flops_per_node']:9.1f}  ncu {v['ncu_time_ms_serialised']:.3f} ms")
