#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into markdown: headline metrics, fp64 flop count, opcode mix.
usage: ncu_summary.py report.ncu-rep units_in_launch "unit name" > profiles/xxx.md"""
import collections
import csv
import io
import subprocess
import sys

rep, units, unit_name = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit, vals)}


def f(name):
    return float(m[name][0].replace(",", "")) if name in m and m[name][0] not in ("", "n/a") else float("nan")


print(f"# ncu summary: `{m['Kernel Name'][0]}`\n")
print(f"source report: `{rep}` (ncu --set full --clock-control none); {units:.0f} {unit_name} in this launch\n")
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in m:
        print(f"| {k} | {m[k][0]} | {m[k][1]} |")
da, dm, df = (f(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum") for o in ("dadd", "dmul", "dfma"))
flop = da + dm + 2 * df
dur = f("gpu__time_duration.sum")
dur_s = dur * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(m["gpu__time_duration.sum"][1], 1e-3)
print(f"\n* executed fp64 flop (DADD + DMUL + 2 DFMA, thread level) = {flop:.4g} -> **{flop / units:.1f} flop per {unit_name}**")
print(f"* under ncu: {units / dur_s:.4g} {unit_name}/s, {flop / dur_s / 1e12:.2f} TFLOP/s fp64 executed")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
ip = h.index("Predicated-On Thread Instructions Executed") if "Predicated-On Thread Instructions Executed" in h else None
ops, st, thr, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    n = int(r[ie] or 0)
    ops[op] += n; st[op] += int(r[iss] or 0); tot += n
    if ip is not None:
        thr[op] += int(r[ip] or 0)
if flop != flop and ip is not None:        # the per-opcode counters were not in this capture: same count from the SASS page
    flop = thr["DADD"] + thr["DMUL"] + 2 * thr["DFMA"]
    print(f"* executed fp64 flop from the source page (predicated-on thread instructions: DADD {thr['DADD']:.4g} + DMUL {thr['DMUL']:.4g} "
          f"+ 2 x DFMA {thr['DFMA']:.4g}) = {flop:.4g} -> **{flop / units:.1f} flop per {unit_name}**, {flop / dur_s / 1e12:.2f} TFLOP/s under ncu")
print(f"\n## SASS opcode mix (warp-level instructions executed; {tot / (units / 32):.0f} per warp per {unit_name})\n")
print("| opcode | executed | share | stall samples |\n|---|---|---|---|")
for op, n in ops.most_common(16):
    print(f"| {op} | {n} | {100 * n / tot:.1f}% | {st[op]} |")
