#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into text: key metrics, stall reasons, SASS instruction mix.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xyz.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


KEYS = ["Duration", "SM Frequency", "Elapsed Cycles", "Registers Per Thread", "Grid Size", "Block Size", "Theoretical Occupancy",
        "Achieved Occupancy", "Achieved Active Warps Per SM", "Executed Ipc Active", "Issue Slots Busy", "Compute (SM) Throughput",
        "Memory Throughput", "DRAM Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Eligible Warps Per Scheduler",
        "Active Warps Per Scheduler", "Issued Warp Per Scheduler", "No Eligible", "Executed Instructions", "Static Shared Memory Per Block",
        "Dynamic Shared Memory Per Block", "Avg. Active Threads Per Warp", "Avg. Not Predicated Off Threads Per Warp", "Branch Efficiency",
        "Waves Per SM", "Shared Memory Configuration Size"]
rows = list(csv.reader(io.StringIO(ncu("--page", "details", "--csv"))))
hdr = rows[0]
ki, ni, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
print("kernel:", rows[1][ki])
seen = set()
for r in rows[1:]:
    if r[ni] in KEYS and r[ni] not in seen:
        seen.add(r[ni])
        print(f"  {r[ni]:45s} {r[vi]:>16s} {r[ui]}")
raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
rh = raw[0]
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
             "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum",
             "sm__inst_executed_pipe_fmalite.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum"):
    if name in rh:
        print(f"  {name:75s} {raw[2][rh.index(name)]:>18s} {raw[1][rh.index(name)]}")
src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h = src[1]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = collections.Counter()
ops = collections.Counter()
samples = 0
for r in src[2:]:
    for s in stalls:
        try:
            tot[s] += int(r[ix[s]])
        except ValueError:
            pass
    try:
        samples += int(r[ix["# Samples"]])
        n = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    t = r[ix["Source"]].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    ops[op.split(".")[0]] += n
print(f"stall samples (total {samples}):")
for k, v in tot.most_common(9):
    print(f"  {k:28s} {v:8d} {100 * v / max(samples, 1):5.1f}%")
tt = sum(ops.values())
print(f"SASS instruction mix (warp-level executed, total {tt}):")
for k, v in ops.most_common(22):
    print(f"  {k:12s} {v:12d} {100 * v / tt:5.1f}%")
