#!/usr/bin/env python
"""Per-kernel table from a multi-kernel .ncu-rep (SpeedOfLight / Memory / Occupancy sections):
    python tools/ncu_table.py rep.ncu-rep > profiles/xyz.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
want = ["Duration", "Memory Throughput", "DRAM Throughput", "Compute (SM) Throughput", "L1/TEX Hit Rate", "L2 Hit Rate",
        "Achieved Occupancy", "Registers Per Thread", "Grid Size", "Block Size", "Mem Busy", "Max Bandwidth"]
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0])
    d = per.setdefault(key, {})
    if r[ix["Metric Name"]] in want and r[ix["Metric Name"]] not in d:
        d[r[ix["Metric Name"]]] = (r[ix["Metric Value"]], r[ix["Metric Unit"]])
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
rh = raw[0]; rix = {n: i for i, n in enumerate(rh)}
dram = {}
for r in raw[2:]:
    try:
        rd, wr = float(r[rix["dram__bytes_read.sum"]]), float(r[rix["dram__bytes_write.sum"]])
        dram[r[rix["ID"]]] = (rd, rh and raw[1][rix["dram__bytes_read.sum"]], wr, raw[1][rix["dram__bytes_write.sum"]])
    except (KeyError, ValueError):
        pass
print(f"{'kernel':34s} {'grid':>7s} {'blk':>4s} {'regs':>4s} {'time us':>8s} {'mem %':>6s} {'dram %':>6s} {'sm %':>6s} {'L1 hit':>6s} {'L2 hit':>6s} {'occ %':>6s}  dram read/write")
for (kid, name), d in per.items():
    g = lambda k: d.get(k, ("", ""))[0]
    dur, unit = d.get("Duration", ("0", "us"))
    us = float(dur.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    dr = dram.get(kid)
    drs = f"{dr[0]:.2f} {dr[1]} / {dr[2]:.2f} {dr[3]}" if dr else ""
    print(f"{name[:34]:34s} {g('Grid Size'):>7s} {g('Block Size'):>4s} {g('Registers Per Thread'):>4s} {us:8.2f} {g('Memory Throughput'):>6s} {g('DRAM Throughput'):>6s} "
          f"{g('Compute (SM) Throughput'):>6s} {g('L1/TEX Hit Rate'):>6s} {g('L2 Hit Rate'):>6s} {g('Achieved Occupancy'):>6s}  {drs}")
