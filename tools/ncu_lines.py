#!/usr/bin/env python
"""Per-source-line hot spots of an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; ix = None; lines = []; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if len(r) > 6 and r[0] == "Line No": ix = {n: i for i, n in enumerate(r)}; continue
    if ix is None or len(r) < 8 or not r[0].strip().isdigit(): continue
    try: s = int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]] or 0)
    except ValueError: continue
    tot += s; lines.append((s, n, fname, r[0], r[1].strip()[:120]))
print("total samples", tot)
for s, n, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/max(tot,1):5.1f}% {s:7d} inst {n:10d}  {f}:{ln}: {src}")
