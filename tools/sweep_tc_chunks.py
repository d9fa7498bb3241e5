#!/usr/bin/env python
"""Chamfer tensor-core search: kernel time vs candidate-split factor for the per-GPU batches of the strong-scaling run
(dev tool; sets the library's GEOM_DEBUG_TC_CHUNKS knob)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g
from geometrics_b200 import _lib, synth
dev = torch.device("cuda:0")
L = g.lib()
for b in (4, 8, 16, 32):
    sets = [(synth.random_points(b, 10000, 41 + i).to(dev), synth.random_points(b, 10000, 141 + i).to(dev)) for i in range(max(2, 64 // b))]
    row = {"b": b}
    for c in (0, 1, 2, 3, 4, 6, 8):
        L.geom_debug_set(_lib.GEOM_DEBUG_TC_CHUNKS, float(c))
        f = lambda i: g.chamfer_forward(*sets[i % len(sets)], want_dist=False)
        for i in range(3): f(i)
        _lib.prof_begin()
        for i in range(10): f(i)
        k = _lib.prof_end()
        row["auto" if c == 0 else f"c{c}"] = round(k["geom::chamfer_tc::scan_kernel"][0] / 10, 4)
    L.geom_debug_set(_lib.GEOM_DEBUG_TC_CHUNKS, 0.0)
    print(json.dumps(row), flush=True)
