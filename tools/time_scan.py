#!/usr/bin/env python
"""Chamfer tensor-core scan kernel time at C2 (and a small batch) + exactness against the CUDA-core path (dev tool)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g
from geometrics_b200 import _lib, synth
dev = torch.device("cuda:0")
for b in (32, 4):
    sets = [(synth.random_points(b, 10000, 41 + i).to(dev), synth.random_points(b, 10000, 141 + i).to(dev)) for i in range(max(2, 96 // b))]
    f = lambda i: g.chamfer_forward(*sets[i % len(sets)], want_dist=False)
    for i in range(3): f(i)
    _lib.prof_begin()
    for i in range(20): f(i)
    k = _lib.prof_end()
    want = g.chamfer_forward(*sets[0], tensor=False)
    got = g.chamfer_forward(*sets[0])
    ok = all(bool(torch.equal(a, c)) for a, c in zip(got, want))
    print(json.dumps({"b": b, "scan_ms": round(k["geom::chamfer_tc::scan_kernel"][0] / 20, 4), "pack_ms": round(k["geom::chamfer_tc::pack_kernel"][0] / 20, 4), "exact": ok}), flush=True)
