#!/usr/bin/env python
"""Time the Chamfer NN search at BASELINE config 2 (32 x 10000 vs 10000) for the FP32 brute-force path and the
tensor-core filtered path, and check that both return identical indices.  CUDA events, ring of inputs > L2.
    python tools/time_chamfer.py [B N M reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda:0")
B, N, M, reps = (int(x) for x in (sys.argv[1:5] + ["32", "10000", "10000", "30"][len(sys.argv) - 1:]))
ring = 12
xs = [(synth.random_points(B, N, 41 + i).to(dev), synth.random_points(B, M, 1041 + i).to(dev)) for i in range(ring)]


def timeit(fn, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def prof(fn):
    _lib.prof_begin()
    for i in range(8):
        fn(i)
    return {k.split("::", 1)[1]: round(v[0] / v[1], 5) for k, v in _lib.prof_end().items()}


out = {"B": B, "N": N, "M": M}
for name, kw in (("fp32", {"tensor": False}), ("tensor", {"tensor": True})):
    f = lambda i: g.chamfer_forward(*xs[i % ring], want_dist=False, **kw)
    ms = timeit(f)
    out[name] = {"ms": round(ms, 4), "pairs_per_s": B * N * M / (ms * 1e-3), "kernels_ms": prof(f)}
same = True
for i in range(3):
    a = g.chamfer_forward(*xs[i], want_dist=True, tensor=False)
    b = g.chamfer_forward(*xs[i], want_dist=True, tensor=True)
    same &= all(torch.equal(p.view(torch.int32), q.view(torch.int32)) for p, q in zip(a, b))
out["identical"] = bool(same)
print(json.dumps(out), flush=True)
