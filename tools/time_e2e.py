#!/usr/bin/env python
"""Time geom_chamfer_fwd_host (pinned host buffers in, indices out) at BASELINE config 2 for several chunk counts.
    python tools/time_e2e.py [chunks ...]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
B, N, M = 32, 10000, 10000
hx1, hx2 = synth.random_points(B, N, 41).pin_memory(), synth.random_points(B, M, 1041).pin_memory()
hi1 = torch.empty(B, N, dtype=torch.int32).pin_memory()
hi2 = torch.empty(B, M, dtype=torch.int32).pin_memory()
out = {}
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5, 6, 8]:
    for tensor in (None, False):
        for _ in range(3):
            g.chamfer_forward_host(hx1, hx2, hi1, hi2, device=dev, chunks=chunks, tensor=tensor)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(30):
            g.chamfer_forward_host(hx1, hx2, hi1, hi2, device=dev, chunks=chunks, tensor=tensor)
        out[f"chunks={chunks} {'auto' if tensor is None else 'fp32'}"] = round((time.perf_counter() - t0) / 30 * 1e3, 4)
# plain copies for scale
d1, d2 = torch.empty(B, N, 3, device=dev), torch.empty(B, M, 3, device=dev)
di = torch.empty(B, N, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    d1.copy_(hx1, non_blocking=True); d2.copy_(hx2, non_blocking=True)
    torch.cuda.synchronize()
out["h2d_only_ms"] = round((time.perf_counter() - t0) / 30 * 1e3, 4)
t0 = time.perf_counter()
for _ in range(30):
    hi1.copy_(di, non_blocking=True); hi2.copy_(di, non_blocking=True)
    torch.cuda.synchronize()
out["d2h_only_ms"] = round((time.perf_counter() - t0) / 30 * 1e3, 4)
print(json.dumps(out, indent=1))
