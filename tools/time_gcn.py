#!/usr/bin/env python
"""0N-GCN layer at C4 (32 x 2562 x 256 -> 256) and the training shapes: per-kernel times (dev tool)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g
from geometrics_b200 import _lib, synth
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
for name, (v, f), b, fin, fout in (("C4", synth.icosphere(4), 32, 256, 256), ("train 963->192", synth.uv_sphere(32, 15), 16, 963, 192)):
    adj = g.adj_init(f)["adj"].to(dev)
    x, w, bias = 0.1 * torch.randn(b, v.shape[0], fin, device=dev), torch.randn(fin, fout, device=dev) / fin ** 0.5, torch.randn(fout, device=dev)
    for _ in range(3): g.zeron_gcn(x, w, bias, adj, 3)
    _lib.prof_begin()
    for _ in range(20): g.zeron_gcn(x, w, bias, adj, 3)
    k = _lib.prof_end()
    print(json.dumps({"case": name, **{kk.split("::", 1)[-1]: round(vv[0] / 20, 4) for kk, vv in k.items() if vv[0] > 0}}), flush=True)
