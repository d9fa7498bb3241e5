#!/usr/bin/env python
"""Run ONE hot kernel a few times at its BASELINE config so ncu can capture it:
    ncu --set full -k regex:<kernel> -s 2 -c 1 ... python tools/prof_one.py chamfer|tri|gcn|gcn482|sample|pool|loss"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import synth  # noqa: E402
import geometrics_b200.utils as U  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "chamfer"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if what == "chamfer":
    a, b = synth.random_points(32, 10000, 41).to(dev), synth.random_points(32, 10000, 1041).to(dev)
    fn = lambda: g.chamfer_forward(a, b, want_dist=False, tensor=False if os.environ.get("GEOM_BENCH_FP32") else None)
elif what == "tri":
    v, f = synth.uv_sphere(50, 50)
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v, 32), f))
    p = synth.random_points(32, 8192, 41).to(dev)
    fn = lambda: g.tri_forward(p, t1, t2, t3)
elif what in ("gcn", "gcn482"):
    torch.backends.cuda.matmul.allow_tf32 = False
    if what == "gcn":
        adj = g.adj_init(synth.icosphere(4)[1])["adj"].to(dev)
        x, w, bias = 0.1 * torch.randn(32, 2562, 256, device=dev), torch.randn(256, 256, device=dev) / 16, torch.randn(256, device=dev)
    else:
        adj = g.adj_init(synth.uv_sphere(32, 15)[1])["adj"].to(dev)
        x, w, bias = 0.1 * torch.randn(16, 482, 963, device=dev), torch.randn(963, 192, device=dev) / 31, torch.randn(192, device=dev)
    fn = lambda: g.zeron_gcn(x, w, bias, adj, 3)
elif what == "sample":
    vi, fi = synth.icosphere(4)
    verts, fi = synth.mesh_batch(vi, 32).to(dev), fi.to(dev)
    fn = lambda: U.batch_sample(verts, fi, 10000)
elif what == "pool":
    vs, _ = synth.uv_sphere(32, 15)
    pv = synth.mesh_batch(vs, 16).to(dev)
    info = torch.stack((torch.rand(16, device=dev) * 360., 25. + 5. * torch.rand(16, device=dev), 0.6 + 0.4 * torch.rand(16, device=dev)), dim=1)
    maps = [torch.randn(16, c, d, d, device=dev) for c, d in ((64, 56), (128, 28), (256, 14), (512, 7))]
    fn = lambda: U.batched_pooling(maps, pv, info)
elif what == "loss":
    vs, fs = synth.uv_sphere(32, 15)
    pv, fs = synth.mesh_batch(vs, 16).to(dev).requires_grad_(True), fs.to(dev)
    gt = (0.45 * torch.nn.functional.normalize(torch.randn(16, 3000, 3), dim=-1)).to(dev)
    fn = lambda: U.batch_point_to_surface(pv, {"faces": fs}, gt, num=3000).backward()
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", what)
