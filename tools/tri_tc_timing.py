#!/usr/bin/env python
"""Phase timers of the tensor-core TriDistance scan kernel (development build: make -C geometrics_b200/csrc timing)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometrics_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "geometrics_b200", "libgeom_b200_timing.so")
import geometrics_b200 as g
from geometrics_b200 import synth
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    v, f = synth.uv_sphere(50, 50)
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v, 32), f)); p = synth.random_points(32, 8192, 41).to(dev)
elif which == "rand":
    t1, t2, t3 = (t.to(dev) for t in synth.random_triangles(32, 5000)); p = synth.random_points(32, 8192, 41).to(dev)
else:
    v, f = synth.uv_sphere(32, 15)
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v, 16), f)); p = synth.random_points(16, 3000, 3).to(dev)
for _ in range(3):
    g.tri_forward(p, t1, t2, t3, tensor=True)
torch.cuda.synchronize()
L = g.lib()
L.geom_debug_read_tri_timing.argtypes = [C.c_void_p, C.c_size_t]
n = 1024 * 16 * 16
buf = np.zeros(n, np.uint64)
assert L.geom_debug_read_tri_timing(buf.ctypes.data, n) == 0
ncta = min(1024, p.shape[0] * ((p.shape[1] + 511) // 512))
t = buf.reshape(1024, 16, 16)[:ncta].astype(np.float64)
ph = {"sweep1": t[:, :, 1] - t[:, :, 0], "locate+exact": t[:, :, 2] - t[:, :, 1], "sweep2": t[:, :, 3] - t[:, :, 2],
      "wait bar1": t[:, :, 4] - t[:, :, 3], "pool eval": t[:, :, 5] - t[:, :, 4], "wait bar2": t[:, :, 6] - t[:, :, 5],
      "total": t[:, :, 6] - t[:, :, 0]}
print(which, "CTAs", ncta)
for k, v in ph.items():
    print(f"  {k:14s} cycles: mean {v.mean():10.0f}  median {np.median(v):10.0f}  max {v.max():10.0f}   per-CTA max-over-warps mean {v.max(1).mean():10.0f}")
for slot, name in ((8, "eval rounds"), (9, "work items (groups)"), (10, "append steps")):
    c = t[:, :, slot]
    print(f"  {name:18s} per warp: mean {c.mean():8.1f} max {c.max():8.0f}   per CTA sum mean {c.sum(1).mean():9.1f}")
st = ph["pool eval"]; ev = t[:, :, 8]
print("  cycles per eval round (pool eval / rounds), warps with >= 8 rounds:", float((st[ev >= 8] / ev[ev >= 8]).mean()) if (ev >= 8).any() else None)
