#!/usr/bin/env python
"""TriDistance timings (tensor-core filtered search vs CUDA-core brute force vs the reference's own kernel) at BASELINE
config 3 and the training shape, with per-kernel device times.  Prints JSON lines.  BENCH INFRASTRUCTURE (the
reference kernel from oracle/_ref is only a comparison arm)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=20, warm=3):
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


def prof(fn, reps=5):
    _lib.prof_begin()
    for i in range(reps):
        fn(i)
    return {k.split("::", 1)[1]: round(v[0] / v[1], 5) for k, v in _lib.prof_end().items()}


def main():
    try:
        from oracle import RefCuda, have_ref_cuda
        ref = RefCuda() if have_ref_cuda() else None
    except Exception:
        ref = None
    cases = []
    v, faces = synth.uv_sphere(50, 50)
    ring = []
    for i in range(4):
        t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v, 32, seed=i), faces))
        ring.append((synth.random_points(32, 8192, 70 + i).to(dev), t1, t2, t3))
    cases.append(("C3 32x(8192 pts vs 5000 faces) cube points", ring))
    vb = synth.mesh_batch(v, 32)
    gg = torch.Generator().manual_seed(7)
    near = (vb[:, torch.randint(0, vb.shape[1], (8192,), generator=gg)] * (1 + 0.02 * torch.randn(32, 8192, 1, generator=gg))).contiguous().to(dev)
    cases.append(("C3 near-surface points", [(near,) + ring[0][1:]]))
    r1, r2, r3 = (t.to(dev) for t in synth.random_triangles(32, 5000))
    cases.append(("C3 random triangles", [(ring[0][0], r1, r2, r3)]))
    v2, f2 = synth.uv_sphere(32, 15)
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v2, 16), f2))
    cases.append(("train 16x(3000 pts vs 960 faces)", [(synth.random_points(16, 3000, 3).to(dev), t1, t2, t3)]))
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v2, 4), f2))
    cases.append(("small 4x(3000 pts vs 960 faces)", [(synth.random_points(4, 3000, 3).to(dev), t1, t2, t3)]))
    for name, rg in cases:
        rec = {"bench": name}
        for label, kw in (("tensor", {"tensor": True}), ("cuda_cores", {"tensor": False})):
            f = lambda i: g.tri_forward(*rg[i % len(rg)], **kw)  # noqa: E731
            rec[label + "_ms"] = round(timeit(f, reps=10), 5)
            rec[label + "_kernels"] = prof(f)
        if ref:
            rec["ref_ms"] = round(timeit(lambda i: ref.tri_forward_t(*rg[i % len(rg)], sync=False), reps=3), 5)
            a = g.tri_forward(*rg[0], tensor=True)
            b = ref.tri_forward_t(*rg[0])
            rec["identical_to_reference"] = all(bool(torch.equal(x, y)) for x, y in zip(a, b))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
