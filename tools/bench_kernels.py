#!/usr/bin/env python
"""Per-kernel timings at the BASELINE configs (C2 Chamfer, C3 TriDistance, C4 0N-GCN, sampler, losses)
for this library and, beside each, the reference implementation of the same step on the same GPU:
the reference's own .cu kernels (oracle/_ref/libref_cuda.so) for the searches, and the reference's
torch expressions (tests/ref_torch.py) for the sampler / losses / dense-adjacency GCN.
CUDA events, 3 warm-ups, inputs cycled through a ring larger than L2.  Prints JSON lines.
TEST/BENCH INFRASTRUCTURE (uses oracle/ and tests/ only as comparison arms)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import _lib, synth  # noqa: E402
import geometrics_b200.utils as U  # noqa: E402
import ref_torch as R  # noqa: E402

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


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


def prof(fn, reps=10):
    _lib.prof_begin()
    for i in range(reps):
        fn(i)
    return {k.split("::", 1)[1]: round(v[0] / v[1], 5) for k, v in _lib.prof_end().items()}


def emit(name, ms, ref_ms=None, alg_bytes=None, units=None, unit_name=None, kernels=None, note=None):
    rec = {"bench": name, "ms": round(ms, 5)}
    if ref_ms is not None:
        rec.update(ref_ms=round(ref_ms, 5), speedup_vs_ref=round(ref_ms / ms, 2))
    if alg_bytes is not None:
        rec.update(algorithmic_MB=round(alg_bytes / 1e6, 3), hbm_gbs=round(alg_bytes / ms / 1e6, 2),
                   hbm_frac_of_measured=round(alg_bytes / ms / 1e6 / HBM, 5))
    if units is not None:
        rec[unit_name + "_per_s"] = units / (ms * 1e-3)
    if kernels:
        rec["kernels_ms"] = kernels
    if note:
        rec["note"] = note
    print(json.dumps(rec), flush=True)


def main():
    try:
        from oracle import RefCuda, have_ref_cuda
        ref = RefCuda() if have_ref_cuda() else None
    except Exception:
        ref = None

    # ---------------- C2: Chamfer 32 x (10000 vs 10000)
    B, N, M = 32, 10000, 10000
    ring = [(synth.random_points(B, N, 41 + i).to(dev), synth.random_points(B, M, 1041 + i).to(dev)) for i in range(16)]
    f = lambda i: g.chamfer_forward(*ring[i % 16], want_dist=False)
    ms = timeit(f)
    rms = timeit(lambda i: ref.chamfer_forward_t(*ring[i % 16], sync=False), reps=5) if ref else None
    emit("C2 chamfer fwd 32x(10000 vs 10000)", ms, rms, 16 * B * (N + M), B * N * M, "point_pairs", prof(f))

    # ---------------- C1-sized and training-sized Chamfer
    for (b, n, m) in ((1, 1024, 1024), (16, 3000, 3000)):
        rr = [(synth.random_points(b, n, 5 + i).to(dev), synth.random_points(b, m, 55 + i).to(dev)) for i in range(4)]
        f = lambda i: g.chamfer_forward(*rr[i % 4], want_dist=False)
        ms = timeit(f, reps=50)
        rms = timeit(lambda i: ref.chamfer_forward_t(*rr[i % 4], sync=False), reps=20) if ref else None
        emit(f"chamfer fwd {b}x({n} vs {m})", ms, rms, 16 * b * (n + m), b * n * m, "point_pairs", prof(f))

    # ---------------- C3: TriDistance 32 x (8192 points vs 5000 faces)
    v, faces = synth.uv_sphere(50, 50)
    tr = []
    for i in range(8):
        t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v, 32, seed=i), faces))
        tr.append((synth.random_points(32, 8192, 70 + i).to(dev), t1, t2, t3))
    f = lambda i: g.tri_forward(*tr[i % 8])
    ms = timeit(f, reps=10)
    rms = timeit(lambda i: ref.tri_forward_t(*tr[i % 8], sync=False), reps=3) if ref else None
    emit("C3 tri fwd 32x(8192 pts vs 5000 faces)", ms, rms, 32 * 8192 * 24 + 3 * 32 * 5000 * 12, 32 * 8192 * 5000, "point_triangle_evals", prof(f, 5))
    # training-sized: 16 x (3000 vs 960)
    v2, f2 = synth.uv_sphere(32, 15)
    t1, t2, t3 = (t.to(dev) for t in synth.face_vertices(synth.mesh_batch(v2, 16), f2))
    p2 = synth.random_points(16, 3000, 3).to(dev)
    f = lambda i: g.tri_forward(p2, t1, t2, t3)
    ms = timeit(f, reps=50)
    rms = timeit(lambda i: ref.tri_forward_t(p2, t1, t2, t3, sync=False), reps=20) if ref else None
    emit("tri fwd 16x(3000 pts vs 960 faces)", ms, rms, 16 * 3000 * 24 + 3 * 16 * 960 * 12, 16 * 3000 * 960, "point_triangle_evals")

    # ---------------- sampler: 32 meshes (V=2562 icosphere) x 10000 samples
    vi, fi = synth.icosphere(4)
    verts = synth.mesh_batch(vi, 32).to(dev)
    fi = fi.to(dev)
    num = 10000
    gen = torch.Generator(device=dev).manual_seed(1)
    choices = torch.randint(0, fi.shape[0], (32 * num,), device=dev, generator=gen)
    r1 = torch.rand(32 * num, 1, device=dev, generator=gen)
    r2 = torch.rand(32 * num, 1, device=dev, generator=gen)
    f = lambda i: U.batch_sample(verts, fi, num, choices=choices, r1=r1, r2=r2)
    ms = timeit(f, reps=50)
    rms = timeit(lambda i: R.batch_sample_given(verts, fi, choices, r1, r2), reps=20)
    alg = 32 * 2562 * 12 + fi.shape[0] * 24 + 32 * num * 16 + 32 * num * 12
    emit("sampler gather fwd 32x10000 (given draws)", ms, rms, alg, 32 * num, "samples", prof(f))
    f = lambda i: U.batch_sample(verts, fi, num)
    ms = timeit(f, reps=20)

    def ref_full(i):
        a = R.face_areas(verts, fi)
        ch = torch.cat([torch.multinomial(x, num, True) for x in a])
        return R.batch_sample_given(verts, fi, ch, torch.rand(32 * num, 1, device=dev), torch.rand(32 * num, 1, device=dev))
    emit("batch_sample end to end 32x10000 (areas+multinomial+gather)", ms, timeit(ref_full, reps=5), None, 32 * num, "samples", prof(f))

    # ---------------- vertex feature pooling: B=16, V=482, VGG-sized maps (64/128/256/512 ch at 56/28/14/7)
    vsp, _ = synth.uv_sphere(32, 15)
    pverts = synth.mesh_batch(vsp, 16).to(dev)
    gen = torch.Generator(device=dev).manual_seed(2)
    info = torch.stack((torch.rand(16, device=dev, generator=gen) * 360., 25. + 5. * torch.rand(16, device=dev, generator=gen),
                        0.6 + 0.4 * torch.rand(16, device=dev, generator=gen)), dim=1)
    maps = [torch.randn(16, c, d, d, device=dev, generator=gen) for c, d in ((64, 56), (128, 28), (256, 14), (512, 7))]
    f = lambda i: U.batched_pooling(maps, pverts, info)
    ms = timeit(f, reps=50)
    rms = timeit(lambda i: R.batched_pooling(maps, pverts, info), reps=20)
    alg = 16 * 482 * 960 * 4 * 5 + 16 * 482 * 12  # 4 texel reads + 1 write per feature, positions
    emit("batched_pooling fwd B=16 V=482 C=960 (projection + gather)", ms, rms, alg, 16 * 482 * 960, "features", prof(f))
    mg = [m.clone().requires_grad_(True) for m in maps]
    pg = pverts.clone().requires_grad_(True)

    def pool_step(fn):
        def run(i):
            pg.grad = None
            for m in mg:
                m.grad = None
            fn(mg, pg, info).sum().backward()
        return run
    ms = timeit(pool_step(U.batched_pooling), reps=30)
    rms = timeit(pool_step(R.batched_pooling), reps=10)
    emit("batched_pooling fwd+bwd B=16 V=482 C=960", ms, rms, None, 16 * 482 * 960, "features", prof(pool_step(U.batched_pooling)))

    # ---------------- losses at the training size (B=16, 3000 samples, 482-vertex mesh)
    vs, fs = synth.uv_sphere(32, 15)
    pv = synth.mesh_batch(vs, 16).to(dev).requires_grad_(True)
    fs = fs.to(dev)
    gt = (0.45 * torch.nn.functional.normalize(torch.randn(16, 3000, 3), dim=-1)).to(dev)
    adj_info = {"faces": fs}

    def step(i):
        pv.grad = None
        U.batch_point_to_surface(pv, adj_info, gt, num=3000).backward()

    def ref_step(i):
        pv.grad = None
        a = R.face_areas(pv.detach(), fs)
        ch = torch.cat([torch.multinomial(x, 3000, True) for x in a])
        pred = R.batch_sample_given(pv, fs, ch, torch.rand(48000, 1, device=dev), torch.rand(48000, 1, device=dev))
        if ref:
            _, ip, _, ig = ref.chamfer_forward_t(gt, pred.detach(), sync=False)
            tt = [torch.index_select(pv.detach(), 1, fs[:, k]).contiguous() for k in range(3)]
            _, opt, idx = ref.tri_forward_t(gt, *tt, sync=False)
        else:
            _, ip, _, ig = g.chamfer_forward(gt, pred.detach())
            _, opt, idx = g.tri_forward(gt, *[torch.index_select(pv.detach(), 1, fs[:, k]) for k in range(3)])
        d1, _ = R.chamfer_terms(gt, pred, ip, ig)
        d2 = R.point_to_surface_term(gt, pv, fs, opt, idx)
        ((d1 + d2) * 3000).backward()

    ms = timeit(step, reps=20)
    emit("batch_point_to_surface fwd+bwd B=16 num=3000 V=482", ms, timeit(ref_step, reps=5), None, 1, "steps", prof(step),
         note="reference arm = reference .cu searches + the reference's torch loss expressions")

    # ---------------- C4: 0N-GCN layer, V=2562, 256 -> 256, B=32
    adj = g.adj_init(synth.icosphere(4)[1])["adj"].to(dev)
    xs = [0.1 * torch.randn(32, 2562, 256, device=dev) for _ in range(4)]
    w = (torch.randn(256, 256, device=dev) / 16)
    bias = torch.randn(256, device=dev)
    csr = g.csr_from_dense(adj)
    torch.backends.cuda.matmul.allow_tf32 = False
    f = lambda i: g.zeron_gcn(xs[i % 4], w, bias, csr, 3)
    ms = timeit(f, reps=20)
    rms = timeit(lambda i: R.zeron_gcn(xs[i % 4], w, bias, adj, 3), reps=10)
    alg = 2 * 32 * 2562 * 256 * 4 + 256 * 256 * 4 + csr.nnz * 8 + 2563 * 4
    flops = 2 * 32 * 2562 * 256 * 256 + 2 * 32 * csr.nnz * 85
    emit("C4 0N-GCN layer fwd B=32 V=2562 256->256 (k=85)", ms, rms, alg, flops, "flop", prof(f),
         note="reference arm = torch dense-adjacency expression (cuBLAS fp32, TF32 off)")
    f = lambda i: g.zeron_gcn(xs[i % 4], w, bias, csr, 3, tensor=True)
    ms = timeit(f, reps=20)
    emit("C4 0N-GCN layer fwd, tensor-core GEMM (tcgen05 3xTF32)", ms, rms, alg, flops, "flop", prof(f),
         note="same layer with GEOM_GCN_TENSOR")
    # the reference block's real layer shapes at V=482, B=16
    adj2 = g.adj_init(synth.uv_sphere(32, 15)[1])["adj"].to(dev)
    for fin, fout in ((963, 192), (192, 192), (192, 3)):
        x = 0.1 * torch.randn(16, 482, fin, device=dev)
        ww = torch.randn(fin, fout, device=dev) / fin ** 0.5
        bb = torch.randn(fout, device=dev)
        f = lambda i: g.zeron_gcn(x, ww, bb, adj2, 3)
        ms = timeit(f, reps=50)
        rms = timeit(lambda i: R.zeron_gcn(x, ww, bb, adj2, 3), reps=50)
        emit(f"0N-GCN layer fwd B=16 V=482 {fin}->{fout}", ms, rms, None, 2 * 16 * 482 * fin * fout, "flop", prof(f))
        if fout >= 16:
            f = lambda i: g.zeron_gcn(x, ww, bb, adj2, 3, tensor=True)
            emit(f"0N-GCN layer fwd B=16 V=482 {fin}->{fout}, tensor-core GEMM", timeit(f, reps=50), rms, None, 2 * 16 * 482 * fin * fout, "flop", prof(f))
    # whole deformation block fwd+bwd (B=16, V=482, 963 features)
    block = g.BatchMeshDeformationBlock(963, 482).to(dev)
    pos = synth.mesh_batch(synth.uv_sphere(32, 15)[0], 16).to(dev)
    pooled = 0.1 * torch.randn(16, 482, 960, device=dev)

    def blk(i):
        block.zero_grad(set_to_none=True)
        f_, c_ = block(pos, pooled, adj2)
        (c_.sum() + f_.mean()).backward()

    ms = timeit(blk, reps=10)
    orig = g.layers.zeron_gcn
    g.layers.zeron_gcn = lambda x, w_, b_, a, split, act=None: (lambda y: y if act is None else act(y))(
        R.zeron_gcn(x, w_[0] if w_.dim() == 3 else w_, b_, adj2, split))
    try:
        rms = timeit(blk, reps=10)
    finally:
        g.layers.zeron_gcn = orig
    emit("BatchMeshDeformationBlock fwd+bwd B=16 V=482", ms, rms, None, 1, "steps", prof(blk, 3))


if __name__ == "__main__":
    main()
