#!/usr/bin/env python
"""BASELINE config 5: the full GEOMetrics training step (GEOMetrics.py:101-174) on synthetic data --
3 image encoders -> feature pooling -> 3 mesh deformation blocks -> surface + edge + Laplacian
(+ latent) losses -> backward -> Adam -- with the hot path served by geometrics_b200, optionally
sharded over GPUs with DDP (one process per GPU, launched with torchrun).

    python tools/full_step.py [--batch 32] [--steps 10] [--impl ours|torch] [--latent] [--sync-bn]
    torchrun --nproc-per-node 8 tools/full_step.py --batch 32      # global batch 32, 4 per rank

The image encoder is out of scope for the library (stock cuDNN convs), so it is restated here in plain torch
only to make the step runnable; `batched_pooling` is served by the library (csrc/pool.cu) and by the
plain-torch restatement below in the comparison arm.  `--impl torch` swaps
the hot path for the reference's torch expressions (tests/ref_torch.py) and, when
oracle/_ref/libref_cuda.so travelled along, the reference's own .cu kernels: the comparison arm.
HARNESS / BENCH INFRASTRUCTURE, not part of the product path.
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geometrics_b200 as g  # noqa: E402
import geometrics_b200.utils as U  # noqa: E402
from geometrics_b200 import synth  # noqa: E402


class ImageEncoder(nn.Module):
    """Same channel plan / strides / taps as the reference VGG (models.py:12-130): 4-channel input,
    taps after layers 8, 11, 14, 18 -> 64@56^2, 128@28^2, 256@14^2, 512@7^2."""

    PLAN = [(4, 16, 1), (16, 16, 1), (16, 32, 2), (32, 32, 1), (32, 32, 1), (32, 64, 2), (64, 64, 1), (64, 64, 1),
            (64, 128, 2), (128, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 256, 1),
            (256, 512, 2), (512, 512, 1), (512, 512, 1), (512, 512, 1)]
    TAPS = (8, 11, 14, 18)

    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList(nn.Sequential(nn.Conv2d(i, o, 3, padding=1, stride=s), nn.BatchNorm2d(o), nn.ReLU(inplace=True))
                                    for i, o, s in self.PLAN)

    def forward(self, x):
        taps = []
        for k, layer in enumerate(self.layers, 1):
            x = layer(x)
            if k in self.TAPS:
                taps.append(x)
        return taps


def camera(param):
    """utils.py:286-313."""
    theta, phi = math.pi * param[:, 0] / 180., math.pi * param[:, 1] / 180.
    cam_y = param[:, 2] * torch.sin(phi)
    t = param[:, 2] * torch.cos(phi)
    pos = torch.stack((t * torch.cos(theta), cam_y, t * torch.sin(theta)), dim=1)
    az = pos
    ay = torch.tensor([0., 1., 0.], device=param.device).expand_as(az)
    ax = torch.cross(ay, az, dim=1)
    ay = torch.cross(az, ax, dim=1)
    mat = torch.stack([F.normalize(v, dim=1) for v in (ax, ay, az)], dim=1)
    return mat, pos


def batched_pooling(blocks, verts, img_info):
    """utils.py:316-389: project vertices into the image, bilinear gather of the 4 feature maps."""
    mat, pos = camera(img_info)
    pt = torch.matmul(verts * .57 - pos.unsqueeze(1), mat.permute(0, 2, 1))
    x, y, z = pt[..., 0], pt[..., 1], pt[..., 2]
    xs = ((-y) / (-z) * 248 + 112.) / 223.
    ys = (x / (-z) * 248 + 112.) / 223.
    b, v = xs.shape
    outs = []
    for blk in blocks:
        dim = blk.shape[-1]
        cx, cy = torch.clamp(xs * dim, 0, dim - 1), torch.clamp(ys * dim, 0, dim - 1)
        x1, y1, x2, y2 = torch.floor(cx), torch.floor(cy), torch.ceil(cx), torch.ceil(cy)
        wa, wb, wg, wh = x2 - cx, cx - x1, y2 - cy, cy - y1
        flat = blk.reshape(b, blk.shape[1], dim * dim)

        def take(xi, yi):
            idx = (xi.long() * dim + yi.long()).unsqueeze(1).expand(b, blk.shape[1], v)
            return torch.gather(flat, 2, idx)

        f = (wa * wg).unsqueeze(1) * take(x1, y1) + (wh * wa).unsqueeze(1) * take(x1, y2) \
            + (wg * wb).unsqueeze(1) * take(x2, y1) + (wb * wh).unsqueeze(1) * take(x2, y2)
        outs.append(f.permute(0, 2, 1))
    return torch.cat(outs, dim=2)


def reference_ops(dev):
    """The comparison arm: the reference's torch expressions + (if present) its own CUDA kernels."""
    import ref_torch as R
    try:
        from oracle import RefCuda, have_ref_cuda
        ref = RefCuda() if have_ref_cuda() else None
    except Exception:
        ref = None

    def surface(pred_vert, adj_info, gt, num, f1=False):
        faces = adj_info["faces"]
        areas = R.face_areas(pred_vert.detach(), faces)
        ch = torch.cat([torch.multinomial(a, num, True) for a in areas])  # the reference's per-item loop
        n_ = pred_vert.shape[0] * num
        pred = R.batch_sample_given(pred_vert, faces, ch, torch.rand(n_, 1, device=dev), torch.rand(n_, 1, device=dev))
        tris = [torch.index_select(pred_vert.detach(), 1, faces[:, k]).contiguous() for k in range(3)]
        if ref is not None:
            _, ip, _, ig = ref.chamfer_forward_t(gt, pred.detach(), sync=False)
            _, opt, idx = ref.tri_forward_t(gt, *tris, sync=False)
        else:
            _, ip, _, ig = g.chamfer_forward(gt, pred.detach())
            _, opt, idx = g.tri_forward(gt, *tris)
        d1, _ = R.chamfer_terms(gt, pred, ip, ig)
        d2 = R.point_to_surface_term(gt, pred_vert, faces, opt, idx)
        loss = (d1 + d2) * 3000
        return (loss, R.f_score(gt, pred.detach(), ip, ig, num)) if f1 else loss

    def lap(p, adj_info):
        orig = adj_info["adj_orig"]
        ns = (torch.matmul(orig, p) - p) * (1. / (torch.sum(orig, dim=1) - 1)).view(-1, 1)
        return p - ns

    def dense_gcn(x, w, b, a, split, act=None):
        y = R.zeron_gcn(x, w[0] if w.dim() == 3 else w, b, dense_adj[0], split)
        return y if act is None else act(y)

    dense_adj = [None]
    return surface, lap, dense_gcn, dense_adj, ("reference .cu kernels" if ref is not None else "library searches") + " + torch expressions"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="GLOBAL batch (split over ranks)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=3000)
    ap.add_argument("--impl", default="ours", choices=["ours", "torch"])
    ap.add_argument("--latent", action="store_true")
    ap.add_argument("--sync-bn", action="store_true")
    args = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(41)  # identical initial weights on every rank (DDP would broadcast rank 0's anyway)
    from geometrics_b200.sharding import shard_bounds
    lo, hi = shard_bounds(args.batch, rank, world)
    bsz = hi - lo

    verts, faces = synth.uv_sphere(32, 15)  # the 482-vertex sphere family of the reference's 482.obj
    adj_info = {k: v.to(dev) for k, v in U.adj_init(faces).items()}
    init_pos = verts.to(dev)
    V = init_pos.shape[0]

    encoders = [ImageEncoder().to(dev) for _ in range(3)]
    blocks = [g.BatchMeshDeformationBlock(963, V).to(dev), g.BatchMeshDeformationBlock(1155, V).to(dev),
              g.BatchMeshDeformationBlock(1155, V).to(dev)]
    mesh_encoder = g.BatchMeshEncoder(50).to(dev).eval() if args.latent else None
    mods = encoders + blocks
    if args.sync_bn and world > 1:
        mods = [nn.SyncBatchNorm.convert_sync_batchnorm(m) for m in mods]
    if world > 1:
        # bn14 exists in the reference block but is never used in forward (models.py:235): unused parameters
        mods = [nn.parallel.DistributedDataParallel(m, device_ids=[local], find_unused_parameters=True) for m in mods]
    encoders, blocks = mods[:3], mods[3:]
    params = [p for m in mods for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=1e-4)

    surface, lap, note = U.batch_point_to_surface, U.batch_get_lap_info, "geometrics_b200 kernels"
    pooling = U.batched_pooling if args.impl != "torch" else batched_pooling
    orig_gcn = g.layers.zeron_gcn
    if args.impl == "torch":
        surface, lap, dense_gcn, dense_adj, note = reference_ops(dev)
        dense_adj[0] = adj_info["adj"]
        g.layers.zeron_gcn = dense_gcn
        g.BatchMeshDeformationBlock.FUSED_BN = False  # torch BatchNorm1d / relu / add / div, as the reference runs them

    # the GLOBAL batch is generated with one seed and sliced, so that every GPU count trains on the same 32 items
    gen = torch.Generator().manual_seed(100)
    gb = args.batch
    images = torch.rand(gb, 4, 224, 224, generator=gen)[lo:hi].to(dev)
    img_info = torch.stack((torch.rand(gb, generator=gen) * 360, 25 + 5 * torch.rand(gb, generator=gen),
                            0.6 + 0.4 * torch.rand(gb, generator=gen)), dim=1)[lo:hi].to(dev)
    gt = (torch.rand(gb, args.samples, 3, generator=gen) - 0.5)[lo:hi].to(dev)
    latent = torch.randn(gb, 50, generator=gen)[lo:hi].to(dev)
    torch.manual_seed(1000 + rank)  # sampler draws differ per rank, as they would in training

    def step():
        opt.zero_grad(set_to_none=True)
        ipb = init_pos.unsqueeze(0).expand(bsz, V, 3)
        f1_, f2_, f3_ = (e(images) for e in encoders)
        vf = pooling(f1_, ipb.contiguous(), img_info)
        vf, p1 = blocks[0](ipb, vf, adj_info["adj"])
        p1 = ipb + p1
        vf = torch.cat((vf, pooling(f2_, p1, img_info)), dim=-1)
        vf, p2 = blocks[1](p1, vf, adj_info["adj"])
        p2 = p2 + p1
        vf = torch.cat((vf, pooling(f3_, p2, img_info)), dim=-1)
        _, p3 = blocks[2](p2, vf, adj_info["adj"])
        p3 = p3 + p2
        s1 = surface(p1, adj_info, gt, num=args.samples)
        s2 = surface(p2, adj_info, gt, num=args.samples)
        s3, f1 = surface(p3, adj_info, gt, num=args.samples, f1=True)
        surf = s1 * .2 + s2 * .2 + s3 * 2
        edge = sum(U.batch_calc_edge(p, adj_info) for p in (p1, p2, p3)) * 300
        l1 = torch.mean(torch.sum((lap(init_pos, adj_info) - lap(p1, adj_info)) ** 2, 2)) * 1500
        l2 = torch.mean(torch.sum((lap(p1, adj_info) - lap(p2, adj_info)) ** 2, 2)) * 1500 + torch.mean(torch.sum((p1 - p2) ** 2, 2)) * 100
        l3 = torch.mean(torch.sum((lap(p2, adj_info) - lap(p3, adj_info)) ** 2, 2)) * 1500 + torch.mean(torch.sum((p2 - p3) ** 2, 2)) * 100
        lap_loss = .2 * (l1 * .3 + l2 + l3)
        lat = torch.zeros((), device=dev)
        if mesh_encoder is not None:
            lat = .0005 * torch.mean(torch.abs(mesh_encoder(p3, adj_info["adj"]) - latent), dim=1).mean()
        loss = edge + surf + lap_loss + lat
        loss.backward()
        opt.step()
        return loss.detach(), f1

    try:
        first = None
        for _ in range(args.warmup):
            out = step()
            if first is None:
                first = out[0].clone()  # loss of the very first step: same weights and data on every GPU count
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
    finally:
        g.layers.zeron_gcn = orig_gcn
        g.BatchMeshDeformationBlock.FUSED_BN = True
    first = first.double().reshape(1)
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
        dist.all_reduce(first)  # equal shards: the global-batch loss is the mean of the per-rank means
        first /= world
    if rank == 0:
        print(json.dumps({"bench": "full GEOMetrics training step (config 5), synthetic", "impl": args.impl, "hot_path": note,
                          "global_batch": args.batch, "n_gpus": world, "samples": args.samples, "verts": V,
                          "ms_per_step": dt * 1e3, "steps_per_s": 1 / dt, "loss": float(out[0]), "f1": float(out[1]),
                          "first_step_loss_global": float(first),
                          "latent": bool(args.latent), "sync_bn": bool(args.sync_bn)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
