#!/usr/bin/env python
"""Generates tests/golden/losses_ref.npz by running the REFERENCE's own loss-side functions on the CPU of this container:
batch_sample (/root/reference/utils.py:590-633), batch_point_to_point (:393-438), batch_point_to_surface (:441-502),
calc_point_to_line + classes edge / Plane (:506-587), batch_calc_edge (:636-651), batch_get_lap_info (:654-662).

utils.py cannot be imported as a module here (matplotlib, import-time CUDA JIT of the chamfer/tri extensions, hard
`.cuda()` calls), so the definitions are compiled from the reference file where it lies (nothing is copied into this
repository) and executed with
  * `.cuda()` mapped to the identity (CPU run),
  * the module-level singletons `chamfer_dist` / `tri_dist` (utils.py:31-35) bound to the CPU oracle, which the -m gpu
    tests pin bit for bit to the reference's own kernels (tests/test_gpu_chamfer.py / test_gpu_tri.py
    ::test_oracle_matches_reference_cuda),
  * `torch.multinomial` and `Uniform.sample_n` wrapped so that the random draws the reference made are RECORDED: the
    tests feed the same (choices, r1, r2) to the CUDA kernels, which makes every golden a pure function of its inputs.
    python tests/golden/make_golden_losses.py          (needs /root/reference; the .npz is committed)"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/utils.py"
WANTED = ("batch_sample", "batch_point_to_point", "batch_point_to_surface", "calc_point_to_line", "edge", "Plane",
          "batch_calc_edge", "batch_get_lap_info", "calc_adj", "normalize_adj", "adj_init")


def inputs():
    """Seeded inputs shared with the tests: 2 jittered 482-vertex UV spheres (the reference's own 482.obj topology class),
    ground-truth points near the surface, some exactly on vertices / edges (option 1-6 all occur)."""
    from geometrics_b200 import synth

    g = torch.Generator().manual_seed(41)
    v, f = synth.uv_sphere(32, 15)
    verts = synth.mesh_batch(v, 2, 0.01, 7)
    gt = verts[:, torch.randint(0, verts.shape[1], (600,), generator=g)] * (1 + 0.05 * torch.randn(2, 600, 1, generator=g))
    gt[:, :40] = verts[:, :40]                                  # exactly on vertices
    gt[:, 80:240] = 1.25 * verts[:, 100:260]                    # radially beyond a vertex: vertex regions (options 2-3) ...
    gt[:, 240:260] = 1.3 * verts[:, 0:1] + 0.01 * torch.randn(2, 20, 3, generator=g)  # ... and beyond the pole, slot A of the cap faces (option 1)
    gt[:, 40:80] = 0.5 * (verts[:, f[:40, 0]] + verts[:, f[:40, 1]])  # exactly on edge midpoints
    return verts.contiguous(), f, gt.contiguous()


class Recorder:
    def __init__(self):
        self.choices, self.uniform = [], []


def reference_namespace(rec):
    from oracle import Oracle

    orc = Oracle()
    tree = ast.parse(open(REF).read())
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in WANTED]
    assert {n.name for n in wanted} == set(WANTED), {n.name for n in wanted}

    def chamfer_dist(xyz1, xyz2):  # chamfer_distance.py:36-38 returns (idx1, idx2)
        _, i1, _, i2 = orc.chamfer_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
        return torch.from_numpy(i1), torch.from_numpy(i2)

    def tri_dist(xyz, t1, t2, t3):  # tri_distance.py:41-43 returns (dist, point, index)
        d, p, i = orc.tri_fwd(*(t.detach().contiguous().numpy() for t in (xyz, t1, t2, t3)))
        return torch.from_numpy(d), torch.from_numpy(p), torch.from_numpy(i)

    # a torch look-alike that records the random draws
    shim = types.ModuleType("torch_shim")
    shim.__dict__.update({k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})

    def multinomial(a, num, replacement=False):
        out = torch.multinomial(a, num, replacement)
        rec.choices.append(out.clone())
        return out

    class Uniform:
        def __init__(self, lo, hi):
            self.d = torch.distributions.Uniform(lo, hi)

        def sample_n(self, n):
            out = self.d.sample((n,))
            rec.uniform.append(out.clone())
            return out

    shim.multinomial = multinomial
    shim.distributions = types.SimpleNamespace(Uniform=Uniform)
    ns = {"torch": shim, "np": np, "chamfer_dist": chamfer_dist, "tri_dist": tri_dist}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU run of code written with hard .cuda() calls
    rec = Recorder()
    ns = reference_namespace(rec)
    verts, faces, gt = inputs()
    out = {}
    info = ns["adj_init"](faces)        # {'adj', 'adj_orig', 'faces'} (utils.py:96-131)
    num = 500

    # ---- a1: batch_sample ----
    torch.manual_seed(41)
    v = verts.clone().requires_grad_(True)
    pts = ns["batch_sample"](v, faces, num=num)
    w = torch.linspace(0.5, 1.5, pts.numel()).view_as(pts)
    (pts * w).sum().backward()
    out.update(sample_points=pts.detach().numpy(), sample_choices=torch.cat(rec.choices).numpy(),
               sample_r1=rec.uniform[0].numpy(), sample_r2=rec.uniform[1].numpy(), sample_dverts=v.grad.numpy())

    # ---- a6: batch_point_to_point (loss, F1) ----
    for tag, fn in (("p2p", "batch_point_to_point"), ("p2s", "batch_point_to_surface")):
        rec.choices.clear(), rec.uniform.clear()
        v = verts.clone().requires_grad_(True)
        g_in = gt[:, :num] if tag == "p2p" else gt      # batch_point_to_point needs as many gt points as samples (:409-411)
        loss, f1 = ns[fn](v, info, g_in, num=num, f1=True)
        loss.backward()
        out.update({f"{tag}_loss": loss.detach().numpy(), f"{tag}_f1": np.float64(f1), f"{tag}_dverts": v.grad.numpy(),
                    f"{tag}_choices": torch.cat(rec.choices).numpy(), f"{tag}_r1": rec.uniform[0].numpy(),
                    f"{tag}_r2": rec.uniform[1].numpy()})

    # ---- a10: calc_point_to_line on its own, every option, with gradients to the triangle vertices ----
    t1, t2, t3 = (torch.index_select(verts, 1, faces[:, k]) for k in range(3))
    _, opt, idx = ns["tri_dist"](gt, t1, t2, t3)
    flat = (idx.long() + faces.shape[0] * torch.arange(2).unsqueeze(-1)).view(-1)
    tri = [t.reshape(-1, 3)[flat].clone().requires_grad_(True) for t in (t1, t2, t3)]
    d2 = ns["calc_point_to_line"](gt.reshape(-1, 3), tri, opt.view(-1))
    d2.backward()
    out.update(ptl_options=opt.numpy(), ptl_index=idx.numpy(), ptl_dist=d2.detach().numpy(),
               ptl_da=tri[0].grad.numpy(), ptl_db=tri[1].grad.numpy(), ptl_dc=tri[2].grad.numpy())
    assert set(np.unique(opt.numpy()).tolist()) == set(range(7)), np.unique(opt.numpy())

    # ---- f2: regularisers ----
    v = verts.clone().requires_grad_(True)
    e = ns["batch_calc_edge"](v, info)
    e.backward()
    out.update(edge_loss=e.detach().numpy(), edge_dverts=v.grad.numpy())
    v = verts.clone().requires_grad_(True)
    lap = ns["batch_get_lap_info"](v, info)
    (lap ** 2).sum().backward()
    out.update(lap=lap.detach().numpy(), lap_dverts=v.grad.numpy())

    np.savez_compressed(os.path.join(HERE, "losses_ref.npz"), **out)
    print("wrote losses_ref.npz:", {k: tuple(np.shape(a)) for k, a in out.items()})
    print("options histogram:", np.bincount(opt.numpy().ravel(), minlength=7))


if __name__ == "__main__":
    main()
