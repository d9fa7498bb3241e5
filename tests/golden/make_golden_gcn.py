"""Generate golden vectors for the 0N-GCN layers / blocks by running the REFERENCE's own layers.py and
models.py (imported from /root/reference, CPU) on seeded inputs with seeded random-init weights.
Run in the build container:

    python tests/golden/make_golden_gcn.py

Writes tests/golden/gcn_ref.npz.  The tests rebuild the same weights (same constructor order, same
torch.manual_seed -> same uniform_ stream) and the same inputs from the seeds; a weight checksum is
stored so that any drift in initialisation shows up as such, not as a numerics failure.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from geometrics_b200 import synth  # noqa: E402
from geometrics_b200.utils import adj_init  # noqa: E402  (plain torch, device-agnostic)

V_SEG, V_RINGS = 32, 15  # the 482-vertex UV sphere family of the reference's 482.obj
B = 2


def inputs():
    verts, faces = synth.uv_sphere(V_SEG, V_RINGS)
    adj = adj_init(faces)["adj"]
    g = torch.Generator().manual_seed(1234)
    pos = verts.unsqueeze(0) + 0.01 * torch.randn(B, verts.shape[0], 3, generator=g)
    pooled = 0.1 * torch.randn(B, verts.shape[0], 960, generator=g)
    x120 = 0.5 * torch.randn(B, verts.shape[0], 120, generator=g)
    return verts, faces, adj, pos, pooled, x120


def checksum(module):
    return float(sum(p.double().abs().sum() for p in module.state_dict().values()))


def build(ns):
    """Construct the modules in a fixed order under fixed seeds; `ns` is the reference's or ours."""
    mods = {}
    torch.manual_seed(101); mods["img"] = ns.Batch_Image_ZERON_GCNGCN(963, 192)
    torch.manual_seed(102); mods["bz"] = ns.BatchZERON_GCN(120, 150)
    torch.manual_seed(103); mods["z"] = ns.ZERON_GCN(120, 150)
    torch.manual_seed(104); mods["max"] = ns.GCNMax(120, 50)
    torch.manual_seed(105); mods["bmax"] = ns.BatchGCNMax(120, 50)
    torch.manual_seed(106); mods["block"] = ns.BatchMeshDeformationBlock(963, 482)
    torch.manual_seed(107); mods["enc"] = ns.BatchMeshEncoder(50)
    torch.manual_seed(108); mods["uenc"] = ns.MeshEncoder(50)
    return mods


def run(mods, adj, pos, pooled, x120):
    out = {}
    full = torch.cat((pos, pooled), dim=-1)
    out["img"] = mods["img"](full, adj, lambda x: x)
    out["bz"] = mods["bz"](x120, adj, F.elu)
    out["z"] = mods["z"](x120[0], adj, F.relu)
    out["max"] = mods["max"](x120[0], adj, F.elu)
    out["bmax"] = mods["bmax"](x120, adj, F.elu)
    mods["block"].train()
    f, c = mods["block"](pos, pooled, adj)
    out["block_train_feat"], out["block_train_coords"] = f, c
    mods["block"].eval()
    f, c = mods["block"](pos, pooled, adj)
    out["block_eval_feat"], out["block_eval_coords"] = f, c
    out["enc"] = mods["enc"](pos, adj)
    out["uenc"] = mods["uenc"](pos[0], adj)
    return out


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    import types

    import layers as ref_layers  # the reference's layers.py
    import models as ref_models  # the reference's models.py (does `from layers import *`)

    ns = types.SimpleNamespace(**{k: getattr(ref_layers, k) for k in dir(ref_layers)},
                               BatchMeshDeformationBlock=ref_models.BatchMeshDeformationBlock,
                               BatchMeshEncoder=ref_models.BatchMeshEncoder, MeshEncoder=ref_models.MeshEncoder)
    verts, faces, adj, pos, pooled, x120 = inputs()
    mods = build(ns)
    wsums = {k: checksum(m) for k, m in mods.items()}  # before the forward passes touch BN statistics
    with torch.no_grad():
        out = run(mods, adj, pos, pooled, x120)
    save = {k: v.numpy() for k, v in out.items()}
    save["block_train_feat"] = save["block_train_feat"][:, ::8]  # keep the fixture small
    save["block_eval_feat"] = save["block_eval_feat"][:, ::8]
    save["img"] = save["img"][:, ::4]
    save["bz"] = save["bz"][:, ::4]
    save["z"] = save["z"][::4]
    for k, v in wsums.items():
        save["wsum_" + k] = np.float64(v)
    # gradients of a scalar through the image layer (backward parity)
    mods["img"].zero_grad()
    x = torch.cat((pos, pooled), dim=-1).requires_grad_(True)
    y = mods["img"](x, adj, lambda t: t)
    (y * torch.linspace(-1, 1, 192)).sum().backward()
    save["img_dx"] = x.grad[:, ::64].numpy()
    save["img_dw"] = mods["img"].weight1.grad[0, ::8].numpy()
    save["img_db"] = mods["img"].bias.grad.numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gcn_ref.npz"), **save)
    print({k: v.shape for k, v in save.items() if hasattr(v, "shape")})
