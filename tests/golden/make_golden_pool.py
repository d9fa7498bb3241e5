#!/usr/bin/env python
"""Generates tests/golden/pool_ref.npz by running the REFERENCE's own batch_camera_info / batched_pooling
(/root/reference/utils.py:283-389) on the CPU of this container.

utils.py cannot be imported as a module here (matplotlib, import-time CUDA JIT of the chamfer/tri extensions), so
the two function definitions are compiled from the reference file where it lies (nothing is copied into this
repository) and executed with `.cuda()` / `torch.cuda.LongTensor` mapped to their CPU equivalents.
    python tests/golden/make_golden_pool.py          (needs /root/reference; the .npz is committed)"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils.py"


def inputs():
    """Seeded inputs shared with the tests: 2 meshes x 96 vertices, 4 small feature maps."""
    g = torch.Generator().manual_seed(41)
    b, v = 2, 96
    verts = torch.rand(b, v, 3, generator=g) - 0.5
    verts[:, :6] *= 3.0                       # some vertices project outside the image: clamped coordinates
    img_info = torch.stack((torch.rand(b, generator=g) * 360., 25. + 5. * torch.rand(b, generator=g),
                            0.6 + 0.4 * torch.rand(b, generator=g)), dim=1)
    blocks = [torch.randn(b, c, d, d, generator=g) for c, d in ((5, 56), (7, 28), (33, 14), (40, 7))]
    return verts, img_info, blocks


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("batch_camera_info", "batched_pooling")]
    assert len(wanted) == 2
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns["batch_camera_info"], ns["batched_pooling"]


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU run of code written with hard .cuda() calls
    torch.cuda.LongTensor = torch.LongTensor
    cam, pooling = reference_functions()
    verts, img_info, blocks = inputs()
    verts = verts.clone().requires_grad_(True)
    blocks = [blk.clone().requires_grad_(True) for blk in blocks]
    cam_mat, cam_pos = cam(img_info)
    feats = pooling(blocks, verts, img_info)
    up = torch.linspace(-1, 1, feats.shape[-1]).view(1, 1, -1) * torch.linspace(0.5, 1.5, feats.shape[1]).view(1, -1, 1)
    (feats * up).sum().backward()
    np.savez_compressed(os.path.join(HERE, "pool_ref.npz"), cam_mat=cam_mat.numpy(), cam_pos=cam_pos.numpy(),
                        feats=feats.detach().numpy(), d_verts=verts.grad.numpy(),
                        **{f"d_block{i}": blk.grad.numpy() for i, blk in enumerate(blocks)})
    print("wrote pool_ref.npz", tuple(feats.shape), "zero rows:", int((feats.detach().abs().sum(-1) == 0).sum()))


if __name__ == "__main__":
    main()
