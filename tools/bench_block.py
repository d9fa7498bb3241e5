#!/usr/bin/env python
"""BatchMeshDeformationBlock forward+backward at the reference's training shape (B=16, V=482, 963 input features):
this library (fused BN/ReLU/residual kernel, own GCN backward) vs the reference composition on the same GPU (torch
dense-adjacency GCN expression layers.py:107-116 + nn.BatchNorm1d + F.relu, cuBLAS fp32).  Prints JSON lines."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import _lib, synth  # noqa: E402
import ref_torch as R  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, reps=20, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    for B in (16, 32):
        adj = g.adj_init(synth.uv_sphere(32, 15)[1])["adj"].to(dev)
        block = g.BatchMeshDeformationBlock(963, 482).to(dev)
        pos = synth.mesh_batch(synth.uv_sphere(32, 15)[0], B).to(dev).requires_grad_(True)
        pooled = (0.1 * torch.randn(B, 482, 960, device=dev)).requires_grad_(True)

        def blk(i):
            block.zero_grad(set_to_none=True)
            pos.grad = pooled.grad = None
            f_, c_ = block(pos, pooled, adj)
            (c_.sum() + f_.mean()).backward()

        def fwd_only(i):
            with torch.no_grad():
                block(pos, pooled, adj)

        rec = {"bench": f"BatchMeshDeformationBlock B={B} V=482 (963 -> 192 x13 -> 3)"}
        rec["ours_fwd_bwd_ms"] = round(timeit(blk), 4)
        rec["ours_fwd_ms"] = round(timeit(fwd_only), 4)
        _lib.prof_begin()
        for i in range(3):
            blk(i)
        rec["ours_kernels_ms_per_step"] = {k.split("::", 1)[1]: [round(v[0] / 3, 4), v[1] // 3] for k, v in _lib.prof_end().items()}
        block.FUSED_BN = False
        rec["ours_unfused_bn_fwd_bwd_ms"] = round(timeit(blk), 4)
        orig = g.layers.zeron_gcn
        g.layers.zeron_gcn = lambda x, w_, b_, a, split, act=None: (lambda y: y if act is None else act(y))(
            R.zeron_gcn(x, w_[0] if w_.dim() == 3 else w_, b_, adj, split))
        try:
            rec["reference_fwd_bwd_ms"] = round(timeit(blk), 4)
            rec["reference_fwd_ms"] = round(timeit(fwd_only), 4)
        finally:
            g.layers.zeron_gcn = orig
            block.FUSED_BN = True
        rec["speedup_fwd_bwd"] = round(rec["reference_fwd_bwd_ms"] / rec["ours_fwd_bwd_ms"], 2)
        rec["speedup_fwd"] = round(rec["reference_fwd_ms"] / rec["ours_fwd_ms"], 2)
        # the same step replayed from a CUDA graph (what a training loop with static shapes can do): device time only
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
