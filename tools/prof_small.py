#!/usr/bin/env python
"""ONE launch of every small kernel of the path at its training shape, bracketed by cudaProfilerStart/Stop, for a single
short ncu session (the big search kernels have their own captures):
    ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy \
        --section LaunchStats -k regex:'<names>' -o gpurun_out/r02_small python tools/prof_small.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g  # noqa: E402
from geometrics_b200 import synth  # noqa: E402
import geometrics_b200.utils as U  # noqa: E402
from geometrics_b200.layers import gcn_bn_stage, csr_from_dense  # noqa: E402

dev = torch.device("cuda:0")
vs, fs = synth.uv_sphere(32, 15)
fs = fs.to(dev)
info = g.adj_init(fs)
csr = csr_from_dense(info["adj"])
gt = (0.45 * torch.nn.functional.normalize(torch.randn(16, 3000, 3), dim=-1)).to(dev)
img_info = torch.stack((torch.rand(16, device=dev) * 360., 25. + 5. * torch.rand(16, device=dev), 0.6 + 0.4 * torch.rand(16, device=dev)), dim=1)
maps = [torch.randn(16, c, d, d, device=dev, requires_grad=True) for c, d in ((64, 56), (128, 28), (256, 14), (512, 7))]
layer = g.Batch_Image_ZERON_GCNGCN(192, 192).to(dev)
bn = torch.nn.BatchNorm1d(482).to(dev)
x1, x2 = synth.random_points(16, 3000, 41).to(dev), synth.random_points(16, 3000, 1041).to(dev)


def one_pass():
    pv = synth.mesh_batch(vs, 16).to(dev).requires_grad_(True)
    loss, f1 = U.batch_point_to_surface(pv, info, gt, num=3000, f1=True)      # area, sample, searches, loss, closest (+bwd)
    reg = U.batch_calc_edge(pv, info) + (U.batch_get_lap_info(pv, info) ** 2).mean()   # edge fwd/bwd, CSR SpMM
    pooled = U.batched_pooling(maps, pv, img_info)                              # projection + gather (+bwd)
    feat = pooled[..., :192].contiguous().requires_grad_(True)
    out = gcn_bn_stage(feat, layer, bn, csr, res=feat)                          # GEMM, SpMM, BN stage (+ their backward)
    (loss + reg + out.mean() + pooled.mean()).backward()
    g.chamfer_forward(x1, x2, tensor=False, want_dist=False)                    # CUDA-core two-way + refine at the training size
    torch.cuda.synchronize()


for _ in range(2):
    one_pass()
torch.cuda.profiler.start()
one_pass()
torch.cuda.profiler.stop()
print("done")
