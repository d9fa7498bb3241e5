import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g
from geometrics_b200 import synth
dev = torch.device("cuda:0"); B = 16
adj = g.adj_init(synth.uv_sphere(32, 15)[1])["adj"].to(dev)
block = g.BatchMeshDeformationBlock(963, 482).to(dev)
pos = synth.mesh_batch(synth.uv_sphere(32, 15)[0], B).to(dev).requires_grad_(True)
pooled = (0.1 * torch.randn(B, 482, 960, device=dev)).requires_grad_(True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    block.zero_grad(set_to_none=True); pos.grad = pooled.grad = None
    f_, c_ = block(pos, pooled, adj); (c_.sum() + f_.mean()).backward()
torch.cuda.synchronize(); print("done")
