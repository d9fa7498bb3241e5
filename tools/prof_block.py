#!/usr/bin/env python
"""Where does the deformation block's fwd+bwd time go?  Wall time per eager step vs the same step replayed from a CUDA
graph (pure device time), plus a cProfile of the host side (dev tool)."""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometrics_b200 as g
from geometrics_b200 import synth
dev = torch.device("cuda:0")
B = 16
adj = g.adj_init(synth.uv_sphere(32, 15)[1])["adj"].to(dev)
block = g.BatchMeshDeformationBlock(963, 482).to(dev)
pos = synth.mesh_batch(synth.uv_sphere(32, 15)[0], B).to(dev).requires_grad_(True)
pooled = (0.1 * torch.randn(B, 482, 960, device=dev)).requires_grad_(True)
def blk():
    block.zero_grad(set_to_none=True)
    pos.grad = pooled.grad = None
    f_, c_ = block(pos, pooled, adj)
    (c_.sum() + f_.mean()).backward()
for _ in range(5): blk()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): blk()
torch.cuda.synchronize()
print("eager wall ms/step", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20): blk()
host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
print("eager host-enqueue ms/step", host)
# graph
try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): blk()
    torch.cuda.current_stream().wait_stream(s)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        blk()
    for _ in range(3): gph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): gph.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph replay ms/step (device time)", e0.elapsed_time(e1) / 20)
except Exception as ex:
    print("graph capture failed:", repr(ex)[:300])
pr = cProfile.Profile(); pr.enable()
for _ in range(10): blk()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:4500])
