"""world_size-2 NCCL test on two B200s (-m gpu; skipped on a single-GPU box): the sharded hot path as SURVEY.md 8e states
it -- each rank searches its slice of the batch with the CUDA kernels, the partial loss sums meet in ONE NCCL all-reduce.
Checked: concatenated per-rank indices == the unsharded search, the global loss on every rank == the unsharded loss
(<= 1e-6 relative: fp32 summation order only), the gradient w.r.t. the local points == the unsharded gradient's slice."""
import os
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, b, n, m, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import geometrics_b200 as g
        from geometrics_b200 import sharding, synth
        from geometrics_b200.utils import all_reduce_sums, chamfer_loss_sums

        x1, x2 = synth.random_points(b, n, 41), synth.random_points(b, m, 42)
        a = sharding.shard(x1, rank, world).to(dev).requires_grad_(True)
        c = sharding.shard(x2, rank, world).to(dev)
        sums, i1, i2 = chamfer_loss_sums(a, c)          # search + fused loss sums (CUDA kernels)
        tot = all_reduce_sums(sums, dist.group.WORLD)   # the one collective of the path
        loss = (tot[0] / (b * n) + tot[1] / (b * m)) * 3000
        loss.backward()
        # the tensor-core TriDistance search shards the same way: indices only, no collective
        v, f = synth.uv_sphere(32, 15)
        t1, t2, t3 = synth.face_vertices(synth.mesh_batch(v, b), f)
        pts = synth.random_points(b, 3000, 43)
        tri = g.tri_forward(*(sharding.shard(t, rank, world).to(dev).contiguous() for t in (pts, t1, t2, t3)))
        torch.cuda.synchronize()
        out_q.put((rank, i1.cpu(), i2.cpu(), float(loss.detach()), a.grad.cpu(), tuple(t.cpu() for t in tri)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("b", [4, 5])
def test_sharded_path_on_two_gpus(cuda, b):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    import geometrics_b200 as g
    from geometrics_b200 import synth
    from geometrics_b200.utils import chamfer_loss_sums

    n, m, world = 3000, 2466, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 300) + b
    procs = [ctx.Process(target=_worker, args=(r, world, port, b, n, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # unsharded run on one GPU
    x1 = synth.random_points(b, n, 41).to(cuda).requires_grad_(True)
    x2 = synth.random_points(b, m, 42).to(cuda)
    sums, i1, i2 = chamfer_loss_sums(x1, x2)
    want = (sums[0] / (b * n) + sums[1] / (b * m)) * 3000
    want.backward()
    assert torch.equal(torch.cat([r[1] for r in res]), i1.cpu()) and torch.equal(torch.cat([r[2] for r in res]), i2.cpu())
    for r in res:
        assert abs(r[3] - float(want.detach())) <= 1e-6 * abs(float(want.detach())), (r[3], float(want.detach()))
    # scatter-adds of the backward are fp32 atomics: order-dependent in the last bits
    assert torch.allclose(torch.cat([r[4] for r in res]), x1.grad.cpu(), rtol=1e-5, atol=1e-7)
    v, f = synth.uv_sphere(32, 15)
    t1, t2, t3 = synth.face_vertices(synth.mesh_batch(v, b), f)
    pts = synth.random_points(b, 3000, 43)
    full = g.tri_forward(*(t.to(cuda) for t in (pts, t1, t2, t3)))
    for k in range(3):
        assert torch.equal(torch.cat([r[5][k] for r in res]), full[k].cpu())
