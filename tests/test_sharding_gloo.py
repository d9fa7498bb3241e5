"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, the single all-reduce of
the partial loss sums and its autograd behaviour.  The per-shard sums come from the CPU oracle (the
CUDA kernels are covered by the -m gpu tests); what is checked here is that shard -> partial sums ->
all-reduce -> global mean reproduces the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, b, n, m, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geometrics_b200 import sharding, synth
        from geometrics_b200.utils import all_reduce_sums
        from oracle import Oracle

        orc = Oracle()
        x1, x2 = synth.random_points(b, n, 41), synth.random_points(b, m, 42)
        a, c = sharding.shard(x1, rank, world), sharding.shard(x2, rank, world)
        lo, hi = sharding.shard_bounds(b, rank, world)
        assert a.shape[0] == hi - lo
        _, i1, _, i2 = orc.chamfer_fwd(a.numpy(), c.numpy())
        s1, s2 = orc.chamfer_loss_sums(a.numpy(), c.numpy(), i1, i2)
        part = torch.tensor([s1, s2], dtype=torch.float32, requires_grad=True)
        means = sharding.global_means(part, [a.shape[0] * n, c.shape[0] * m], dist.group.WORLD)
        loss = (means[0] + means[1]) * 3000
        loss.backward()
        # plain sum path as well
        tot = all_reduce_sums(torch.tensor([s1, s2], dtype=torch.float32), dist.group.WORLD)
        assert all_reduce_sums(part.detach()) is not None and torch.equal(all_reduce_sums(part.detach()), part.detach())  # group=None: local
        out_q.put((rank, lo, hi, i1, i2, float(loss), part.grad.tolist(), tot.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("b", [4, 5])
def test_sharded_chamfer_loss_equals_global(oracle, b):
    n, m, world = 300, 257, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + b
    procs = [ctx.Process(target=_worker, args=(r, world, port, b, n, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from geometrics_b200 import synth

    x1, x2 = synth.random_points(b, n, 41).numpy(), synth.random_points(b, m, 42).numpy()
    _, gi1, _, gi2 = oracle.chamfer_fwd(x1, x2)
    s1, s2 = oracle.chamfer_loss_sums(x1, x2, gi1, gi2)
    want = (s1 / (b * n) + s2 / (b * m)) * 3000
    # shards tile the batch, indices concatenate to the global search result
    assert [r[1] for r in res] == [0, res[0][2]] and res[-1][2] == b
    assert (np.concatenate([r[3] for r in res]) == gi1).all() and (np.concatenate([r[4] for r in res]) == gi2).all()
    for r in res:
        assert abs(r[5] - want) <= 1e-6 * abs(want)              # every rank holds the global loss
        np.testing.assert_allclose(r[6], [3000 / (b * n), 3000 / (b * m)], rtol=1e-6)  # d loss / d local sums
        np.testing.assert_allclose(r[7], [s1, s2], rtol=1e-6)


def test_shard_bounds_cover_batch():
    from geometrics_b200.sharding import shard_bounds

    for b in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
