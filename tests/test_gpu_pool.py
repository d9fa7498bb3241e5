"""GPU: the fused vertex-pooling kernel (geom_pool_fwd / geom_pool_bwd) against the torch
restatement of utils.py:316-389 (tests/ref_torch.py, itself pinned to the reference's own output on CPU by
tests/test_pool_cpu.py) and against that golden output directly."""
import os

import numpy as np
import pytest
import torch

import ref_torch as R
from golden import make_golden_pool as G
from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_forward_bit_exact_vs_reference_golden(cuda):
    """Same projected coordinates (computed on the CPU like the golden run) -> the reference's bits."""
    import geometrics_b200.utils as U

    gold = np.load(os.path.join(GOLDEN, "pool_ref.npz"))
    verts, img_info, blocks = G.inputs()
    xs, ys = R.project_vertices(verts, img_info)
    got = U.pool_features([b.to(cuda) for b in blocks], xs.to(cuda), ys.to(cuda))
    assert np.array_equal(got.cpu().numpy(), gold["feats"])
    assert int((got.abs().sum(-1) == 0).sum()) >= 1  # clamped vertices: the floor == ceil zero-weight quirk


@pytest.mark.parametrize("shape", [(2, 96, ((5, 56), (7, 28), (33, 14), (40, 7))), (3, 482, ((64, 56), (128, 28), (256, 14), (512, 7))),
                                   (1, 31, ((3, 9),)), (2, 33, ((1, 5), (70, 3)))])
def test_forward_and_backward_vs_torch(cuda, shape, monkeypatch):
    import geometrics_b200.utils as U

    b, v, maps = shape
    g = torch.Generator().manual_seed(3)
    verts = (torch.rand(b, v, 3, generator=g) - 0.5)
    verts[:, :4] *= 3.0
    img_info = torch.stack((torch.rand(b, generator=g) * 360., 25. + 5. * torch.rand(b, generator=g),
                            0.6 + 0.4 * torch.rand(b, generator=g)), dim=1).to(cuda)
    blocks = [torch.randn(b, c, d, d, generator=g).to(cuda) for c, d in maps]
    up = torch.randn(b, v, sum(c for c, _ in maps), generator=g).to(cuda)
    res = []
    for fn in (U.batched_pooling, R.batched_pooling):
        p = verts.to(cuda).requires_grad_(True)
        bl = [x.clone().requires_grad_(True) for x in blocks]
        f = fn(bl, p, img_info)
        (f * up).sum().backward()
        res.append((f.detach(), p.grad, [x.grad for x in bl]))
    (f0, dp0, db0), (f1, dp1, db1) = res
    # the fused projection equals torch's to fp32 rounding, the bilinear gather is continuous in the coordinates
    assert float((f0 - f1).abs().max()) <= 2e-4 * float(f1.abs().max()), f"forward differs: {float((f0 - f1).abs().max())}"
    # d(features)/d(position) is piecewise constant in the texel cell: a coordinate within rounding of a texel
    # boundary may sit in the neighbouring cell on one of the two paths, so allow a handful of outlier vertices
    scale = float(dp1.abs().max())
    outliers = ((dp0 - dp1).abs().amax(dim=-1) > 1e-3 * scale).float().mean()
    assert float(outliers) <= 0.005, (float(outliers), float((dp0 - dp1).abs().max()), scale)
    for a, r in zip(db0, db1):
        assert float((a - r).abs().max()) <= 2e-4 * max(1.0, float(r.abs().max()))
    # given the SAME projected coordinates the gather is bit-identical to the torch expression
    xs, ys = R.project_vertices(verts.to(cuda), img_info)
    assert torch.equal(U.pool_features(blocks, xs, ys), R.pool_features(blocks, xs, ys))
    monkeypatch.setattr(U, "FUSED_PROJECTION", False)
    assert torch.equal(U.batched_pooling(blocks, verts.to(cuda), img_info), R.batched_pooling(blocks, verts.to(cuda), img_info))


def test_fused_projection_vs_torch(cuda):
    """geom_project_fwd / _bwd against the reference expression evaluated by torch on the same device."""
    import geometrics_b200.utils as U

    g = torch.Generator().manual_seed(9)
    verts = (torch.rand(5, 700, 3, generator=g) - 0.5).to(cuda)
    info = torch.stack((torch.rand(5, generator=g) * 720. - 180., 25. + 5. * torch.rand(5, generator=g),
                        0.6 + 0.4 * torch.rand(5, generator=g)), dim=1).to(cuda)
    gx, gy = torch.randn(5, 700, generator=g).to(cuda), torch.randn(5, 700, generator=g).to(cuda)
    out = []
    for fn in (U.project_vertices, R.project_vertices):
        p = verts.clone().requires_grad_(True)
        xs, ys = fn(p, info)
        (xs * gx + ys * gy).sum().backward()
        out.append((xs.detach(), ys.detach(), p.grad))
    for a, r in zip(*out):
        assert float((a - r).abs().max()) <= 2e-5 * max(1.0, float(r.abs().max())), float((a - r).abs().max())


def test_gather_into_block_input_buffer(cuda):
    """out / col0: features written next to the vertex positions (what models.py:239 concatenates)."""
    import geometrics_b200.utils as U

    verts, img_info, blocks = G.inputs()
    verts, img_info, blocks = verts.to(cuda), img_info.to(cuda), [b.to(cuda) for b in blocks]
    xs, ys = U.project_vertices(verts, img_info)
    want = U.pool_features(blocks, xs, ys)
    buf = torch.full((2, 96, 3 + want.shape[2] + 2), -7.0, device=cuda)
    buf[:, :, :3] = verts
    out = U.pool_features(blocks, xs, ys, out=buf, col0=3)
    assert out.data_ptr() == buf.data_ptr()
    assert torch.equal(buf[:, :, 3:3 + want.shape[2]], want) and torch.equal(buf[:, :, :3], verts)
    assert bool((buf[:, :, 3 + want.shape[2]:] == -7.0).all())


def test_gather_into_buffer_keeps_the_gradient_of_the_prefilled_columns(cuda):
    """Training pattern: buf[:, :, :3] = verts, then pool_features(..., out=buf, col0=3).  The gradient that reaches the
    position columns of the buffer must flow back to verts (and through xs / ys and the maps as without `out`)."""
    import geometrics_b200.utils as U

    verts, img_info, blocks = G.inputs()
    img_info = img_info.to(cuda)
    w = None
    grads = []
    for use_out in (False, True):
        v = verts.to(cuda).clone().requires_grad_(True)
        blk = [b.to(cuda).clone().requires_grad_(True) for b in blocks]
        xs, ys = U.project_vertices(v, img_info)
        if use_out:
            c = sum(b.shape[1] for b in blk)
            buf = torch.zeros(2, 96, 3 + c, device=cuda)
            buf[:, :, :3] = v          # in-place copy of a tensor that requires grad: buf now carries autograd history
            full = U.pool_features(blk, xs, ys, out=buf, col0=3)
        else:
            full = torch.cat((v, U.pool_features(blk, xs, ys)), dim=2)
        if w is None:
            w = torch.randn(full.shape, device=cuda, generator=torch.Generator(device=cuda).manual_seed(5))
        (full * w).sum().backward()
        grads.append([v.grad] + [b.grad for b in blk])
    for a, r in zip(*grads):
        assert float((a - r).abs().max()) <= 2e-5 * max(1.0, float(r.abs().max())), float((a - r).abs().max())


def test_argument_validation(cuda):
    import geometrics_b200.utils as U

    xs = torch.rand(2, 10, device=cuda)
    with pytest.raises(RuntimeError):
        U.pool_features([torch.rand(2, 3, 4, 4)], xs, xs)            # CPU map
    with pytest.raises(ValueError):
        U.pool_features([torch.rand(2, 3, 4, 5, device=cuda)], xs, xs)  # not square
    with pytest.raises(ValueError):
        U.pool_features([torch.rand(3, 3, 4, 4, device=cuda)], xs, xs)  # batch mismatch
