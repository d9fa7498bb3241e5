"""CPU: the torch restatement of the reference's vertex pooling (tests/ref_torch.py) against golden vectors
produced by the reference's own batched_pooling / batch_camera_info (tests/golden/make_golden_pool.py)."""
import os

import numpy as np
import torch

import ref_torch as R
from golden import make_golden_pool as G
from helpers import GOLDEN


def test_restatement_matches_reference_golden():
    gold = np.load(os.path.join(GOLDEN, "pool_ref.npz"))
    verts, img_info, blocks = G.inputs()
    cam_mat, cam_pos = R.batch_camera_info(img_info)
    assert np.array_equal(cam_mat.numpy(), gold["cam_mat"]) and np.array_equal(cam_pos.numpy(), gold["cam_pos"])
    verts = verts.clone().requires_grad_(True)
    blocks = [b.clone().requires_grad_(True) for b in blocks]
    feats = R.batched_pooling(blocks, verts, img_info)
    assert np.array_equal(feats.detach().numpy(), gold["feats"])  # same ops in the same order: same bits
    assert int((feats.detach().abs().sum(-1) == 0).sum()) >= 1    # clamped vertices -> the zero-weight quirk
    up = torch.linspace(-1, 1, feats.shape[-1]).view(1, 1, -1) * torch.linspace(0.5, 1.5, feats.shape[1]).view(1, -1, 1)
    (feats * up).sum().backward()
    np.testing.assert_allclose(verts.grad.numpy(), gold["d_verts"], rtol=1e-5, atol=1e-6)
    for i, b in enumerate(blocks):
        np.testing.assert_allclose(b.grad.numpy(), gold[f"d_block{i}"], rtol=1e-5, atol=1e-6)


def test_mirror_projection_matches_reference_golden():
    """geometrics_b200.utils.batch_camera_info / project_vertices are device-agnostic torch: check them on CPU."""
    from geometrics_b200 import utils as U

    gold = np.load(os.path.join(GOLDEN, "pool_ref.npz"))
    verts, img_info, _ = G.inputs()
    cam_mat, cam_pos = U.batch_camera_info(img_info)
    assert np.array_equal(cam_mat.numpy(), gold["cam_mat"]) and np.array_equal(cam_pos.numpy(), gold["cam_pos"])
    xs, ys = U.project_vertices(verts, img_info)
    rx, ry = R.project_vertices(verts, img_info)
    assert torch.equal(xs, rx) and torch.equal(ys, ry)


def _kernel_formulas(verts, info, gxs=None, gys=None):
    """The fused projection kernel's arithmetic (csrc/pool.cu project_fwd_kernel / project_bwd_kernel) restated
    line for line in float64 torch, to check the closed-form camera and the hand-derived gradient on the CPU."""
    import math

    def rem360(x):
        r = torch.fmod(x, 360.)
        return torch.where((r != 0) & (r < 0), r + 360., r)

    p = info.double()
    theta, phi = rem360(p[:, 0] * math.pi / 180.), rem360(p[:, 1] * math.pi / 180.)
    t = p[:, 2] * torch.cos(phi)
    c = torch.stack((t * torch.cos(theta), p[:, 2] * torch.sin(phi), t * torch.sin(theta)), dim=1)
    x = torch.stack((c[:, 2], torch.zeros_like(t), -c[:, 0]), dim=1)
    y = torch.stack((c[:, 1] * x[:, 2] - c[:, 2] * x[:, 1], c[:, 2] * x[:, 0] - c[:, 0] * x[:, 2], c[:, 0] * x[:, 1] - c[:, 1] * x[:, 0]), dim=1)
    ax, ay, az = (v / v.norm(dim=1, keepdim=True) for v in (x, y, c))
    q = verts.double() * .57 - c.unsqueeze(1)
    X = (q * ax.unsqueeze(1)).sum(-1)
    Y = (q * ay.unsqueeze(1)).sum(-1)
    Z = (q * az.unsqueeze(1)).sum(-1)
    xs = ((-Y) / (-Z) * 248. + 112.) / 223.
    ys = (X / (-Z) * 248. + 112.) / 223.
    if gxs is None:
        return xs, ys
    s = 248. / 223.
    dX = -gys.double() * s / Z
    dY = gxs.double() * s / Z
    dZ = (-gxs.double() * Y + gys.double() * X) * s / (Z * Z)
    g = .57 * (dX.unsqueeze(-1) * ax.unsqueeze(1) + dY.unsqueeze(-1) * ay.unsqueeze(1) + dZ.unsqueeze(-1) * az.unsqueeze(1))
    return xs, ys, g


def test_fused_projection_formulas_match_reference_expression():
    verts, img_info, _ = G.inputs()
    img_info = torch.cat((img_info, torch.tensor([[-30., 400., 0.8], [725., -10., 0.7]])))  # negative / > 360 degree angles
    verts = torch.cat((verts, verts.flip(0)))
    v = verts.double().requires_grad_(True)
    xs, ys = R.project_vertices(v, img_info.double())
    gx, gy = torch.randn(xs.shape, dtype=torch.float64), torch.randn(xs.shape, dtype=torch.float64)
    (xs * gx + ys * gy).sum().backward()
    kx, ky, kg = _kernel_formulas(verts, img_info, gx, gy)
    assert torch.allclose(kx, xs.detach(), rtol=1e-12, atol=1e-12) and torch.allclose(ky, ys.detach(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(kg, v.grad, rtol=1e-10, atol=1e-12)
