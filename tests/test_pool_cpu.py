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
