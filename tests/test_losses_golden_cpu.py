"""The plain-torch restatements in tests/ref_torch.py against tests/golden/losses_ref.npz, which was produced by
EXECUTING the reference's own functions (tests/golden/make_golden_losses.py: batch_sample, calc_point_to_line,
batch_point_to_point / _surface, batch_calc_edge, batch_get_lap_info compiled from /root/reference/utils.py).
This pins the restatements the GPU tests also use; the GPU tests compare the kernels with the same goldens directly."""
import os

import numpy as np
import torch

import ref_torch as R
from helpers import GOLDEN, ROOT  # noqa: F401

G = np.load(os.path.join(GOLDEN, "losses_ref.npz"))


def _inputs():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_losses", os.path.join(GOLDEN, "make_golden_losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.inputs()


def test_sampler_restatement_matches_reference():
    verts, faces, _ = _inputs()
    v = verts.clone().requires_grad_(True)
    pts = R.batch_sample_given(v, faces, torch.from_numpy(G["sample_choices"]), torch.from_numpy(G["sample_r1"]),
                               torch.from_numpy(G["sample_r2"]))
    assert np.array_equal(pts.detach().numpy().view(np.uint32), G["sample_points"].view(np.uint32))
    w = torch.linspace(0.5, 1.5, pts.numel()).view_as(pts)
    (pts * w).sum().backward()
    np.testing.assert_allclose(v.grad.numpy(), G["sample_dverts"], rtol=1e-6, atol=1e-7)


def test_point_to_line_restatement_matches_reference():
    verts, faces, gt = _inputs()
    opt, idx = torch.from_numpy(G["ptl_options"]), torch.from_numpy(G["ptl_index"])
    v = verts.clone().requires_grad_(True)
    d2 = R.point_to_surface_term(gt, v, faces, opt, idx)
    np.testing.assert_allclose(float(d2), float(G["ptl_dist"]), rtol=1e-6)
    d2.backward()
    # the golden holds gradients w.r.t. the gathered triangle vertices: scatter them back to the mesh vertices
    want = torch.zeros(2 * verts.shape[1], 3)
    flat = (idx.long() + faces.shape[0] * torch.arange(2).unsqueeze(-1)).view(-1)
    fi = faces[flat % faces.shape[0]] + (flat // faces.shape[0]).unsqueeze(-1) * verts.shape[1]
    for k, key in enumerate(("ptl_da", "ptl_db", "ptl_dc")):
        want.index_add_(0, fi[:, k], torch.from_numpy(G[key]))
    np.testing.assert_allclose(v.grad.view(-1, 3).numpy(), want.numpy(), rtol=1e-5, atol=1e-7)


def test_oracle_search_feeds_the_reference_functions(oracle):
    """The options / indices stored in the golden are the oracle's (= the reference kernel's, pinned on the GPU box)."""
    verts, faces, gt = _inputs()
    t1, t2, t3 = (torch.index_select(verts, 1, faces[:, k]).contiguous() for k in range(3))
    _, p, i = oracle.tri_fwd(gt.numpy(), t1.numpy(), t2.numpy(), t3.numpy())
    assert np.array_equal(p, G["ptl_options"]) and np.array_equal(i, G["ptl_index"])
    assert set(np.unique(p).tolist()) == set(range(7))
