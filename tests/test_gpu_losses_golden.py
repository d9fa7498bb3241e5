"""The CUDA kernels of rows a1 / a6 / a10 / f2 against tests/golden/losses_ref.npz -- the outputs of the REFERENCE's own
functions (tests/golden/make_golden_losses.py executes batch_sample, batch_point_to_point, batch_point_to_surface,
calc_point_to_line, batch_calc_edge and batch_get_lap_info from /root/reference/utils.py with their random draws
recorded).  Tolerances: sampler forward bit-exact; everything else <= 1e-5 relative (BASELINE north_star)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "losses_ref.npz"))
NUM = 500


def _inputs(cuda):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_losses", os.path.join(GOLDEN, "make_golden_losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    verts, faces, gt = mod.inputs()
    return verts.to(cuda), faces.to(cuda), gt.to(cuda)


def _draws(tag, cuda):
    return tuple(torch.from_numpy(G[f"{tag}_{k}"]).to(cuda) for k in ("choices", "r1", "r2"))


def test_batch_sample_matches_reference(cuda):
    import geometrics_b200.utils as U

    verts, faces, _ = _inputs(cuda)
    v = verts.clone().requires_grad_(True)
    c, r1, r2 = _draws("sample", cuda)
    import ref_torch as R

    pts = U.batch_sample(v, faces, NUM, choices=c, r1=r1, r2=r2)
    got, want = pts.detach().cpu().numpy(), G["sample_points"]
    # The golden was computed by the reference's expression on the CPU.  The kernel is bit-identical to that expression
    # evaluated by torch on the GPU; CPU and GPU torch may differ in the last bit of a value, hence one ulp here.
    on_gpu = R.batch_sample_given(verts, faces, c, r1, r2).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), on_gpu.view(np.uint32)), "kernel vs the torch expression on the same device"
    ulp = np.spacing(np.abs(want).astype(np.float32))
    bad = np.abs(got - want) > ulp
    print("values differing from the CPU golden at all:", int((got != want).sum()), "of", got.size)
    assert not bad.any(), f"{int(bad.sum())} of {got.size} sampled coordinates differ from the reference by more than 1 ulp"
    w = torch.linspace(0.5, 1.5, pts.numel()).view_as(pts).to(cuda)
    (pts * w).sum().backward()
    np.testing.assert_allclose(v.grad.cpu().numpy(), G["sample_dverts"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["p2p", "p2s"])
def test_losses_match_reference(cuda, tag):
    import geometrics_b200 as g
    import geometrics_b200.utils as U

    verts, faces, gt = _inputs(cuda)
    info = g.adj_init(faces)
    v = verts.clone().requires_grad_(True)
    fn = U.batch_point_to_point if tag == "p2p" else U.batch_point_to_surface
    g_in = gt[:, :NUM] if tag == "p2p" else gt
    loss, f1 = fn(v, info, g_in, num=NUM, f1=True, sample_draws=_draws(tag, cuda))
    np.testing.assert_allclose(float(loss), float(G[f"{tag}_loss"]), rtol=1e-5)
    assert abs(float(f1) - float(G[f"{tag}_f1"])) <= 1e-9
    loss.backward()
    np.testing.assert_allclose(v.grad.cpu().numpy(), G[f"{tag}_dverts"], rtol=1e-4, atol=2e-5 * np.abs(G[f"{tag}_dverts"]).max())


def test_point_to_line_matches_reference(cuda):
    """geom_tri_fwd -> geom_tri_loss_fwd/_bwd against calc_point_to_line (utils.py:506-550) on all seven options."""
    import geometrics_b200.utils as U

    verts, faces, gt = _inputs(cuda)
    v = verts.clone().requires_grad_(True)
    total, opt, idx = U.point_to_surface_sum(gt, v, faces)
    assert np.array_equal(opt.cpu().numpy(), G["ptl_options"]) and np.array_equal(idx.cpu().numpy(), G["ptl_index"])
    np.testing.assert_allclose(float(total) / (gt.shape[0] * gt.shape[1]), float(G["ptl_dist"]), rtol=1e-5)
    (total / (gt.shape[0] * gt.shape[1])).backward()
    want = torch.zeros(2 * verts.shape[1], 3)
    idx_c, faces_c = torch.from_numpy(G["ptl_index"]), faces.cpu()
    flat = (idx_c.long() + faces_c.shape[0] * torch.arange(2).unsqueeze(-1)).view(-1)
    fi = faces_c[flat % faces_c.shape[0]] + (flat // faces_c.shape[0]).unsqueeze(-1) * verts.shape[1]
    for k, key in enumerate(("ptl_da", "ptl_db", "ptl_dc")):
        want.index_add_(0, fi[:, k], torch.from_numpy(G[key]))
    np.testing.assert_allclose(v.grad.view(-1, 3).cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5 * float(want.abs().max()))


def test_regularisers_match_reference(cuda):
    import geometrics_b200 as g
    import geometrics_b200.utils as U

    verts, faces, _ = _inputs(cuda)
    info = g.adj_init(faces)
    v = verts.clone().requires_grad_(True)
    e = U.batch_calc_edge(v, info)
    np.testing.assert_allclose(float(e), float(G["edge_loss"]), rtol=1e-5)
    e.backward()
    np.testing.assert_allclose(v.grad.cpu().numpy(), G["edge_dverts"], rtol=1e-5, atol=1e-9)
    v = verts.clone().requires_grad_(True)
    lap = U.batch_get_lap_info(v, info)
    np.testing.assert_allclose(lap.detach().cpu().numpy(), G["lap"], rtol=1e-5, atol=1e-6)
    (lap ** 2).sum().backward()
    np.testing.assert_allclose(v.grad.cpu().numpy(), G["lap_dverts"], rtol=1e-5, atol=1e-6)
