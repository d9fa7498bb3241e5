"""GPU parity tests of the sampler, the Chamfer-loss kernels and the point-to-surface loss kernels
against (1) plain-torch fp32 restatements of the reference expressions (tests/ref_torch.py) run on
the same device with the same random draws and (2) the CPU oracle (double accumulation).
Tolerance: <=1e-5 relative (BASELINE north_star) unless bit-exactness is stated."""
import numpy as np
import pytest
import torch

import ref_torch as R
from helpers import assert_same_f32

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _mesh(cuda, b=3, seg=32, rings=15, seed=5):
    from geometrics_b200 import synth

    v, f = synth.uv_sphere(seg, rings)
    return synth.mesh_batch(v, b, 0.01, seed).to(cuda), f.to(cuda)


# ------------------------------------------------------------------------------ sampler
def test_face_areas(cuda):
    import geometrics_b200.utils as U

    verts, faces = _mesh(cuda)
    raw = U.face_areas(verts, faces, normalise=False)
    assert_same_f32(raw.cpu().numpy(), R.face_areas(verts, faces, normalise=False).cpu().numpy(), "raw areas")
    norm = U.face_areas(verts, faces)
    torch.testing.assert_close(norm, R.face_areas(verts, faces), rtol=1e-6, atol=0)
    torch.testing.assert_close(norm.sum(1), torch.ones(3, device=cuda), rtol=1e-5, atol=0)


def test_sample_forward_bit_exact_and_backward(cuda):
    import geometrics_b200.utils as U

    verts, faces = _mesh(cuda)
    b, num = verts.shape[0], 4000
    g = torch.Generator(device=cuda).manual_seed(3)
    choices = torch.randint(0, faces.shape[0], (b * num,), device=cuda, generator=g)
    r1 = torch.rand(b * num, 1, device=cuda, generator=g)
    r2 = torch.rand(b * num, 1, device=cuda, generator=g)
    v1 = verts.clone().requires_grad_(True)
    v2 = verts.clone().requires_grad_(True)
    got = U.batch_sample(v1, faces, num, choices=choices, r1=r1, r2=r2)
    want = R.batch_sample_given(v2, faces, choices, r1, r2)
    assert_same_f32(got.detach().cpu().numpy(), want.detach().cpu().numpy(), "sampled points")
    w = torch.randn(b, num, 3, device=cuda, generator=g)
    (got * w).sum().backward()
    (want * w).sum().backward()
    torch.testing.assert_close(v1.grad, v2.grad, rtol=RTOL, atol=1e-6)


def test_sample_rng_order_and_distribution(cuda):
    import geometrics_b200.utils as U

    verts, faces = _mesh(cuda, b=2)
    num = 20000
    # reference_rng=True consumes the CUDA generator exactly as utils.py:606-628 does
    torch.manual_seed(41)
    pts = U.batch_sample(verts, faces, num, reference_rng=True)
    torch.manual_seed(41)
    areas = R.face_areas(verts, faces)
    choices = torch.cat([torch.multinomial(a, num, True) for a in areas])
    dist_uni = torch.distributions.Uniform(torch.tensor([0.0], device=cuda), torch.tensor([1.0], device=cuda))
    r1 = dist_uni.sample((2 * num,))
    r2 = dist_uni.sample((2 * num,))
    want = R.batch_sample_given(verts, faces, choices, r1, r2)
    same = (pts == want).all(-1).float().mean().item()
    assert same > 0.999, same  # a last-ulp difference in a normalised area may flip a rare draw
    # chi-square of face choice frequencies against the areas
    torch.manual_seed(7)
    a = U.face_areas(verts, faces)
    ch = torch.multinomial(a, 200000, True)
    for i in range(2):
        cnt = torch.bincount(ch[i], minlength=faces.shape[0]).double()
        exp = a[i].double() * 200000
        chi2 = float(((cnt - exp) ** 2 / exp).sum())
        assert chi2 < faces.shape[0] + 6 * (2 * faces.shape[0]) ** 0.5, chi2
    # samples lie on the surface: distance to the chosen triangle's plane is ~0
    assert pts.shape == (2, num, 3) and torch.isfinite(pts).all()


# ------------------------------------------------------------------------------ chamfer loss
@pytest.mark.parametrize("shape", [(2, 700, 900), (1, 1024, 1024), (3, 2466, 3000), (1, 5, 3)])
def test_chamfer_loss_fwd_bwd(oracle, cuda, shape):
    import geometrics_b200 as g
    import geometrics_b200.utils as U
    from geometrics_b200 import synth

    b, n, m = shape
    gt = synth.random_points(b, n, 1).to(cuda)
    pr = synth.random_points(b, m, 2).to(cuda).requires_grad_(True)
    sums, idx1, idx2 = U.chamfer_loss_sums(gt, pr)
    # forward vs the oracle's double sums on the same indices, and vs the torch restatement
    s1, s2 = oracle.chamfer_loss_sums(gt.cpu().numpy(), pr.detach().cpu().numpy(), idx1.cpu().numpy(), idx2.cpu().numpy())
    np.testing.assert_allclose(sums.detach().cpu().numpy(), [s1, s2], rtol=RTOL)
    pr2 = pr.detach().clone().requires_grad_(True)
    d1, d2 = R.chamfer_terms(gt, pr2, idx1, idx2)
    torch.testing.assert_close(sums[1] / (b * m), d1, rtol=RTOL, atol=0)
    torch.testing.assert_close(sums[0] / (b * n), d2, rtol=RTOL, atol=0)
    # backward vs autograd of the torch restatement, and vs the oracle's restated nnd_backward
    w = torch.tensor([0.7, 1.3], device=cuda)
    (sums * w).sum().backward()
    (d2 * (0.7 * b * n) + d1 * (1.3 * b * m)).backward()
    torch.testing.assert_close(pr.grad, pr2.grad, rtol=1e-4, atol=1e-6)
    _, og2 = oracle.chamfer_loss_bwd(gt.cpu().numpy(), pr.detach().cpu().numpy(), idx1.cpu().numpy(), idx2.cpu().numpy(), 0.7, 1.3)
    np.testing.assert_allclose(pr.grad.cpu().numpy(), og2, rtol=1e-4, atol=1e-6)


def test_chamfer_loss_grad_both_inputs_and_dist_outputs(oracle, cuda):
    import geometrics_b200 as g
    from geometrics_b200 import _lib, synth

    b, n, m = 2, 300, 411
    x1 = synth.random_points(b, n, 5).to(cuda)
    x2 = synth.random_points(b, m, 6).to(cuda)
    d1, i1, d2, i2 = g.chamfer_forward(x1, x2)
    L = g.lib()
    sums = torch.empty(2, device=cuda)
    o1, o2 = torch.empty(b, n, device=cuda), torch.empty(b, m, device=cuda)
    wsb = L.geom_chamfer_loss_workspace_bytes(b, n, m)
    ws = torch.empty(wsb // 8 + 2, dtype=torch.float64, device=cuda)
    _lib.check(L.geom_chamfer_loss_fwd(b, n, x1.data_ptr(), m, x2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                       o1.data_ptr(), o2.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel() * 8, None), "fwd")
    torch.cuda.synchronize()
    # the per-point terms are exactly the search's own distances (same rounding sequence)
    assert torch.equal(o1, d1) and torch.equal(o2, d2)
    g1, g2 = torch.empty_like(x1), torch.empty_like(x2)
    w = torch.tensor([0.25, 0.5], device=cuda)
    _lib.check(L.geom_chamfer_loss_bwd(b, n, x1.data_ptr(), m, x2.data_ptr(), i1.data_ptr(), i2.data_ptr(), w.data_ptr(),
                                       g1.data_ptr(), g2.data_ptr(), None), "bwd")
    og1, og2 = oracle.chamfer_loss_bwd(x1.cpu().numpy(), x2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy(), 0.25, 0.5)
    np.testing.assert_allclose(g1.cpu().numpy(), og1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g2.cpu().numpy(), og2, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------ point to surface
def _all_option_inputs(cuda, b=2, n=3000, seed=9):
    verts, faces = _mesh(cuda, b=b, seed=seed)
    g = torch.Generator().manual_seed(seed)
    # points near the surface (all seven regions occur) plus far-away ones
    pts = (0.45 * torch.nn.functional.normalize(torch.randn(b, n, 3, generator=g), dim=-1)
           * (1 + 0.15 * torch.randn(b, n, 1, generator=g))).to(cuda)
    return pts, verts, faces


def test_tri_loss_fwd_bwd_vs_torch(cuda):
    import geometrics_b200.utils as U

    pts, verts, faces = _all_option_inputs(cuda)
    v1 = verts.clone().requires_grad_(True)
    total, option, index = U.point_to_surface_sum(pts, v1, faces)
    assert {0, 4, 5} <= set(option.unique().tolist())  # interior + edge regions (all 7 are forced in the next test)
    v2 = verts.clone().requires_grad_(True)
    want = R.point_to_surface_term(pts, v2, faces, option, index)
    b, n, _ = pts.shape
    torch.testing.assert_close(total[0] / (b * n), want, rtol=RTOL, atol=0)
    (total[0] / (b * n) * 3000).backward()
    (want * 3000).backward()
    torch.testing.assert_close(v1.grad, v2.grad, rtol=2e-4, atol=2e-6)


def test_tri_loss_forced_options_vs_torch(cuda):
    """Every option on every row (not only the one the search picked): exercises each backward branch
    on generic, non-orthogonal configurations."""
    from geometrics_b200 import _lib
    import geometrics_b200 as g

    pts, verts, faces = _all_option_inputs(cuda, b=1, n=1400, seed=4)
    n = pts.shape[1]
    gen = torch.Generator().manual_seed(1)
    option = (torch.arange(n) % 7).int().view(1, n).to(cuda)
    index = torch.randint(0, faces.shape[0], (1, n), generator=gen).int().to(cuda)
    L = g.lib()
    total = torch.empty(1, device=cuda)
    dist = torch.empty(1, n, device=cuda)
    closest = torch.empty(1, n, 3, device=cuda)
    ws = torch.empty(L.geom_tri_loss_workspace_bytes(1, n) // 8 + 1, dtype=torch.float64, device=cuda)
    _lib.check(L.geom_tri_loss_fwd(1, n, pts.data_ptr(), verts.shape[1], verts.data_ptr(), faces.shape[0], faces.data_ptr(),
                                   option.data_ptr(), index.data_ptr(), dist.data_ptr(), closest.data_ptr(), total.data_ptr(),
                                   ws.data_ptr(), ws.numel() * 8, None), "fwd")
    v2 = verts.clone().requires_grad_(True)
    want = R.point_to_surface_term(pts, v2, faces, option, index)
    torch.testing.assert_close(total[0] / n, want, rtol=RTOL, atol=0)
    torch.testing.assert_close(dist.sum() / n, want, rtol=1e-4, atol=0)
    torch.testing.assert_close(((closest - pts) ** 2).sum(-1), dist, rtol=1e-4, atol=1e-9)
    want.backward()
    gv = torch.empty_like(verts)
    w = torch.tensor([1.0 / n], device=cuda)
    _lib.check(L.geom_tri_loss_bwd(1, n, pts.data_ptr(), verts.shape[1], verts.data_ptr(), faces.shape[0], faces.data_ptr(),
                                   option.data_ptr(), index.data_ptr(), w.data_ptr(), gv.data_ptr(), None), "bwd")
    torch.testing.assert_close(gv, v2.grad, rtol=2e-4, atol=1e-7)


# ------------------------------------------------------------------------------ composed losses
@pytest.mark.parametrize("f1", [False, True])
def test_batch_point_to_point_and_surface(cuda, f1):
    """The mirrored utils functions against the reference composition with identical random draws."""
    import geometrics_b200.utils as U
    import geometrics_b200 as g

    verts, faces = _mesh(cuda, b=4, seed=12)
    adj_info = {"faces": faces}
    num = 3000
    gt = (0.45 * torch.nn.functional.normalize(torch.randn(4, num, 3, generator=torch.Generator().manual_seed(2)), dim=-1)).to(cuda)

    def reference(surface):
        v = verts.clone().requires_grad_(True)
        areas = U.face_areas(v, faces)
        choices = torch.multinomial(areas, num, True).view(-1)
        r1 = torch.rand(4 * num, 1, device=cuda)
        r2 = torch.rand(4 * num, 1, device=cuda)
        pred = R.batch_sample_given(v, faces, choices, r1, r2)
        _, id_p, _, id_g = g.chamfer_forward(gt, pred.detach())
        dist_1, dist_2 = R.chamfer_terms(gt, pred, id_p, id_g)
        if surface:
            t1, t2, t3 = (torch.index_select(v.detach(), 1, faces[:, k]) for k in range(3))
            _, opt, idx = g.tri_forward(gt, t1, t2, t3)
            dist_2 = R.point_to_surface_term(gt, v, faces, opt, idx)
        loss = (dist_1 + dist_2) * 3000
        loss.backward()
        return loss.detach(), v.grad, R.f_score(gt, pred.detach(), id_p, id_g, num)

    for surface, fn in ((False, U.batch_point_to_point), (True, U.batch_point_to_surface)):
        torch.manual_seed(123)
        want_loss, want_grad, want_f1 = reference(surface)
        torch.manual_seed(123)
        v = verts.clone().requires_grad_(True)
        out = fn(v, adj_info, gt, num=num, f1=f1)
        loss, score = out if f1 else (out, None)
        loss.backward()
        torch.testing.assert_close(loss.detach(), want_loss, rtol=RTOL, atol=0)
        torch.testing.assert_close(v.grad, want_grad, rtol=5e-4, atol=1e-5)
        if f1:
            assert abs(score - want_f1) < 1e-9
