"""Host-side mirror of the hot-path functions of the reference's ``utils.py``: same names, argument
meaning and return values, with the chains of small torch kernels replaced by calls into
``libgeom_b200.so`` (sm_100a).  Everything here is CUDA-only; there is no CPU fallback.

Reference map:
  batch_sample              utils.py:590-633
  batch_point_to_point      utils.py:393-438
  batch_point_to_surface    utils.py:441-502  (+ calc_point_to_line :506-550, edge :553-570, Plane :572-587)
  calc_adj / normalize_adj / adj_init   utils.py:96-131
  chamfer_dist / tri_dist   module singletons, utils.py:31-35
"""
from __future__ import annotations

import torch

from . import _lib
from .chamfer_distance import ChamferDistance, chamfer_forward
from .tri_distance import TriDistance, tri_forward

chamfer_dist = ChamferDistance()  # utils.py:32
tri_dist = TriDistance()          # utils.py:35


# ------------------------------------------------------------------------------------------------
# adjacency (one-time setup; plain torch, device-agnostic)
# ------------------------------------------------------------------------------------------------
def calc_adj(faces: torch.Tensor) -> torch.Tensor:
    """Dense binary adjacency with self loops (utils.py:113-131)."""
    v1, v2, v3 = faces[:, 0], faces[:, 1], faces[:, 2]
    num_verts = int(faces.max())
    adj = torch.eye(num_verts + 1, device=faces.device)
    adj[(v1, v2)] = 1
    adj[(v1, v3)] = 1
    adj[(v2, v1)] = 1
    adj[(v2, v3)] = 1
    adj[(v3, v1)] = 1
    adj[(v3, v2)] = 1
    return adj


def normalize_adj(mx: torch.Tensor) -> torch.Tensor:
    """Row-normalise (utils.py:96-101): adj[i, j] = (1 / rowsum_i) * mx[i, j]."""
    r_inv = (1.0 / mx.sum(1)).view(-1)
    r_inv[r_inv != r_inv] = 0.0
    return r_inv.unsqueeze(1) * mx


def adj_init(faces: torch.Tensor) -> dict:
    """utils.py:104-112: {'adj': row-normalised, 'adj_orig': binary, 'faces': faces}."""
    adj = calc_adj(faces)
    return {"adj": normalize_adj(adj), "adj_orig": adj.clone(), "faces": faces}


# ------------------------------------------------------------------------------------------------
# face sampler
# ------------------------------------------------------------------------------------------------
def _faces_i64(faces: torch.Tensor, dev) -> torch.Tensor:
    if faces.dim() != 2 or faces.shape[1] != 3:
        raise ValueError(f"faces must have shape (F, 3), got {tuple(faces.shape)}")
    if faces.dtype != torch.int64 or faces.device != dev or not faces.is_contiguous():
        faces = faces.to(device=dev, dtype=torch.int64).contiguous()
    return faces


def face_areas(verts: torch.Tensor, faces: torch.Tensor, normalise: bool = True) -> torch.Tensor:
    """(B,F) face areas of every mesh (utils.py:596-601), divided by the per-mesh total when
    ``normalise`` (utils.py:602).  No gradient, as in the reference (areas only feed multinomial)."""
    _lib.require_cuda_f32("verts", verts, 3)
    L = _lib.lib()
    verts = verts.detach().contiguous()
    faces = _faces_i64(faces, verts.device)
    b, v, _ = verts.shape
    f = faces.shape[0]
    with _lib.on_device(verts.device):
        areas = torch.empty((b, f), dtype=torch.float32, device=verts.device)
        _lib.check(L.geom_face_areas(b, v, verts.data_ptr(), f, faces.data_ptr(), areas.data_ptr(),
                                     1 if normalise else 0, _lib.current_stream_ptr(verts.device)), "geom_face_areas")
    return areas


class SampleFunction(torch.autograd.Function):
    """points = barycentric combination of the chosen faces' vertices; gradient to ``verts`` only
    (utils.py:614-631: the three index_selects are the only differentiable path)."""

    @staticmethod
    def forward(ctx, verts, faces, choices, r1, r2):
        L = _lib.lib()
        vc = verts.detach().contiguous()
        b, v, _ = vc.shape
        num = choices.numel() // max(b, 1)
        dev = vc.device
        with _lib.on_device(dev):
            points = torch.empty((b, num, 3), dtype=torch.float32, device=dev)
            _lib.check(L.geom_sample_fwd(b, v, vc.data_ptr(), faces.shape[0], faces.data_ptr(), num,
                                         choices.data_ptr(), r1.data_ptr(), r2.data_ptr(), points.data_ptr(),
                                         _lib.current_stream_ptr(dev)), "geom_sample_fwd")
        ctx.save_for_backward(faces, choices, r1, r2)
        ctx.shape = (b, v, num)
        return points

    @staticmethod
    def backward(ctx, grad_points):
        faces, choices, r1, r2 = ctx.saved_tensors
        b, v, num = ctx.shape
        L = _lib.lib()
        g = grad_points.contiguous()
        dev = g.device
        with _lib.on_device(dev):
            gv = torch.empty((b, v, 3), dtype=torch.float32, device=dev)
            _lib.check(L.geom_sample_bwd(b, v, faces.shape[0], faces.data_ptr(), num, choices.data_ptr(),
                                         r1.data_ptr(), r2.data_ptr(), g.data_ptr(), gv.data_ptr(),
                                         _lib.current_stream_ptr(dev)), "geom_sample_bwd")
        return gv, None, None, None, None


def batch_sample(verts: torch.Tensor, faces: torch.Tensor, num: int = 10000, *, choices=None, r1=None, r2=None,
                 reference_rng: bool = False, generator=None) -> torch.Tensor:
    """Area-weighted uniform surface samples (B,num,3), differentiable w.r.t. ``verts``
    (utils.py:590-633).

    Random draws: one batched ``torch.multinomial`` over the normalised areas, then two uniform
    vectors of B*num, in that order.  ``reference_rng=True`` instead issues the per-mesh multinomial
    calls in the reference's order (utils.py:606-612) so that the same CUDA seed gives the same
    draws.  ``choices`` (B*num int64 face ids), ``r1`` and ``r2`` (B*num uniforms; u = sqrt(r1),
    v = r2) can be supplied to make the result a pure function of its inputs (parity tests)."""
    _lib.require_cuda_f32("verts", verts, 3)
    dev = verts.device
    faces = _faces_i64(faces, dev)
    b = verts.shape[0]
    if choices is None:
        areas = face_areas(verts, faces)
        if reference_rng:
            choices = torch.cat([torch.multinomial(a, num, True, generator=generator) for a in areas])
        else:
            choices = torch.multinomial(areas, num, True, generator=generator).view(-1)
    choices = choices.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    if r1 is None:
        r1 = torch.rand(b * num, 1, device=dev, generator=generator)  # dist_uni.sample_n, utils.py:627
    if r2 is None:
        r2 = torch.rand(b * num, 1, device=dev, generator=generator)  # utils.py:628
    r1 = r1.to(device=dev, dtype=torch.float32).contiguous().view(-1)
    r2 = r2.to(device=dev, dtype=torch.float32).contiguous().view(-1)
    if choices.numel() != b * num or r1.numel() != b * num or r2.numel() != b * num:
        raise ValueError("choices / r1 / r2 must hold B*num entries")
    return SampleFunction.apply(verts, faces, choices, r1, r2)


# ------------------------------------------------------------------------------------------------
# Chamfer loss
# ------------------------------------------------------------------------------------------------
class _AllReduceSum(torch.autograd.Function):
    """Sum partial loss sums over the ranks of ``group``; every rank keeps the same upstream gradient.

    Gradient convention: every rank ends up with d(GLOBAL-mean loss)/d(its own points and parameters) -- the backward
    passes the upstream gradient through unchanged.  Summing parameter gradients over ranks (all-reduce SUM) therefore
    gives the exact gradient of the global mean; DistributedDataParallel AVERAGES over ranks, which yields 1/W of it:
    scale the loss by the world size (or the learning rate) when training under DDP with ``group=``."""

    @staticmethod
    def forward(ctx, sums, group):
        import torch.distributed as dist

        out = sums.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        return grad, None


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """Sum over the ranks of ``group``.  ``group=None`` means "not sharded": no communication (pass
    ``torch.distributed.group.WORLD`` explicitly to reduce over the default group)."""
    import torch.distributed as dist

    if group is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    return _AllReduceSum.apply(sums, group)


def world_size(group=None) -> int:
    import torch.distributed as dist

    if group is None or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class ChamferLossFunction(torch.autograd.Function):
    """(xyz1, xyz2, idx1, idx2) -> sums (2,): sums[0] = sum |xyz2[idx1] - xyz1|^2 over xyz1's rows,
    sums[1] = sum |xyz1[idx2] - xyz2|^2 over xyz2's rows.  Gradient to both point sets with the
    indices held constant (what autograd gives for utils.py:413-417)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, idx1, idx2, f1=None):
        """``f1`` = (scale, threshold): also return the per-item F1 counts (B,) int32 x 2 from the same launch."""
        L = _lib.lib()
        a, c = xyz1.detach().contiguous(), xyz2.detach().contiguous()
        b, n, _ = a.shape
        m = c.shape[1]
        dev = a.device
        with _lib.on_device(dev):
            sums = torch.empty(2, dtype=torch.float32, device=dev)
            wsb = L.geom_chamfer_loss_workspace_bytes(b, n, m)
            ws = torch.empty(max(wsb // 16, 1) * 2, dtype=torch.float64, device=dev)
            if f1 is None:
                _lib.check(L.geom_chamfer_loss_fwd(b, n, a.data_ptr(), m, c.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                                                   None, None, sums.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                                   _lib.current_stream_ptr(dev)), "geom_chamfer_loss_fwd")
            else:
                cnt = torch.empty(2, b, dtype=torch.int32, device=dev)
                _lib.check(L.geom_chamfer_loss_f1_fwd(b, n, a.data_ptr(), m, c.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                                                      None, None, sums.data_ptr(), float(f1[0]), float(f1[1]),
                                                      cnt[0].data_ptr(), cnt[1].data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                                      _lib.current_stream_ptr(dev)), "geom_chamfer_loss_f1_fwd")
        ctx.save_for_backward(a, c, idx1, idx2)
        if f1 is None:
            return sums
        ctx.mark_non_differentiable(cnt)
        return sums, cnt

    @staticmethod
    def backward(ctx, grad_sums, *_):
        a, c, idx1, idx2 = ctx.saved_tensors
        L = _lib.lib()
        b, n, _ = a.shape
        m = c.shape[1]
        dev = a.device
        w = grad_sums.to(torch.float32).contiguous()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with _lib.on_device(dev):
            g1 = torch.empty_like(a) if need1 else None
            g2 = torch.empty_like(c) if need2 else None
            _lib.check(L.geom_chamfer_loss_bwd(b, n, a.data_ptr(), m, c.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                                               w.data_ptr(), g1.data_ptr() if need1 else None,
                                               g2.data_ptr() if need2 else None, _lib.current_stream_ptr(dev)),
                       "geom_chamfer_loss_bwd")
        return g1, g2, None, None, None


F1_SCALE, F1_THRESHOLD = 0.57, 1e-2  # utils.py:425-426, :490-491


def chamfer_loss_sums(xyz1: torch.Tensor, xyz2: torch.Tensor, f1: bool = False):
    """NN search + loss sums: -> (sums (2,) differentiable, idx1 (B,n), idx2 (B,m)); with ``f1`` also the per-item
    F1 counts (2,B) int32 (rows of xyz1 / of xyz2 within the threshold of their neighbour) from the same launch."""
    _, idx1, _, idx2 = chamfer_forward(xyz1, xyz2, want_dist=False)
    if not f1:
        return ChamferLossFunction.apply(xyz1, xyz2, idx1, idx2), idx1, idx2
    sums, cnt = ChamferLossFunction.apply(xyz1, xyz2, idx1, idx2, (F1_SCALE, F1_THRESHOLD))
    return sums, idx1, idx2, cnt


def f_score_from_counts(cnt: torch.Tensor, num: int) -> torch.Tensor:
    """F1 (utils.py:428-436) from the fused counts: cnt[0] = ground-truth points whose nearest prediction is within
    the threshold (recall), cnt[1] = predictions near the ground truth (precision), both divided by ``num`` as the
    reference does.  Returns a 0-d float64 DEVICE tensor: nothing synchronises until the caller reads it
    (``float(f)``, formatting, comparison) -- the reference returned a Python float after B host syncs."""
    recall = cnt[0].double() / float(num)
    precision = cnt[1].double() / float(num)
    return (2 * (precision * recall) / (precision + recall + 1e-8)).sum() / cnt.shape[1]


def _draws(sample_draws):
    if sample_draws is None:
        return {}
    choices, r1, r2 = sample_draws
    return {"choices": choices, "r1": r1, "r2": r2}


def batch_point_to_point(pred_vert, adj_info, gt_points, num=1000, f1=False, group=None, sample_draws=None):
    """Chamfer loss between ``num`` surface samples of the predicted meshes and the ground-truth
    points (utils.py:393-438): ``(dist_1 + dist_2) * 3000`` [, f_score].  With ``group`` the batch is
    taken to be sharded over that process group and the means are global (one all-reduce of two
    scalars).  ``sample_draws`` = (choices, r1, r2) replaces the sampler's random draws (parity tests)."""
    pred_points = batch_sample(pred_vert, adj_info["faces"], num=num, **_draws(sample_draws))
    gt = gt_points.contiguous()
    sums, id_p, id_g, *cnt = chamfer_loss_sums(gt, pred_points, f1=f1)
    sums = all_reduce_sums(sums, group)
    ws = world_size(group)
    dist_2 = sums[0] / float(gt.shape[0] * gt.shape[1] * ws)                    # utils.py:417
    dist_1 = sums[1] / float(pred_points.shape[0] * pred_points.shape[1] * ws)  # utils.py:416
    loss = (dist_1 + dist_2) * 3000
    if f1:
        return loss, f_score_from_counts(cnt[0], num)
    return loss


# ------------------------------------------------------------------------------------------------
# point-to-surface loss
# ------------------------------------------------------------------------------------------------
class TriLossFunction(torch.autograd.Function):
    """sum over all ground-truth points of |closest point on the winning triangle - p|^2, with the
    closest point recomputed differentiably from the region option (utils.py:472-481, 506-550)."""

    @staticmethod
    def forward(ctx, gt_points, verts, faces, option, index):
        L = _lib.lib()
        p, v = gt_points.detach().contiguous(), verts.detach().contiguous()
        b, n, _ = p.shape
        dev = p.device
        with _lib.on_device(dev):
            total = torch.empty(1, dtype=torch.float32, device=dev)
            wsb = L.geom_tri_loss_workspace_bytes(b, n)
            ws = torch.empty(max(wsb // 8, 1), dtype=torch.float64, device=dev)
            _lib.check(L.geom_tri_loss_fwd(b, n, p.data_ptr(), v.shape[1], v.data_ptr(), faces.shape[0],
                                           faces.data_ptr(), option.data_ptr(), index.data_ptr(), None, None,
                                           total.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                           _lib.current_stream_ptr(dev)), "geom_tri_loss_fwd")
        ctx.save_for_backward(p, v, faces, option, index)
        return total

    @staticmethod
    def backward(ctx, grad_total):
        p, v, faces, option, index = ctx.saved_tensors
        L = _lib.lib()
        b, n, _ = p.shape
        dev = p.device
        w = grad_total.to(torch.float32).contiguous()
        with _lib.on_device(dev):
            gv = torch.empty_like(v)
            _lib.check(L.geom_tri_loss_bwd(b, n, p.data_ptr(), v.shape[1], v.data_ptr(), faces.shape[0],
                                           faces.data_ptr(), option.data_ptr(), index.data_ptr(), w.data_ptr(),
                                           gv.data_ptr(), _lib.current_stream_ptr(dev)), "geom_tri_loss_bwd")
        return None, gv, None, None, None


def point_to_surface_sum(gt_points, pred_vert, faces):
    """TriDistance search + differentiable closest-point sum -> (sum (1,), option, index)."""
    _lib.require_cuda_f32("pred_vert", pred_vert, 3)
    faces = _faces_i64(faces, pred_vert.device)
    with torch.no_grad():  # the search consumes positions only (its outputs carry no gradient)
        tri1, tri2, tri3 = (torch.index_select(pred_vert, 1, faces[:, k]) for k in range(3))  # utils.py:467-469
        _, option, index = tri_forward(gt_points, tri1, tri2, tri3)
    gt = gt_points.contiguous()
    return TriLossFunction.apply(gt, pred_vert, faces, option, index), option, index


def batch_point_to_surface(pred_vert, adj_info, gt_points, num=1000, f1=False, group=None, sample_draws=None):
    """Training loss (utils.py:441-502): Chamfer term from the surface samples to the ground truth
    plus point-to-surface term from the ground truth to the predicted triangles,
    ``(dist_1 + dist_2) * 3000`` [, f_score].  ``sample_draws`` as in :func:`batch_point_to_point`."""
    pred_points = batch_sample(pred_vert, adj_info["faces"], num=num, **_draws(sample_draws))
    gt = gt_points.contiguous()
    sums, id_p, id_g, *cnt = chamfer_loss_sums(gt, pred_points, f1=f1)
    tri_sum, _, _ = point_to_surface_sum(gt, pred_vert, adj_info["faces"])
    part = all_reduce_sums(torch.cat((sums[1:2], tri_sum)), group)
    ws = world_size(group)
    dist_1 = part[0] / float(pred_points.shape[0] * pred_points.shape[1] * ws)  # utils.py:462
    dist_2 = part[1] / float(gt.shape[0] * gt.shape[1] * ws)                    # utils.py:481, 549
    loss = (dist_1 + dist_2) * 3000
    if f1:
        return loss, f_score_from_counts(cnt[0], num)
    return loss


# ------------------------------------------------------------------------------------------------
# regularisers
# ------------------------------------------------------------------------------------------------
class EdgeLossFunction(torch.autograd.Function):
    """verts (B,V,3), faces (F,3) int64 -> scalar mean squared edge length: geom_edge_loss_fwd / _bwd."""

    @staticmethod
    def forward(ctx, verts, faces):
        L = _lib.lib()
        vc = verts.detach().contiguous()
        b, v, _ = vc.shape
        f = faces.shape[0]
        dev = vc.device
        with _lib.on_device(dev):
            acc = torch.empty(1, dtype=torch.float64, device=dev)
            _lib.check(L.geom_edge_loss_fwd(b, v, vc.data_ptr(), f, faces.data_ptr(), acc.data_ptr(),
                                            _lib.current_stream_ptr(dev)), "geom_edge_loss_fwd")
        ctx.save_for_backward(vc, faces)
        ctx.scale = 1.0 / (3.0 * b * f) if b * f else 0.0
        return (acc * ctx.scale).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        vc, faces = ctx.saved_tensors
        b, v, _ = vc.shape
        dev = vc.device
        gc = g.detach().to(torch.float32).contiguous()
        with _lib.on_device(dev):
            gv = torch.empty_like(vc)
            _lib.check(L.geom_edge_loss_bwd(b, v, vc.data_ptr(), faces.shape[0], faces.data_ptr(), gc.data_ptr(), ctx.scale,
                                            gv.data_ptr(), _lib.current_stream_ptr(dev)), "geom_edge_loss_bwd")
        return gv, None


def batch_calc_edge(verts, info):
    """Mean squared edge length over the three edges of every face (utils.py:636-651).  CUDA fp32 (B,V,3) input: one
    fused kernel each way (fp64 accumulation; equal to the torch expression to fp32 rounding); otherwise the
    reference's torch expression."""
    faces = info["faces"]
    if verts.is_cuda and verts.dtype == torch.float32 and verts.dim() == 3 and verts.shape[-1] == 3 and faces.numel() > 0:
        return EdgeLossFunction.apply(verts, _faces_i64(faces, verts.device))
    p1 = torch.index_select(verts, 1, faces[:, 0])
    p2 = torch.index_select(verts, 1, faces[:, 1])
    p3 = torch.index_select(verts, 1, faces[:, 2])
    e1, e2, e3 = p2 - p1, p3 - p1, p2 - p3
    return (torch.sum(e1 ** 2, -1).mean() + torch.sum(e2 ** 2, -1).mean() + torch.sum(e3 ** 2, -1).mean()) / 3.


_degree_cache: dict = {}


def batch_get_lap_info(positions, adj_info):
    """Uniform Laplacian coordinates (utils.py:654-662): positions minus the mean of the neighbours.
    The reference's dense ``adj_orig @ positions`` (called six times per step, GEOMetrics.py:156-160)
    becomes a CSR SpMM over the three coordinate columns."""
    from .layers import csr_matmul

    import weakref

    orig = adj_info["adj_orig"]
    key = (orig.data_ptr(), orig._version, orig.shape[0])
    hit = _degree_cache.get(key)
    # an entry is valid only while the tensor it was computed from is alive: the allocator can hand the same address
    # (and version 0) to a different adjacency of the same size
    scaler = hit[1] if hit is not None and hit[0]() is orig else None
    if scaler is None:
        if len(_degree_cache) > 16:
            _degree_cache.clear()
        scaler = (1. / (torch.sum(orig, dim=1) - 1)).view(-1, 1)
        _degree_cache[key] = (weakref.ref(orig), scaler)
    neighbor_sum = (csr_matmul(orig, positions) - positions) * scaler
    return positions - neighbor_sum


# ------------------------------------------------------------------------------------------------
# vertex feature pooling (utils.py:283-389)
# ------------------------------------------------------------------------------------------------
def batch_camera_info(param: torch.Tensor):
    """utils.py:283-312: camera matrix (B,3,3) and position (B,3) from (azimuth deg, elevation deg, distance)
    rows.  Same expressions in the same order (incl. the ``% 360.`` applied to radians), device-agnostic."""
    import math

    theta = (math.pi * param[:, 0] / 180.) % 360.
    phi = (math.pi * param[:, 1] / 180.) % 360.
    cam_y = param[:, 2] * torch.sin(phi)
    temp = param[:, 2] * torch.cos(phi)
    cam_x = temp * torch.cos(theta)
    cam_z = temp * torch.sin(theta)
    cam_pos = torch.cat((cam_x.unsqueeze(1), cam_y.unsqueeze(1), cam_z.unsqueeze(1)), dim=1)
    axis_z = cam_pos.clone()
    axis_y = torch.tensor([0., 1., 0.], dtype=param.dtype, device=param.device).unsqueeze(0).expand(axis_z.shape[0], 3)
    axis_x = torch.linalg.cross(axis_y, axis_z, dim=1)
    axis_y = torch.linalg.cross(axis_z, axis_x, dim=1)
    axis_x = axis_x / torch.sqrt(torch.sum(axis_x ** 2, dim=1)).unsqueeze(-1)
    axis_y = axis_y / torch.sqrt(torch.sum(axis_y ** 2, dim=1)).unsqueeze(-1)
    axis_z = axis_z / torch.sqrt(torch.sum(axis_z ** 2, dim=1)).unsqueeze(-1)
    cam_mat = torch.cat((axis_x.unsqueeze(1), axis_y.unsqueeze(1), axis_z.unsqueeze(1)), dim=1)
    return cam_mat, cam_pos


class ProjectFunction(torch.autograd.Function):
    """(verts (B,V,3), img_info (B,3)) -> (xs, ys): geom_project_fwd / geom_project_bwd (one launch each)."""

    @staticmethod
    def forward(ctx, verts, img_info):
        L = _lib.lib()
        vc, ic = verts.detach().contiguous(), img_info.detach().contiguous()
        b, v, _ = vc.shape
        dev = vc.device
        with _lib.on_device(dev):
            xs = torch.empty((b, v), dtype=torch.float32, device=dev)
            ys = torch.empty((b, v), dtype=torch.float32, device=dev)
            _lib.check(L.geom_project_fwd(b, v, vc.data_ptr(), ic.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                                          _lib.current_stream_ptr(dev)), "geom_project_fwd")
        ctx.save_for_backward(vc, ic)
        return xs, ys

    @staticmethod
    def backward(ctx, gxs, gys):
        L = _lib.lib()
        vc, ic = ctx.saved_tensors
        b, v, _ = vc.shape
        dev = vc.device
        gxs = gxs.contiguous() if gxs is not None else None
        gys = gys.contiguous() if gys is not None else None
        with _lib.on_device(dev):
            gv = torch.empty_like(vc)
            _lib.check(L.geom_project_bwd(b, v, vc.data_ptr(), ic.data_ptr(), gxs.data_ptr() if gxs is not None else None,
                                          gys.data_ptr() if gys is not None else None, gv.data_ptr(),
                                          _lib.current_stream_ptr(dev)), "geom_project_bwd")
        return gv, None


FUSED_PROJECTION = True  # project_vertices on CUDA fp32 inputs: one kernel instead of ~30 (B,3)/(B,V) torch ops


def project_vertices(verts_pos: torch.Tensor, img_info: torch.Tensor):
    """utils.py:319-330: image-plane coordinates of the vertices as fractions of the image size -> (xs, ys), (B,V).
    CUDA fp32 inputs take the fused kernel (equal to the torch expression below to fp32 rounding); anything else,
    or an ``img_info`` that requires grad, evaluates the reference's expression with torch ops."""
    if (FUSED_PROJECTION and verts_pos.is_cuda and img_info.is_cuda and verts_pos.dtype == torch.float32
            and img_info.dtype == torch.float32 and verts_pos.dim() == 3 and verts_pos.shape[-1] == 3
            and tuple(img_info.shape) == (verts_pos.shape[0], 3) and not img_info.requires_grad):
        return ProjectFunction.apply(verts_pos, img_info)
    cam_mat, cam_pos = batch_camera_info(img_info)
    pt = torch.matmul((verts_pos * .57) - cam_pos.unsqueeze(1), cam_mat.permute(0, 2, 1))
    x, y, z = pt[:, :, 0], pt[:, :, 1], pt[:, :, 2]
    f = 248
    h = (-y) / (-z) * f + 224 / 2.0
    w = x / (-z) * f + 224 / 2.0
    return h / 223., w / 223.


class PoolFunction(torch.autograd.Function):
    """(xs, ys, *maps) -> (B, V, sum C): geom_pool_fwd / geom_pool_bwd."""

    @staticmethod
    def forward(ctx, xs, ys, out_buf, col0, *maps):
        import ctypes as C

        L = _lib.lib()
        xs, ys = xs.detach().contiguous(), ys.detach().contiguous()
        maps = tuple(m.detach().contiguous() for m in maps)
        b, v = xs.shape
        dev = xs.device
        chans = [int(m.shape[1]) for m in maps]
        dims = [int(m.shape[-1]) for m in maps]
        total = sum(chans)
        with _lib.on_device(dev):
            out = out_buf if out_buf is not None else torch.empty((b, v, total), dtype=torch.float32, device=dev)
            ptrs = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
            ch = (C.c_int * len(maps))(*chans)
            dm = (C.c_int * len(maps))(*dims)
            _lib.check(L.geom_pool_fwd(b, v, len(maps), ptrs, ch, dm, xs.data_ptr(), ys.data_ptr(), out.data_ptr(),
                                       out.shape[2], int(col0), _lib.current_stream_ptr(dev)), "geom_pool_fwd")
        ctx.save_for_backward(xs, ys, *maps)
        ctx.meta = (chans, dims, int(col0), out.shape[2])
        if out_buf is not None:
            ctx.mark_dirty(out_buf)
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C

        L = _lib.lib()
        xs, ys, *maps = ctx.saved_tensors
        chans, dims, col0, ld = ctx.meta
        b, v = xs.shape
        dev = xs.device
        dout = dout.contiguous()
        want_maps = [ctx.needs_input_grad[4 + k] for k in range(len(maps))]
        with _lib.on_device(dev):
            gmaps = [torch.zeros_like(m) if w else None for m, w in zip(maps, want_maps)]
            gxs = torch.empty_like(xs) if ctx.needs_input_grad[0] else None
            gys = torch.empty_like(ys) if ctx.needs_input_grad[1] else None
            ptrs = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
            gptrs = (C.c_void_p * len(maps))(*[gm.data_ptr() if gm is not None else None for gm in gmaps])
            ch = (C.c_int * len(maps))(*chans)
            dm = (C.c_int * len(maps))(*dims)
            _lib.check(L.geom_pool_bwd(b, v, len(maps), ptrs, gptrs, ch, dm, xs.data_ptr(), ys.data_ptr(), dout.data_ptr(),
                                       ld, col0, gxs.data_ptr() if gxs is not None else None,
                                       gys.data_ptr() if gys is not None else None, _lib.current_stream_ptr(dev)), "geom_pool_bwd")
        gbuf = None
        if ctx.needs_input_grad[2]:
            # the pre-filled columns of the caller's buffer (e.g. the vertex positions next to the pooled features) pass
            # their gradient straight through; the columns this op overwrote carry none
            gbuf = dout.clone()
            gbuf[:, :, col0:col0 + sum(chans)] = 0
        return (gxs, gys, gbuf, None, *gmaps)


def pool_features(blocks, xs: torch.Tensor, ys: torch.Tensor, out: torch.Tensor = None, col0: int = 0) -> torch.Tensor:
    """Bilinear gather of every feature map in ``blocks`` (each (B,C,dim,dim)) at the projected coordinates.
    ``out`` (B,V,ld) / ``col0``: gather straight into columns [col0, col0 + sum C) of an existing buffer (e.g. the
    deformation block's input, next to the vertex positions) instead of allocating the result."""
    _lib.require_cuda_f32("xs", xs)
    _lib.require_cuda_f32("ys", ys)
    if xs.shape != ys.shape or xs.dim() != 2:
        raise ValueError(f"xs / ys must be (B,V), got {tuple(xs.shape)} and {tuple(ys.shape)}")
    blocks = list(blocks)
    if len(blocks) > 8:
        raise ValueError("at most 8 feature maps per call")
    for i, blk in enumerate(blocks):
        _lib.require_cuda_f32(f"blocks[{i}]", blk)
        if blk.dim() != 4 or blk.shape[0] != xs.shape[0] or blk.shape[2] != blk.shape[3]:
            raise ValueError(f"blocks[{i}] must be (B,C,dim,dim) with B={xs.shape[0]}, got {tuple(blk.shape)}")
    if out is not None:
        _lib.require_cuda_f32("out", out)
        total = sum(int(blk.shape[1]) for blk in blocks)
        if out.dim() != 3 or tuple(out.shape[:2]) != tuple(xs.shape) or not out.is_contiguous() or col0 < 0 \
                or col0 + total > out.shape[2]:
            raise ValueError(f"out must be a contiguous (B,V,ld) buffer with ld >= col0 + {total}, got {tuple(out.shape)}")
    return PoolFunction.apply(xs, ys, out, int(col0), *blocks)


def batched_pooling(blocks, verts_pos: torch.Tensor, img_info: torch.Tensor) -> torch.Tensor:
    """utils.py:316-389, same signature: per-vertex features (B,V,sum C) gathered from the CNN feature maps at the
    vertices' image positions.  The projection is the reference's handful of (B,V) elementwise ops; the gather
    (4 index_selects, permutes, weights and the cat per map in the reference) is one kernel, differentiable with
    respect to the maps and the vertex positions."""
    _lib.require_cuda_f32("verts_pos", verts_pos, 3)
    xs, ys = project_vertices(verts_pos, img_info)
    return pool_features(blocks, xs, ys)
