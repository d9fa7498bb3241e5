"""Plain-torch fp32 restatements of the reference's floating-point hot-path functions, used only by
the tests as the parity reference for the fused CUDA kernels (device-agnostic: they run on whatever
device their inputs live on).  Each function names the reference lines it follows.  TEST
INFRASTRUCTURE ONLY -- nothing under geometrics_b200/ imports this."""
import torch


def face_areas(verts, faces, normalise=True):
    """utils.py:596-602."""
    x1, x2, x3 = torch.split(torch.index_select(verts, 1, faces[:, 0]) - torch.index_select(verts, 1, faces[:, 1]), 1, dim=-1)
    y1, y2, y3 = torch.split(torch.index_select(verts, 1, faces[:, 1]) - torch.index_select(verts, 1, faces[:, 2]), 1, dim=-1)
    a = (x2 * y3 - x3 * y2) ** 2
    b = (x3 * y1 - x1 * y3) ** 2
    c = (x1 * y2 - x2 * y1) ** 2
    areas = torch.sqrt(a + b + c) / 2
    if not normalise:
        return areas.squeeze(-1)
    return areas.squeeze(-1) / torch.sum(areas, dim=1)


def batch_sample_given(verts, faces, choices, r1, r2):
    """utils.py:614-631 with the random draws passed in: choices (B*num,) int64, r1/r2 (B*num,1)."""
    b, v, _ = verts.shape
    num = choices.numel() // b
    select_faces = faces[choices].view(b, 3, num)
    face_arange = v * torch.arange(0, b, device=verts.device).unsqueeze(-1).expand(b, num)
    select_faces = (select_faces + face_arange.unsqueeze(1)).view(-1, 3)
    flat = verts.view(-1, 3)
    xs = torch.index_select(flat, 0, select_faces[:, 0])
    ys = torch.index_select(flat, 0, select_faces[:, 1])
    zs = torch.index_select(flat, 0, select_faces[:, 2])
    u = torch.sqrt(r1.view(-1, 1))
    vv = r2.view(-1, 1)
    points = (1 - u) * xs + (u * (1 - vv)) * ys + u * vv * zs
    return points.view(b, num, 3)


def chamfer_terms(gt_points, pred_points, id_p, id_g):
    """utils.py:405-417 / 454-462: -> (dist_1, dist_2) means."""
    b, n_gt, _ = gt_points.shape
    n_pr = pred_points.shape[1]
    pr = pred_points.reshape(-1, 3)
    gt = gt_points.contiguous().view(-1, 3)
    rng_p = n_pr * torch.arange(0, b, device=gt.device).unsqueeze(-1).expand(b, n_gt)
    rng_g = n_gt * torch.arange(0, b, device=gt.device).unsqueeze(-1).expand(b, n_pr)
    pred_counters = torch.index_select(pr, 0, (id_p.long() + rng_p).view(-1))
    gt_counters = torch.index_select(gt, 0, (id_g.long() + rng_g).view(-1))
    dist_1 = torch.mean(torch.sum((gt_counters - pr) ** 2, dim=1))
    dist_2 = torch.mean(torch.sum((pred_counters - gt) ** 2, dim=1))
    return dist_1, dist_2


def _project(a, d, p):
    """edge.Project, utils.py:564-569."""
    vec = p - a
    return (vec[:, 0] * d[:, 0] + vec[:, 1] * d[:, 1] + vec[:, 2] * d[:, 2]) / torch.sum(d ** 2, dim=-1)


def calc_point_to_line(p, triangles, point_options):
    """utils.py:506-550 (+ edge :553-570, Plane :572-587): mean squared distance from p (N,3) to the
    closest point selected by point_options on triangles (a, b, c) each (N,3)."""
    a, b, c = triangles
    counter_p = torch.zeros_like(p)
    d_ab, d_bc, d_ca = b - a, c - b, a - c
    uab, ubc, uca = _project(a, d_ab, p), _project(b, d_bc, p), _project(c, d_ca, p)
    tri_norm = torch.cross(a - b, a - c, dim=-1)
    direction = tri_norm / torch.sqrt(torch.sum(tri_norm ** 2, dim=-1)).unsqueeze(-1)
    v = p - a
    dist = (v[:, 0] * direction[:, 0] + v[:, 1] * direction[:, 1] + v[:, 2] * direction[:, 2]).unsqueeze(-1)
    plane_proj = p - dist * direction
    cands = {1: a, 2: b, 3: c, 4: a + uab.unsqueeze(-1) * d_ab, 5: b + ubc.unsqueeze(-1) * d_bc,
             6: c + uca.unsqueeze(-1) * d_ca, 0: plane_proj}
    for k, val in cands.items():
        cond = point_options == k
        counter_p[cond] = val[cond]
    return torch.mean(torch.sum((counter_p - p) ** 2, dim=-1))


def point_to_surface_term(gt_points, pred_vert, faces, point_options, index):
    """utils.py:467-481: gather the winning triangles and evaluate calc_point_to_line."""
    b, n, _ = gt_points.shape
    tris = [torch.index_select(pred_vert, 1, faces[:, k]) for k in range(3)]
    f = tris[0].shape[1]
    rng = f * torch.arange(0, b, device=gt_points.device).unsqueeze(-1).expand(b, n)
    flat_index = (index.long() + rng).view(-1)
    sets = [torch.index_select(t.reshape(-1, 3), 0, flat_index) for t in tris]
    return calc_point_to_line(gt_points.contiguous().view(-1, 3), sets, point_options.view(-1))


def f_score(gt_points, pred_points, id_p, id_g, num):
    """utils.py:424-436, with the per-item loop kept."""
    b = gt_points.shape[0]
    pred_counters = torch.gather(pred_points, 1, id_p.long().unsqueeze(-1).expand(-1, -1, 3))
    gt_counters = torch.gather(gt_points, 1, id_g.long().unsqueeze(-1).expand(-1, -1, 3))
    dist_to_pred = torch.sqrt(torch.sum((.57 * pred_counters - .57 * gt_points) ** 2, dim=-1))
    dist_to_gt = torch.sqrt(torch.sum((.57 * gt_counters - .57 * pred_points) ** 2, dim=-1))
    score = 0
    for i in range(b):
        recall = float(torch.where(dist_to_pred[i] <= 1e-2)[0].shape[0]) / float(num)
        precision = float(torch.where(dist_to_gt[i] <= 1e-2)[0].shape[0]) / float(num)
        score += 2 * (precision * recall) / (precision + recall + 1e-8)
    return score / b


def zeron_gcn(x, weight, bias, adj, split):
    """layers.py:107-116 / 143-152 / 34-41: support = x @ W; first out//split columns smoothed by adj."""
    support = torch.matmul(x, weight)
    k = support.shape[-1] // split
    out = torch.cat((torch.matmul(adj, support[..., :k]), support[..., k:]), dim=-1)
    return out + bias if bias is not None else out


def batch_camera_info(param):
    """utils.py:283-312 restated device-agnostically (same expressions, same order)."""
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
    return torch.cat((axis_x.unsqueeze(1), axis_y.unsqueeze(1), axis_z.unsqueeze(1)), dim=1), cam_pos


def project_vertices(verts_pos, img_info):
    """utils.py:319-330."""
    cam_mat, cam_pos = batch_camera_info(img_info)
    pt = torch.matmul((verts_pos * .57) - cam_pos.unsqueeze(1), cam_mat.permute(0, 2, 1))
    x, y, z = pt[:, :, 0], pt[:, :, 1], pt[:, :, 2]
    h = (-y) / (-z) * 248 + 224 / 2.0
    w = x / (-z) * 248 + 224 / 2.0
    return h / 223., w / 223.


def pool_features(blocks, xs, ys):
    """utils.py:331-389: the bilinear gather, with the reference's association order
    (A*C)*G + (H*D)*A + (G*E)*B + (B*F)*H and its floor == ceil -> zero-weight quirk."""
    b, v = xs.shape
    outs = []
    for blk in blocks:
        dim = blk.shape[-1]
        cx, cy = torch.clamp(xs * dim, 0, dim - 1), torch.clamp(ys * dim, 0, dim - 1)
        x1, y1, x2, y2 = torch.floor(cx), torch.floor(cy), torch.ceil(cx), torch.ceil(cy)
        wa, wb, wg, wh = x2 - cx, cx - x1, y2 - cy, cy - y1
        flat = blk.reshape(b, blk.shape[1], dim * dim)

        def take(xi, yi):
            idx = (xi.long() * dim + yi.long()).unsqueeze(1).expand(b, blk.shape[1], v)
            return torch.gather(flat, 2, idx)

        f = wa.unsqueeze(1) * take(x1, y1) * wg.unsqueeze(1) + wh.unsqueeze(1) * take(x1, y2) * wa.unsqueeze(1) \
            + wg.unsqueeze(1) * take(x2, y1) * wb.unsqueeze(1) + wb.unsqueeze(1) * take(x2, y2) * wh.unsqueeze(1)
        outs.append(f.permute(0, 2, 1))
    return torch.cat(outs, dim=2)


def batched_pooling(blocks, verts_pos, img_info):
    return pool_features(blocks, *project_vertices(verts_pos, img_info))
