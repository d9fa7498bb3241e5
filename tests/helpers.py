"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from geometrics_b200 import synth  # noqa: E402

# (kind, B, n, m, seed): covers config 1, ragged sizes, the reference's tail-skip cases
# ((m mod 512) mod 4 != 0: 1030, 2466, 999), the "empty final tile" cases (513, 10001, 3, 1),
# exact-tie data (lattice, duplicates) and tiny inputs.
CHAMFER_CASES = [
    ("random", 1, 1024, 1024, 41),
    ("random", 2, 777, 1030, 1),
    ("random", 2, 513, 700, 2),
    ("random", 1, 3, 5, 3),
    ("random", 1, 1, 1, 4),
    ("random", 1, 4, 4, 5),
    ("random", 3, 2466, 3000, 6),
    ("random", 2, 1500, 999, 7),
    ("random", 1, 300, 10001, 8),
    ("random", 2, 256, 8, 9),
    ("random", 1, 2, 1025, 10),
    ("lattice", 2, 900, 1100, 11),
    ("lattice", 1, 2048, 515, 12),
    ("dup", 2, 1000, 1000, 13),
    ("dup", 1, 257, 2466, 14),
    ("cluster", 2, 640, 1300, 15),
]


def chamfer_inputs(kind, b, n, m, seed):
    """CPU float32 tensors (B,n,3), (B,m,3)."""
    import torch

    if kind == "random":
        return synth.random_points(b, n, seed), synth.random_points(b, m, seed + 100)
    if kind == "lattice":
        return synth.lattice_points(b, n, 32, seed), synth.lattice_points(b, m, 32, seed + 100)
    if kind == "dup":
        return synth.duplicated_points(b, n, 300, seed), synth.duplicated_points(b, m, 250, seed + 100)
    if kind == "cluster":
        g = torch.Generator().manual_seed(seed)
        c = torch.rand(b, 1, 3, generator=g) - 0.5
        return (c + 1e-3 * torch.randn(b, n, 3, generator=g)).contiguous(), \
               (c + 1e-3 * torch.randn(b, m, 3, generator=g)).contiguous()
    raise ValueError(kind)


TRI_CASES = [
    # (kind, B, n, m, seed)
    ("sphere", 2, 1000, 960, 21),      # 482.obj-like UV sphere 32x15 -> F=960
    ("sphere50", 1, 700, 5000, 22),    # config-3 mesh
    ("random", 2, 513, 1030, 23),      # tail skip (1030 % 512 = 6 -> 2 skipped)
    ("random", 1, 300, 515, 24),       # final tile holds 3 triangles -> empty scan
    ("random", 1, 64, 3, 25),          # m < 4 -> (10000, 0, 0)
    ("random", 1, 1, 4, 26),
    ("degenerate", 2, 400, 700, 27),   # zero-area / zero-edge faces -> NaN distances
    ("degenerate0", 1, 200, 1100, 28), # degenerate face at position 0 of a 512-tile (poisoning)
    ("onmesh", 2, 600, 960, 29),       # points ON the surface: shared-edge/vertex near-ties
    ("big", 1, 257, 2466, 30),
]


def tri_inputs(kind, b, n, m, seed):
    """CPU tensors: points (B,n,3), tri1/2/3 (B,m,3)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(b, n, 3, generator=g) - 0.5
    if kind in ("sphere", "sphere50", "onmesh"):
        seg, rings = (32, 15) if kind != "sphere50" else (50, 50)
        v, f = synth.uv_sphere(seg, rings)
        assert f.shape[0] == m, (f.shape, m)
        vb = synth.mesh_batch(v, b, 0.01, seed)
        t1, t2, t3 = synth.face_vertices(vb, f)
        if kind == "onmesh":
            # barycentric samples on random faces, plus exact vertices and edge midpoints
            fi = torch.randint(0, m, (b, n), generator=g)
            w = torch.rand(b, n, 3, generator=g)
            w[:, : n // 4] = torch.tensor([1.0, 0.0, 0.0])
            w[:, n // 4: n // 2] = torch.tensor([0.5, 0.5, 0.0])
            w = w / w.sum(-1, keepdim=True)
            gat = lambda t: torch.gather(t, 1, fi.unsqueeze(-1).expand(b, n, 3))  # noqa: E731
            pts = w[..., 0:1] * gat(t1) + w[..., 1:2] * gat(t2) + w[..., 2:3] * gat(t3)
        return pts.contiguous(), t1, t2, t3
    t1, t2, t3 = synth.random_triangles(b, m, 0.05, seed)
    if kind.startswith("degenerate"):
        t2 = t2.clone()
        t3 = t3.clone()
        pos = [0, 512] if kind == "degenerate0" else [5, 100, 513, 600]
        for k, p in enumerate(pos):
            if p < m:
                if k % 2 == 0 or kind == "degenerate0":
                    t2[:, p] = t1[:, p]  # zero-length edge AB
                else:
                    t3[:, p] = 2 * t2[:, p] - t1[:, p]  # collinear: zero normal
    return pts.contiguous(), t1.contiguous(), t2.contiguous(), t3.contiguous()


def assert_same_f32(a, b, what=""):
    """Bitwise equality of float arrays, NaN == NaN."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} float values differ"
