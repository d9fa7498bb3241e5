"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything is generated on the CPU with an explicit ``torch.Generator`` so that the same seed gives
the same tensors on every box; callers move them to the GPU.  Seed 41 is the reference's default
(``GEOMetrics.py:14``); points live in the [-0.5, 0.5) box like the reference's own dummy points
(``utils.py:48``) and its 482.obj sphere.
"""
from __future__ import annotations

import math

import torch

DEFAULT_SEED = 41


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def random_points(b: int, n: int, seed: int = DEFAULT_SEED) -> torch.Tensor:
    return torch.rand(b, n, 3, generator=_gen(seed)) - 0.5


def lattice_points(b: int, n: int, cells: int = 128, seed: int = DEFAULT_SEED) -> torch.Tensor:
    """Points on a k/cells lattice: many exactly equal distances (the voxel-derived ground truth of
    ``data_prep.py:337-345`` has this structure), so the lowest-index tie rule is exercised."""
    return torch.randint(0, cells, (b, n, 3), generator=_gen(seed)).float() / cells - 0.5


def duplicated_points(b: int, n: int, unique: int, seed: int = DEFAULT_SEED) -> torch.Tensor:
    """`unique` random points repeated until n (``data_prep.py:385-386`` duplicates clouds this way)."""
    base = random_points(b, unique, seed)
    reps = (n + unique - 1) // unique
    return base.repeat(1, reps, 1)[:, :n].contiguous()


def uv_sphere(segments: int = 50, rings: int = 50, radius: float = 0.45):
    """UV sphere like the reference's 482.obj (32 segments x 15 rings): V = segments*rings + 2,
    F = 2*segments*rings.  (50, 50) gives V=2502, F=5000 = BASELINE config 3."""
    verts = [(0.0, radius, 0.0)]
    for r in range(1, rings + 1):
        phi = math.pi * r / (rings + 1)
        for s in range(segments):
            th = 2.0 * math.pi * s / segments
            verts.append((radius * math.sin(phi) * math.cos(th), radius * math.cos(phi),
                          radius * math.sin(phi) * math.sin(th)))
    verts.append((0.0, -radius, 0.0))
    south = len(verts) - 1
    faces = []
    ring = lambda r, s: 1 + (r - 1) * segments + (s % segments)  # noqa: E731
    for s in range(segments):
        faces.append((0, ring(1, s + 1), ring(1, s)))
        faces.append((south, ring(rings, s), ring(rings, s + 1)))
    for r in range(1, rings):
        for s in range(segments):
            a, b_, c, d = ring(r, s), ring(r, s + 1), ring(r + 1, s), ring(r + 1, s + 1)
            faces.append((a, b_, d))
            faces.append((a, d, c))
    return torch.tensor(verts, dtype=torch.float32), torch.tensor(faces, dtype=torch.int64)


def icosphere(subdiv: int = 4, radius: float = 0.45):
    """Icosphere: subdiv=4 gives V=2562, F=5120 (BASELINE config 4)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [tuple(c / math.sqrt(1 + t * t) for c in p) for p in v]
    for _ in range(subdiv):
        cache = {}

        def mid(a, b_):
            key = (min(a, b_), max(a, b_))
            if key not in cache:
                p = tuple((verts[a][i] + verts[b_][i]) / 2.0 for i in range(3))
                ln = math.sqrt(sum(c * c for c in p))
                verts.append(tuple(c / ln for c in p))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b_, c in f:
            ab, bc, ca = mid(a, b_), mid(b_, c), mid(c, a)
            nf += [(a, ab, ca), (b_, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return torch.tensor(verts, dtype=torch.float32) * radius, torch.tensor(f, dtype=torch.int64)


def mesh_batch(verts: torch.Tensor, b: int, jitter: float = 0.01, seed: int = DEFAULT_SEED) -> torch.Tensor:
    """(V,3) -> (B,V,3) with per-item N(0, jitter^2) vertex noise."""
    return verts.unsqueeze(0) + jitter * torch.randn(b, verts.shape[0], 3, generator=_gen(seed + 1))


def face_vertices(verts_b: torch.Tensor, faces: torch.Tensor):
    """tri1, tri2, tri3 exactly as the caller builds them (``utils.py:467-469``)."""
    return tuple(torch.index_select(verts_b, 1, faces[:, k]).contiguous() for k in range(3))


def random_triangles(b: int, m: int, spread: float = 0.05, seed: int = DEFAULT_SEED):
    g = _gen(seed + 2)
    a = torch.rand(b, m, 3, generator=g) - 0.5
    return a, a + spread * torch.randn(b, m, 3, generator=g), a + spread * torch.randn(b, m, 3, generator=g)
