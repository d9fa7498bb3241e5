#!/usr/bin/env python
"""Round-2 groundwork (CPU only, TEST/ANALYSIS INFRASTRUCTURE: uses oracle/): a numpy model of the planned tensor-core
FILTERED TriDistance search, checked bit for bit against the brute-force oracle (tri_distance.cu semantics incl. the
512-tile structure, tail skip, NaN poisoning by a tile's first triangle, lowest-index ties).

Filter: approximate squared centroid distances dt (the quantity a bf16 tcgen05 contraction would deliver; modelled here as
exact + noise <= eps), a per-triangle lower bound  LB_j = max(0, sqrt(max(dt_j - eps, 0)) - r_j)^2 * (1 - 1e-4)  on the
reference's COMPUTED distance (tools/tri_filter_study.py: never violated), an upper bound U = exact computed distance of
the live triangle with the smallest dt, and exact evaluation (oracle arithmetic) of  {j : LB_j <= U}  plus the first
triangle of every 512-tile (a NaN there poisons the tile for that point: its triangles are dropped, as in the reference)
plus every (near-)degenerate triangle (their computed distance obeys no geometric bound).  Reports whether (dist, option, index) equal the brute force for every point and how many triangles per
point survive the filter (= the exact work left for the CUDA-core arithmetic).
    python tools/tri_filter_proto.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

TILE = 512


def exact_pairs(orc, p, a, b, c):
    """Computed distance / option of independent (point, triangle) pairs through the oracle's per-pair arithmetic."""
    d, opt, _ = orc.tri_fwd(p[:, None, :], a[:, None, :], b[:, None, :], c[:, None, :], exact=True)
    return d[:, 0], opt[:, 0]


def filtered_search(orc, pts, A, B, C, eps_scale=2.5e-4, rng=None):
    n, m = pts.shape[0], A.shape[0]
    cen = (A.astype(np.float64) + B + C) / 3
    r = np.max([np.linalg.norm(X - cen, axis=1) for X in (A, B, C)], axis=0)
    rmax = float(max(np.abs(pts).max(), np.abs(A).max(), np.abs(B).max(), np.abs(C).max()))
    eps = eps_scale * rmax * rmax
    dt = ((pts[:, None, :].astype(np.float64) - cen[None]) ** 2).sum(-1)
    if rng is not None:
        dt = dt + rng.uniform(-eps, eps, dt.shape)           # the tensor-core approximation error
    scanned = np.zeros(m, bool)                                # tail skip: last (tile length mod 4) of every tile
    firsts = []
    for k2 in range(0, m, TILE):
        end_k = min(m, k2 + TILE) - k2
        scanned[k2:k2 + end_k - (end_k & 3)] = True
        if end_k - (end_k & 3) > 0:
            firsts.append(k2)
    firsts = np.array(firsts, int)
    lb = np.maximum(0.0, np.sqrt(np.maximum(dt - eps, 0.0)) - r[None]) ** 2 * (1 - 1e-4) - 1e-12
    lb[:, ~scanned] = np.inf
    # (near-)degenerate triangles: their computed distance obeys no geometric bound -> always evaluated
    nrm2 = (np.cross(B.astype(np.float64) - A, C.astype(np.float64) - A) ** 2).sum(1)
    e2 = ((B.astype(np.float64) - A) ** 2).sum(1) * ((C.astype(np.float64) - A) ** 2).sum(1)
    suspect = scanned & ~(nrm2 > 1e-10 * e2)
    # tiles poisoned for a point (their first triangle evaluates to NaN) are ignored by the reference as a whole
    fd = np.stack([exact_pairs(orc, pts, np.repeat(A[[k]], n, 0), np.repeat(B[[k]], n, 0), np.repeat(C[[k]], n, 0))[0] for k in firsts], 1)
    tile_of = np.arange(m) // TILE
    poisoned = np.zeros((n, (m + TILE - 1) // TILE), bool)
    poisoned[:, firsts // TILE] = np.isnan(fd)
    dead = poisoned[:, tile_of]                                   # (n, m): triangle lies in a poisoned tile
    lb = np.where(dead, np.inf, lb)
    # upper bound: exact computed distance of the live scanned triangle with the smallest dt (inf if not finite)
    j0 = np.where(scanned[None] & ~dead, dt, np.inf).argmin(1)
    u, _ = exact_pairs(orc, pts, A[j0], B[j0], C[j0])
    u = np.where(np.isfinite(u) & ~dead[np.arange(n), j0], u.astype(np.float64), np.inf)
    surv = (lb <= u[:, None]) | (suspect[None] & ~dead)
    surv[:, firsts] = True
    pi, ji = np.nonzero(surv)
    d, opt = exact_pairs(orc, pts[pi], A[ji], B[ji], C[ji])
    # the reference's tile loop over the evaluated subset (skipped triangles could not have been smaller)
    out_d, out_o, out_i = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    start = np.searchsorted(pi, np.arange(n + 1))
    first_set = set(firsts.tolist())
    for i in range(n):
        js, ds, os_ = ji[start[i]:start[i + 1]], d[start[i]:start[i + 1]], opt[start[i]:start[i + 1]]
        res, res_o, res_i = np.float32(0), 0, 0
        pos = 0
        for k2 in range(0, m, TILE):
            best, bo, bi = np.float32(10000.0), 0, 0
            while pos < len(js) and js[pos] < k2 + TILE:
                if js[pos] in first_set and js[pos] == k2 or ds[pos] < best:
                    best, bo, bi = ds[pos], os_[pos], js[pos]
                pos += 1
            if k2 == 0 or res > best:
                res, res_o, res_i = best, bo, bi
        out_d[i], out_o[i], out_i[i] = res, res_o, res_i
    return out_d, out_o, out_i, surv.sum(1)


def uv_sphere(nseg, nring, radius=0.45):
    v = [(0, radius, 0)]
    for i in range(1, nring + 1):
        th = np.pi * i / (nring + 1)
        for j in range(nseg):
            ph = 2 * np.pi * j / nseg
            v.append((radius * np.sin(th) * np.cos(ph), radius * np.cos(th), radius * np.sin(th) * np.sin(ph)))
    v.append((0, -radius, 0))
    v = np.array(v, np.float32)
    f = []
    for j in range(nseg):
        f.append((0, 1 + j, 1 + (j + 1) % nseg))
    for i in range(nring - 1):
        for j in range(nseg):
            a, b = 1 + i * nseg + j, 1 + i * nseg + (j + 1) % nseg
            c, d = a + nseg, b + nseg
            f += [(a, c, b), (b, c, d)]
    last = len(v) - 1
    for j in range(nseg):
        f.append((last, 1 + (nring - 1) * nseg + (j + 1) % nseg, 1 + (nring - 1) * nseg + j))
    return v, np.array(f, int)


def main():
    orc = Oracle()
    rng = np.random.default_rng(41)
    cases = []
    v, f = uv_sphere(50, 50)
    v = v + 0.01 * rng.standard_normal(v.shape).astype(np.float32)
    cases.append(("C3-like mesh: UV sphere 5000 faces, random points in the cube", (rng.random((1500, 3)) - 0.5).astype(np.float32), v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]))
    near = (v[rng.integers(0, len(v), 1500)] * (1 + 0.02 * rng.standard_normal((1500, 1)))).astype(np.float32)
    cases.append(("same mesh, points near the surface (training regime)", near, v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]))
    a = (rng.random((4999, 3)) - 0.5).astype(np.float32)
    cases.append(("4999 random small triangles (tail skip 3)", (rng.random((800, 3)) - 0.5).astype(np.float32), a,
                  a + 0.05 * rng.standard_normal(a.shape).astype(np.float32), a + 0.05 * rng.standard_normal(a.shape).astype(np.float32)))
    A, B, C = v[f[:1030, 0]].copy(), v[f[:1030, 1]].copy(), v[f[:1030, 2]].copy()
    B[512] = A[512]                      # degenerate first triangle of tile 1: NaN poisons that tile
    C[700] = B[700]                      # degenerate inside a tile: its NaN is simply never smaller
    cases.append(("1030 faces, degenerate triangles at 512 (tile poison) and 700", (rng.random((800, 3)) - 0.5).astype(np.float32), A, B, C))
    A2, B2, C2 = A.copy(), B.copy(), C.copy()
    B2[0] = A2[0]
    cases.append(("1030 faces, degenerate triangle 0 (poisons tile 0)", (rng.random((300, 3)) - 0.5).astype(np.float32), A2, B2, C2))
    ok_all = True
    for name, pts, A, B, C in cases:
        wd, wo, wi = orc.tri_fwd(pts[None], A[None], B[None], C[None])
        gd, go, gi, ns = filtered_search(orc, pts, A, B, C, rng=rng)
        same = (np.array_equal(gd.view(np.int32), wd[0].view(np.int32)) | (np.isnan(gd) & np.isnan(wd[0])).all()) \
            and np.array_equal(gi, wi[0]) and np.array_equal(go, wo[0])
        bits = ((gd.view(np.int32) == wd[0].view(np.int32)) | (np.isnan(gd) & np.isnan(wd[0]))) & (gi == wi[0]) & (go == wo[0])
        ok_all &= bool(bits.all())
        print(f"{name}: identical {int(bits.sum())}/{len(bits)} points; survivors per point mean {ns.mean():.1f} median "
              f"{np.median(ns):.0f} max {ns.max()} of {A.shape[0]} triangles", flush=True)
    print("ALL IDENTICAL" if ok_all else "MISMATCH")


if __name__ == "__main__":
    main()
