#!/usr/bin/env python
"""Round-2 groundwork (CPU only, TEST/ANALYSIS INFRASTRUCTURE: uses oracle/): how far below the TRUE point-triangle
distance can the reference TriDistance kernel's COMPUTED distance fall?  A tensor-core filter for TriDistance (like the
Chamfer one) needs a lower bound on the computed value; the reference's region logic and its option-6 bug
(tri_distance.cu:153-189, :180) make the computed value differ from the true distance.  For random triangles of several
shape families and random points this script evaluates the oracle's per-pair distance/option and reports, per family,
violations of candidate lower bounds:
   plane    : squared distance to the triangle's plane
   sphere   : max(0, |p-c| - r)^2       c = centroid, r = max vertex distance from c
   sphere+AB: max(0, |p-c| - r - |AB|)^2  (room for the option-6 point C + AB*u)
    python tools/tri_filter_study.py [triangles per family] [points per triangle]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402


def true_dist2(p, a, b, c):
    """Ericson, Real-Time Collision Detection 5.1.5, vectorised in float64: p (T,P,3); a,b,c (T,1,3)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = p - b
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = p - c
    d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    q = np.empty_like(p)
    done = np.zeros(p.shape[:2], bool)

    def put(mask, val):
        m = mask & ~done
        q[m] = np.broadcast_to(val, p.shape)[m]
        done[m] = True

    with np.errstate(divide="ignore", invalid="ignore"):
        put((d1 <= 0) & (d2 <= 0), a)
        put((d3 >= 0) & (d4 <= d3), b)
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[..., None] * ab)
        put((d6 >= 0) & (d5 <= d6), c)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[..., None] * ac)
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + w[..., None] * (c - b))
        den = 1.0 / (va + vb + vc)
        put(np.ones_like(done), a + ab * (vb * den)[..., None] + ac * (vc * den)[..., None])
    return ((p - q) ** 2).sum(-1)


def families(rng, t):
    a = rng.random((t, 3)) - 0.5
    yield "small (mesh-like, edge ~0.05)", a, a + 0.05 * rng.standard_normal((t, 3)), a + 0.05 * rng.standard_normal((t, 3))
    yield "large (edge ~0.5)", a, a + 0.5 * rng.standard_normal((t, 3)), a + 0.5 * rng.standard_normal((t, 3))
    d = rng.standard_normal((t, 3))
    yield "needle (aspect ~100)", a, a + 0.3 * d, a + 0.3 * d * rng.random((t, 1)) + 0.003 * rng.standard_normal((t, 3))
    e = rng.standard_normal((t, 3))
    yield "obtuse (angle at A > 150 deg)", a, a + 0.2 * e, a - 0.2 * e * rng.random((t, 1)) + 0.02 * rng.standard_normal((t, 3))


def main():
    t = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    pts = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    orc = Oracle()
    rng = np.random.default_rng(41)
    for name, a, b, c in families(rng, t):
        a, b, c = (x.astype(np.float32) for x in (a, b, c))
        p = (rng.random((t, pts, 3)) - 0.5).astype(np.float32)
        p[:, : pts // 4] = (a[:, None] + 0.05 * rng.standard_normal((t, pts // 4, 3))).astype(np.float32)  # near the triangle
        d, opt, _ = orc.tri_fwd(p, a[:, None], b[:, None], c[:, None], exact=True)
        a64, b64, c64, p64 = (x.astype(np.float64) for x in (a[:, None], b[:, None], c[:, None], p))
        td = true_dist2(p64, a64, b64, c64)
        n = np.cross(b64 - a64, c64 - a64)
        plane = ((p64 - a64) * n).sum(-1) ** 2 / np.maximum((n * n).sum(-1), 1e-300)
        cen = (a64 + b64 + c64) / 3
        r = np.max([np.linalg.norm(x - cen, axis=-1) for x in (a64, b64, c64)], axis=0)
        pc = np.linalg.norm(p64 - cen, axis=-1)
        sphere = np.maximum(0, pc - r) ** 2
        sphere_ab = np.maximum(0, pc - r - np.linalg.norm(b64 - a64, axis=-1)) ** 2
        fin = np.isfinite(d)
        dd = d.astype(np.float64)
        tol = 1e-5 * np.maximum(dd, td) + 1e-12
        rep = {"pairs": int(fin.sum()), "nonfinite": int((~fin).sum()),
               "computed < true": int((fin & (dd < td - tol)).sum()),
               "computed > true": int((fin & (dd > td + tol)).sum()),
               "worst computed/true": float(np.min(np.where(fin & (td > 1e-12), dd / np.maximum(td, 1e-300), 1.0))),
               "below plane": int((fin & (dd < plane - tol)).sum()),
               "below sphere": int((fin & (dd < sphere - tol)).sum()),
               "below sphere+AB": int((fin & (dd < sphere_ab - tol)).sum())}
        under = fin & (dd < td - tol)
        rep["options of computed<true"] = {int(k): int(v) for k, v in zip(*np.unique(opt[under], return_counts=True))}
        print(name, rep, flush=True)


if __name__ == "__main__":
    main()
