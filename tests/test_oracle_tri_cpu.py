"""CPU tests of the TriDistance oracle: known answers for each region option, the option-6 quirk,
tail/NaN semantics, and agreement of the minimum distance with an independent
float64 point-triangle distance (Ericson's closest-point algorithm)."""
import numpy as np
import pytest

from helpers import TRI_CASES, tri_inputs

A = np.array([0, 0, 0], np.float32)
B = np.array([1, 0, 0], np.float32)
C = np.array([0, 1, 0], np.float32)


@pytest.mark.parametrize("p,opt,d", [
    ((0.25, 0.25, 0.5), 0, 0.25),     # above the interior
    ((-1.0, -1.0, 0.0), 1, 2.0),      # vertex A
    ((2.0, -0.5, 0.0), 2, 1.25),      # vertex B
    ((-0.5, 2.0, 0.0), 3, 1.25),      # vertex C
    ((0.5, -1.0, 0.0), 4, 1.0),       # edge AB
    ((1.0, 1.0, 0.0), 5, 0.5),        # edge BC
])
def test_region_known_answers(oracle, p, opt, d):
    dd, oo = oracle.tri_pair(np.array(p, np.float32), A, B, C)
    assert oo == opt
    assert dd == pytest.approx(d, rel=1e-6)


def test_option6_direction_quirk(oracle):
    """tri_distance.cu:180: the CA-edge point is C + uca*AB (not C + uca*CA)."""
    p = np.array((-1.0, 0.5, 0.0), np.float32)
    dd, oo = oracle.tri_pair(p, A, B, C)
    assert oo == 6
    # uca = dot(p-C, CA)/|CA|^2 = 0.5 ; buggy point = C + 0.5*AB = (0.5, 1, 0); true closest = (0, 0.5, 0)
    assert dd == pytest.approx((1.5) ** 2 + 0.25, rel=1e-6)


def closest_point_f64(p, a, b, c):
    """Ericson, Real-Time Collision Detection 5.1.5 (independent of the reference's formulation)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return a
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return b
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        return a + ab * (d1 / (d1 - d3))
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return c
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        return a + ac * (d2 / (d2 - d6))
    va = d3 * d6 - d5 * d4
    if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
        return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))
    den = 1.0 / (va + vb + vc)
    return a + ab * (vb * den) + ac * (vc * den)


def test_pair_distance_vs_independent_f64(oracle):
    rng = np.random.default_rng(3)
    bad = 0
    for _ in range(3000):
        a, b, c = (rng.random(3) - 0.5 for _ in range(3))
        p = rng.random(3) - 0.5
        dd, oo = oracle.tri_pair(*(v.astype(np.float32) for v in (p, a, b, c)))
        if oo == 6:
            continue  # the reference's CA-edge point is deliberately wrong (quirk above)
        a, b, c, p = (v.astype(np.float32).astype(np.float64) for v in (a, b, c, p))
        q = closest_point_f64(p, a, b, c)
        ref = ((p - q) ** 2).sum()
        # The reference's region logic is not an exact Voronoi decomposition for obtuse triangles,
        # so it may return a slightly larger distance; it must never return a smaller one.
        assert dd >= ref * (1 - 1e-4) - 1e-9
        bad += dd > ref * (1 + 1e-3) + 1e-9
    assert bad < 600


@pytest.mark.parametrize("case", TRI_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}x{c[3]}")
def test_tri_fwd_structure(oracle, case):
    kind, b, n, m, seed = case
    pts, t1, t2, t3 = (t.numpy() for t in tri_inputs(kind, b, n, m, seed))
    dist, point, index = oracle.tri_fwd(pts, t1, t2, t3)
    assert dist.shape == (b, n) and point.dtype == np.int32
    assert ((point >= 0) & (point <= 6)).all()
    assert ((index >= 0) & (index < max(m, 1))).all()
    m_scan = m - ((m % 512) & 3)
    if m < 4:
        assert (dist == 10000).all() and (index == 0).all() and (point == 0).all()
        return
    assert (index < m_scan).all()  # skipped tail triangles can never win
    # the reported distance/option is the pair evaluation of the reported triangle
    for bi in range(b):
        for j in range(0, n, max(1, n // 50)):
            k = index[bi, j]
            d, o = oracle.tri_pair(pts[bi, j], t1[bi, k], t2[bi, k], t3[bi, k])
            if np.isnan(d):
                assert np.isnan(dist[bi, j])
            else:
                assert d == dist[bi, j] and o == point[bi, j]
    if kind in ("sphere", "random", "big"):
        de, pe, ie = oracle.tri_fwd(pts, t1, t2, t3, exact=True)
        assert (de <= dist).all()
        if m_scan == m:
            assert (ie == index).all() and (pe == point).all()


def test_nan_tile_poisoning(oracle):
    """A degenerate triangle at position 0 of a 512-tile makes that tile's best NaN
    (tri_distance.cu:194 with k==0), which then never wins the merge (:202) -- or, for tile 0,
    sticks forever."""
    pts, t1, t2, t3 = (t.numpy() for t in tri_inputs("degenerate0", 1, 200, 1100, 28))
    dist, point, index = oracle.tri_fwd(pts, t1, t2, t3)
    d0 = np.array([oracle.tri_pair(pts[0, j], t1[0, 0], t2[0, 0], t3[0, 0])[0] for j in range(pts.shape[1])])
    poisoned = np.isnan(d0)
    assert poisoned.any() and not poisoned.all()
    assert np.isnan(dist[0][poisoned]).all() and (index[0][poisoned] == 0).all()
    assert not np.isnan(dist[0][~poisoned]).any()
    # move the degenerate triangle off position 0: it alone is ignored (NaN never wins '<')
    t1b, t2b, t3b = t1.copy(), t2.copy(), t3.copy()
    for t in (t1b, t2b, t3b):
        t[:, [0, 1]] = t[:, [1, 0]]
    dist_b, point_b, index_b = oracle.tri_fwd(pts, t1b, t2b, t3b)
    assert not np.isnan(dist_b).any()
    # tile 1 starts at face 512, also degenerate: rows it poisons ignore faces 512..1023
    d512 = np.array([oracle.tri_pair(pts[0, j], t1[0, 512], t2[0, 512], t3[0, 512])[0] for j in range(pts.shape[1])])
    p1 = np.isnan(d512)
    assert p1.any()
    assert ((index_b[0][p1] < 512) | (index_b[0][p1] >= 1024)).all()
