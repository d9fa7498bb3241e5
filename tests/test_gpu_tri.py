"""GPU parity tests of geom_tri_fwd (through the C ABI via geometrics_b200.tri_forward): bit-exact
(dist, option, face index) against the CPU oracle and against the reference's own tri_distance.cu
compiled unmodified (oracle/_ref/libref_cuda.so), degenerate / tail / tie cases included, plus the
BASELINE config-3 size."""
import numpy as np
import pytest
import torch

from helpers import TRI_CASES, tri_inputs, assert_same_f32

pytestmark = pytest.mark.gpu
_ids = lambda c: f"{c[0]}-{c[1]}x{c[2]}x{c[3]}"  # noqa: E731


def _cmp(got, want, what):
    gd, gp, gi = got
    wd, wp, wi = want
    assert_same_f32(gd, wd, f"{what} dist")
    assert (gi == wi).all(), f"{what}: {np.count_nonzero(gi != wi)} of {wi.size} face indices differ"
    assert (gp == wp).all(), f"{what}: {np.count_nonzero(gp != wp)} of {wp.size} options differ"


@pytest.mark.parametrize("tensor", [False, True], ids=["cuda_cores", "tensor"])
@pytest.mark.parametrize("exact", [False, True], ids=["default", "exact_tail"])
@pytest.mark.parametrize("case", TRI_CASES, ids=_ids)
def test_tri_matches_oracle(oracle, cuda, case, exact, tensor):
    import geometrics_b200 as g

    pts, t1, t2, t3 = tri_inputs(*case)
    if tensor and not exact and case[3] < 4:
        pytest.skip("nothing is scanned for m < 4: the library always takes the plain path")
    got = g.tri_forward(*(t.to(cuda) for t in (pts, t1, t2, t3)), exact_tail=exact, tensor=tensor)
    torch.cuda.synchronize()
    want = oracle.tri_fwd(pts.numpy(), t1.numpy(), t2.numpy(), t3.numpy(), exact=exact)
    _cmp(tuple(t.cpu().numpy() for t in got), want, str(case))


@pytest.mark.parametrize("tensor", [False, True], ids=["cuda_cores", "tensor"])
@pytest.mark.parametrize("case", TRI_CASES, ids=_ids)
def test_tri_matches_reference_cuda(ref_cuda, cuda, case, tensor):
    import geometrics_b200 as g

    if tensor and case[3] < 4:
        pytest.skip("nothing is scanned for m < 4: the library always takes the plain path")
    dev = tuple(t.to(cuda) for t in tri_inputs(*case))
    want = ref_cuda.tri_forward_t(*dev)
    got = g.tri_forward(*dev, tensor=tensor)
    torch.cuda.synchronize()
    _cmp(tuple(t.cpu().numpy() for t in got), tuple(t.cpu().numpy() for t in want), str(case))


def test_oracle_matches_reference_cuda(oracle, ref_cuda, cuda):
    """Pins the CPU oracle against the reference GPU kernel (no CPU path, no goldens)."""
    for case in TRI_CASES:
        pts, t1, t2, t3 = tri_inputs(*case)
        want = ref_cuda.tri_forward_t(*(t.to(cuda) for t in (pts, t1, t2, t3)))
        got = oracle.tri_fwd(pts.numpy(), t1.numpy(), t2.numpy(), t3.numpy())
        _cmp(got, tuple(t.cpu().numpy() for t in want), f"oracle {case}")


@pytest.mark.parametrize("tune", ["1,64", "1,512", "2,128", "2,512", "4,256", "4,512"])
def test_every_kernel_variant(oracle, cuda, tune):
    """Points-per-lane and chunk-size variants of the CUDA-core kernel (normally picked from the problem size) all
    give the reference's answer, on cases that exercise tile poisoning, the tail skip and multi-tile merges."""
    import geometrics_b200 as g
    from geometrics_b200 import _lib

    pr, chunk = (float(x) for x in tune.split(","))
    g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_PR, pr)
    g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_CHUNK, chunk)
    try:
        for case in [("degenerate0", 1, 200, 1100, 28), ("big", 1, 257, 2466, 30), ("onmesh", 2, 600, 960, 29),
                     ("random", 2, 513, 1030, 23), ("random", 1, 300, 515, 24)]:
            pts, t1, t2, t3 = tri_inputs(*case)
            for exact in (False, True):
                got = g.tri_forward(*(t.to(cuda) for t in (pts, t1, t2, t3)), exact_tail=exact, tensor=False)
                want = oracle.tri_fwd(pts.numpy(), t1.numpy(), t2.numpy(), t3.numpy(), exact=exact)
                _cmp(tuple(t.cpu().numpy() for t in got), want, f"{tune} {case} exact={exact}")
    finally:
        g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_PR, 0.0)
        g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_CHUNK, 0.0)


def test_module_surface(cuda):
    """TriDistance()(xyz1, tri1, tri2, tri3) -> (dist f32, point i32, index i32), all (B,n); the
    caller's gathers (utils.py:467-469) produce tensors with requires_grad."""
    from tri_distance import TriDistance
    from geometrics_b200 import synth

    v, f = synth.uv_sphere(32, 15)
    vb = synth.mesh_batch(v, 2).to(cuda).requires_grad_(True)
    f = f.to(cuda)
    t1, t2, t3 = (torch.index_select(vb, 1, f[:, k]) for k in range(3))
    gt = torch.rand(2, 3000, 3, device=cuda) - 0.5
    dist, point, index = TriDistance()(gt, t1, t2, t3)
    for t, dt in ((dist, torch.float32), (point, torch.int32), (index, torch.int32)):
        assert t.shape == (2, 3000) and t.dtype == dt and t.device == gt.device and not t.requires_grad
    assert int(point.min()) >= 0 and int(point.max()) <= 6 and int(index.max()) < 960


def test_argument_errors(cuda):
    import geometrics_b200 as g

    p = torch.zeros(2, 8, 3, device=cuda)
    t = torch.zeros(2, 4, 3, device=cuda)
    with pytest.raises(ValueError):
        g.tri_forward(p, t, t, torch.zeros(2, 5, 3, device=cuda))
    with pytest.raises(ValueError):
        g.tri_forward(p, *(torch.zeros(3, 4, 3, device=cuda),) * 3)
    with pytest.raises(TypeError):
        g.tri_forward(p.half(), t, t, t)
    assert g.lib().geom_tri_fwd(2, 8, p.data_ptr(), 4, t.data_ptr(), t.data_ptr(), t.data_ptr(),
                                None, None, None, None, 0, 0, None) == -1
    i = torch.zeros(2, 8, dtype=torch.int32, device=cuda)
    assert g.lib().geom_tri_fwd(2, 8, p.data_ptr(), 4, t.data_ptr(), t.data_ptr(), t.data_ptr(),
                                None, i.data_ptr(), i.data_ptr(), None, 0, 0, None) == -2  # workspace missing
    # no faces: outputs are the zeros the reference's caller pre-allocates (tri_distance.py:22-30)
    d, pt, ix = g.tri_forward(p, *(torch.zeros(2, 0, 3, device=cuda),) * 3)
    assert (d == 0).all() and (pt == 0).all() and (ix == 0).all()


@pytest.mark.parametrize("tensor", [False, True], ids=["cuda_cores", "tensor"])
def test_full_size_config3(ref_cuda, cuda, tensor):
    """BASELINE config 3: 32 x (8192 points vs 5000 faces) on the jittered UV sphere."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    v, f = synth.uv_sphere(50, 50)
    t1, t2, t3 = (t.to(cuda) for t in synth.face_vertices(synth.mesh_batch(v, 32), f))
    pts = synth.random_points(32, 8192, 41).to(cuda)
    got = g.tri_forward(pts, t1, t2, t3, tensor=tensor)
    want = ref_cuda.tri_forward_t(pts, t1, t2, t3)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # second stress set: independent random triangles
    r1, r2, r3 = (t.to(cuda) for t in synth.random_triangles(32, 5000))
    got = g.tri_forward(pts, r1, r2, r3, tensor=tensor)
    want = ref_cuda.tri_forward_t(pts, r1, r2, r3)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # points near the surface (the training regime): shared-edge near-ties everywhere
    vb = synth.mesh_batch(v, 32)
    g_ = torch.Generator().manual_seed(7)
    near = vb[:, torch.randint(0, vb.shape[1], (8192,), generator=g_)] * (1 + 0.02 * torch.randn(32, 8192, 1, generator=g_))
    near = near.contiguous().to(cuda)
    got = g.tri_forward(near, t1, t2, t3, tensor=tensor)
    want = ref_cuda.tri_forward_t(near, t1, t2, t3)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1.0, 50.0), (1e4, 0.0), (1e6, -3e6), (1e-12, 0.0), (1e16, 0.0)])
def test_tensor_path_ranges(oracle, cuda, scale, offset):
    """Coordinate ranges: the bf16 filter margin scales with R^2 (offsets make it useless but never wrong); magnitudes
    it cannot handle (R <= 1e-15, R >= 1e15) take the in-kernel transcription of the reference loop."""
    import geometrics_b200 as g

    for case in [("sphere", 2, 700, 960, 31), ("random", 1, 300, 1030, 32), ("degenerate0", 1, 200, 1100, 33)]:
        pts, t1, t2, t3 = (t * scale + offset for t in tri_inputs(*case))
        got = g.tri_forward(*(t.to(cuda) for t in (pts, t1, t2, t3)), tensor=True)
        want = oracle.tri_fwd(pts.numpy(), t1.numpy(), t2.numpy(), t3.numpy())
        _cmp(tuple(t.cpu().numpy() for t in got), want, f"{case} x{scale}+{offset}")


def test_tensor_path_nonfinite_points(oracle, cuda):
    """NaN / Inf query points: the reference's first-candidate rule decides (tri_distance.cu:194,202)."""
    import geometrics_b200 as g

    pts, t1, t2, t3 = tri_inputs("sphere", 2, 600, 960, 34)
    pts[0, 5, 1] = float("nan")
    pts[1, 17, 0] = float("inf")
    pts[1, 300] = float("-inf")
    for tensor in (True, False):
        got = g.tri_forward(*(t.to(cuda) for t in (pts, t1, t2, t3)), tensor=tensor)
        want = oracle.tri_fwd(pts.numpy(), t1.numpy(), t2.numpy(), t3.numpy())
        _cmp(tuple(t.cpu().numpy() for t in got), want, f"nonfinite tensor={tensor}")


def test_tensor_path_margin(ref_cuda, cuda):
    """The filter stays exact with its bf16 margin shrunk 8x: the shipped margin has headroom."""
    from geometrics_b200 import synth, _lib
    import geometrics_b200 as g

    v, f = synth.uv_sphere(50, 50)
    t1, t2, t3 = (t.to(cuda) for t in synth.face_vertices(synth.mesh_batch(v, 4), f))
    pts = synth.random_points(4, 8192, 9).to(cuda)
    want = ref_cuda.tri_forward_t(pts, t1, t2, t3)
    g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_TC_EPS_FACTOR, 0.125)
    try:
        got = g.tri_forward(pts, t1, t2, t3, tensor=True)
        torch.cuda.synchronize()
    finally:
        g.lib().geom_debug_set(_lib.GEOM_DEBUG_TRI_TC_EPS_FACTOR, 0.0)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_full_size_properties(oracle, cuda):
    """Without the reference library: the reported (dist, option) is the oracle's pair evaluation of
    the reported face, and no sampled face is strictly closer."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    v, f = synth.uv_sphere(50, 50)
    t1c, t2c, t3c = synth.face_vertices(synth.mesh_batch(v, 32), f)
    ptsc = synth.random_points(32, 8192, 41)
    dist, point, index = (t.cpu().numpy() for t in g.tri_forward(*(t.to(cuda) for t in (ptsc, t1c, t2c, t3c))))
    assert ((point >= 0) & (point <= 6)).all() and ((index >= 0) & (index < 5000)).all()
    rng = np.random.default_rng(0)
    P, A, B, C = ptsc.numpy(), t1c.numpy(), t2c.numpy(), t3c.numpy()
    for _ in range(300):
        bi, j = int(rng.integers(32)), int(rng.integers(8192))
        k = index[bi, j]
        d, o = oracle.tri_pair(P[bi, j], A[bi, k], B[bi, k], C[bi, k])
        assert d == dist[bi, j] and o == point[bi, j]
        for kk in rng.integers(0, 5000, 20):
            d2, _ = oracle.tri_pair(P[bi, j], A[bi, kk], B[bi, kk], C[bi, kk])
            assert not (d2 < d) and not (d2 == d and kk < k)
    # a slice of rows against the full oracle scan
    sl = slice(0, 256)
    wd, wp, wi = oracle.tri_fwd(P[:2, sl], A[:2], B[:2], C[:2])
    assert (wi == index[:2, sl]).all() and (wp == point[:2, sl]).all()
    assert_same_f32(wd, dist[:2, sl])
