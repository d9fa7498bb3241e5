"""GPU parity tests of geom_chamfer_fwd (through the C ABI via geometrics_b200.chamfer_forward):
bit-exact distances and indices against (1) the CPU oracle and (2) the reference's own
chamfer_distance.cu compiled unmodified (oracle/_ref/libref_cuda.so) on the same inputs, plus
size-independent properties at the BASELINE config-2 size."""
import numpy as np
import pytest
import torch

from helpers import CHAMFER_CASES, chamfer_inputs, assert_same_f32

pytestmark = pytest.mark.gpu

MODES = [dict(), dict(oneway=True), dict(exact_tail=True), dict(exact_tail=True, oneway=True),
         dict(tensor=False), dict(tensor=False, exact_tail=True)]  # default = library choice (tensor-core search)
_ids = lambda c: f"{c[0]}-{c[1]}x{c[2]}x{c[3]}"  # noqa: E731


def _run(x1, x2, cuda, **kw):
    import geometrics_b200 as g

    d1, i1, d2, i2 = g.chamfer_forward(x1.to(cuda), x2.to(cuda), **kw)
    torch.cuda.synchronize()
    return d1.cpu().numpy(), i1.cpu().numpy(), d2.cpu().numpy(), i2.cpu().numpy()


@pytest.mark.parametrize("mode", MODES, ids=lambda m: "+".join(sorted(m)) or "default")
@pytest.mark.parametrize("case", CHAMFER_CASES, ids=_ids)
def test_chamfer_matches_oracle(oracle, cuda, case, mode):
    x1, x2 = chamfer_inputs(*case)
    got = _run(x1, x2, cuda, **mode)
    want = oracle.chamfer_fwd(x1.numpy(), x2.numpy(), exact=mode.get("exact_tail", False))
    for name, g_, w_ in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
        if name.startswith("dist"):
            assert_same_f32(g_, w_, name)
        else:
            assert (g_ == w_).all(), f"{name}: {np.count_nonzero(g_ != w_)} of {w_.size} indices differ"


@pytest.mark.parametrize("case", CHAMFER_CASES, ids=_ids)
def test_chamfer_matches_reference_cuda(ref_cuda, cuda, case):
    """Default mode is bit-compatible with the reference .cu, tail-skip quirks included."""
    x1, x2 = chamfer_inputs(*case)
    a, b = x1.to(cuda), x2.to(cuda)
    rd1, ri1, rd2, ri2 = ref_cuda.chamfer_forward_t(a, b)
    import geometrics_b200 as g

    d1, i1, d2, i2 = g.chamfer_forward(a, b)
    torch.cuda.synchronize()
    assert torch.equal(i1, ri1) and torch.equal(i2, ri2)
    assert torch.equal(d1, rd1) and torch.equal(d2, rd2)


def test_oracle_matches_reference_cuda(oracle, ref_cuda, cuda):
    """Pins the CPU oracle itself against the reference GPU kernel."""
    for case in CHAMFER_CASES:
        x1, x2 = chamfer_inputs(*case)
        rd1, ri1, rd2, ri2 = ref_cuda.chamfer_forward_t(x1.to(cuda), x2.to(cuda))
        d1, i1, d2, i2 = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
        assert (i1 == ri1.cpu().numpy()).all() and (i2 == ri2.cpu().numpy()).all(), case
        assert_same_f32(d1, rd1.cpu().numpy(), str(case))
        assert_same_f32(d2, rd2.cpu().numpy(), str(case))


def test_module_surface(cuda):
    """ChamferDistance()(xyz1, xyz2) -> (idx1, idx2): int32, on the inputs' device; accepts a
    non-contiguous slice (GEOMetrics.py:227 passes gt_samples[:, :2466]) and requires_grad inputs."""
    from chamfer_distance import ChamferDistance
    import geometrics_b200 as g

    x1 = (torch.rand(2, 3000, 3, device=cuda) - 0.5).requires_grad_(True)
    gt = torch.rand(2, 5000, 3, device=cuda) - 0.5
    idx1, idx2 = ChamferDistance()(gt[:, :2466], x1)
    assert idx1.dtype == torch.int32 and idx1.shape == (2, 2466) and idx1.device == x1.device
    assert idx2.dtype == torch.int32 and idx2.shape == (2, 3000)
    assert not idx1.requires_grad
    _, e1, _, e2 = g.chamfer_forward(gt[:, :2466].contiguous(), x1.detach())
    assert torch.equal(idx1, e1) and torch.equal(idx2, e2)


def test_argument_errors(cuda):
    import geometrics_b200 as g

    a = torch.zeros(2, 8, 3, device=cuda)
    with pytest.raises(ValueError):
        g.chamfer_forward(a, torch.zeros(3, 8, 3, device=cuda))
    with pytest.raises(ValueError):
        g.chamfer_forward(a, torch.zeros(2, 8, 2, device=cuda))
    with pytest.raises(TypeError):
        g.chamfer_forward(a, a.double())
    L = g.lib()
    # workspace too small -> GEOM_ERR_WORKSPACE, null output -> GEOM_ERR_INVALID_ARG
    i = torch.zeros(2, 8, dtype=torch.int32, device=cuda)
    assert L.geom_chamfer_fwd(2, 8, a.data_ptr(), 8, a.data_ptr(), None, i.data_ptr(), None, i.data_ptr(),
                              None, 0, 0, None) == -2
    assert L.geom_chamfer_fwd(2, 8, a.data_ptr(), 8, a.data_ptr(), None, None, None, i.data_ptr(),
                              None, 0, 0, None) == -1
    assert L.geom_chamfer_fwd(-1, 8, a.data_ptr(), 8, a.data_ptr(), None, i.data_ptr(), None, i.data_ptr(),
                              None, 0, 0, None) == -1


def test_empty_inputs(cuda):
    import geometrics_b200 as g

    d1, i1, d2, i2 = g.chamfer_forward(torch.zeros(0, 5, 3, device=cuda), torch.zeros(0, 7, 3, device=cuda))
    assert i1.shape == (0, 5) and i2.shape == (0, 7)
    # no candidates at all: the reference kernel's tile loop never runs and outputs stay zero
    d1, i1, d2, i2 = g.chamfer_forward(torch.rand(2, 5, 3, device=cuda), torch.zeros(2, 0, 3, device=cuda))
    assert (i1 == 0).all() and (d1 == 0).all() and i2.shape == (2, 0)


def test_full_size_config2(ref_cuda, cuda):
    """BASELINE config 2: 32 x (10000 vs 10000), seed 41.  Bit-exact against the reference kernel;
    plus properties that do not need a second implementation."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    x1 = synth.random_points(32, 10000, 41).to(cuda)
    x2 = synth.random_points(32, 10000, 42).to(cuda)
    d1, i1, d2, i2 = g.chamfer_forward(x1, x2)
    rd1, ri1, rd2, ri2 = ref_cuda.chamfer_forward_t(x1, x2)
    assert torch.equal(i1, ri1) and torch.equal(i2, ri2)
    assert torch.equal(d1, rd1) and torch.equal(d2, rd2)
    _full_size_properties(x1, x2, d1, i1, d2, i2)


def _full_size_properties(x1, x2, d1, i1, d2, i2):
    b, n, _ = x1.shape
    m = x2.shape[1]
    assert int(i1.min()) >= 0 and int(i1.max()) < m and int(i2.min()) >= 0 and int(i2.max()) < n
    # the reported distance is the distance to the reported neighbour (fp32, 1e-5 relative)
    nn1 = torch.gather(x2, 1, i1.long().unsqueeze(-1).expand(b, n, 3))
    nn2 = torch.gather(x1, 1, i2.long().unsqueeze(-1).expand(b, m, 3))
    torch.testing.assert_close(((nn1 - x1) ** 2).sum(-1), d1, rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(((nn2 - x2) ** 2).sum(-1), d2, rtol=1e-5, atol=1e-12)
    # minimality on a sample of rows, against torch.cdist-free fp64 brute force
    rows = torch.arange(0, n, 97, device=x1.device)
    dd = ((x1[:, rows, None, :].double() - x2[:, None, :, :].double()) ** 2).sum(-1)  # (B, r, m)
    torch.testing.assert_close(dd.min(-1).values.float(), d1[:, rows], rtol=1e-5, atol=1e-12)
    # swapping the arguments swaps the outputs bit for bit (d(p,q) is symmetric in fp32)
    import geometrics_b200 as g

    s1, j1, s2, j2 = g.chamfer_forward(x2, x1)
    assert torch.equal(j1, i2) and torch.equal(j2, i1) and torch.equal(s1, d2) and torch.equal(s2, d1)
    # one-way and two-way kernels agree
    o1, k1, o2, k2 = g.chamfer_forward(x1, x2, oneway=True)
    assert torch.equal(k1, i1) and torch.equal(k2, i2) and torch.equal(o1, d1) and torch.equal(o2, d2)


def test_full_size_properties_without_reference(cuda):
    """Same properties on tie-heavy lattice data at full size (no reference library needed)."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    x1 = synth.lattice_points(32, 10000, 64, 5).to(cuda)
    x2 = synth.lattice_points(32, 10000, 64, 6).to(cuda)
    d1, i1, d2, i2 = g.chamfer_forward(x1, x2)
    _full_size_properties(x1, x2, d1, i1, d2, i2)
    # lattice data is exact in fp32: lowest index among ties == first argmin of the exact distances
    rows = torch.arange(0, 10000, 211, device=cuda)
    dd = ((x1[:, rows, None, :].double() - x2[:, None, :, :].double()) ** 2).sum(-1)
    assert torch.equal(dd.argmin(-1).int(), i1[:, rows]) or _first_min_equal(dd, i1[:, rows])


def _first_min_equal(dd, idx):
    # torch.argmin does not promise the first minimum on CUDA; check explicitly
    mn = dd.min(-1, keepdim=True).values
    first = (dd == mn).int().argmax(-1)
    return torch.equal(first.int(), idx)


@pytest.mark.parametrize("pinned", [True, False])
def test_host_buffer_entry(oracle, cuda, pinned):
    """geom_chamfer_fwd_host: host tensors in, host indices (and distances) out, chunk-pipelined."""
    import geometrics_b200 as g

    x1, x2 = chamfer_inputs("random", 5, 1500, 1030, 77)
    want = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
    bufs = [torch.empty(5, 1500, dtype=torch.int32), torch.empty(5, 1030, dtype=torch.int32),
            torch.empty(5, 1500), torch.empty(5, 1030)]
    if pinned:
        x1, x2 = x1.pin_memory(), x2.pin_memory()
        bufs = [t.pin_memory() for t in bufs]
    for chunks in (1, 2, 4, 8):
        for t in bufs:
            t.fill_(-1)
        g.chamfer_forward_host(x1, x2, bufs[0], bufs[1], device=cuda, chunks=chunks, dist1_out=bufs[2], dist2_out=bufs[3])
        assert (bufs[0].numpy() == want[1]).all() and (bufs[1].numpy() == want[3]).all(), chunks
        assert_same_f32(bufs[2].numpy(), want[0])
        assert_same_f32(bufs[3].numpy(), want[2])
    with pytest.raises(RuntimeError):
        g.chamfer_forward_host(x1.to(cuda), x2, bufs[0], bufs[1])


# ---------------------------------------------------------------------------- tensor-core filtered path
@pytest.mark.parametrize("exact", [False, True], ids=["default", "exact_tail"])
@pytest.mark.parametrize("case", CHAMFER_CASES, ids=_ids)
def test_tensor_path_matches_oracle(oracle, cuda, case, exact):
    """GEOM_FLAG_TENSOR (tcgen05 bf16 filter + exact fp32 re-evaluation) is bit-identical to the oracle."""
    x1, x2 = chamfer_inputs(*case)
    got = _run(x1, x2, cuda, tensor=True, exact_tail=exact)
    want = oracle.chamfer_fwd(x1.numpy(), x2.numpy(), exact=exact)
    for name, g_, w_ in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
        if name.startswith("dist"):
            assert_same_f32(g_, w_, name)
        else:
            assert (g_ == w_).all(), f"{name}: {np.count_nonzero(g_ != w_)} of {w_.size} indices differ"


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (100.0, 0.0), (1e-3, 0.0), (1.0, 50.0), (1e6, 0.0)])
def test_tensor_path_full_size_and_ranges(cuda, scale, offset):
    """Config-2 size and other coordinate ranges (the filter margin scales with max|coord|^2): equal to
    the brute-force kernel bit for bit; lattice / duplicate data exercise ties."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    b = 32 if scale == 1.0 and offset == 0.0 else 4
    for mk in (synth.random_points, lambda b_, n_, s_: synth.lattice_points(b_, n_, 64, s_)):
        x1 = (mk(b, 10000, 41) * scale + offset).to(cuda)
        x2 = (mk(b, 10000, 42) * scale + offset).to(cuda)
        want = g.chamfer_forward(x1, x2, tensor=False)
        got = g.chamfer_forward(x1, x2, tensor=True)
        auto = g.chamfer_forward(x1, x2)
        for a_, b_, c_ in zip(got, want, auto):
            assert torch.equal(a_, b_) and torch.equal(c_, b_)


def test_tensor_path_margin(cuda):
    """The filter stays exact with its margin shrunk 8x: the shipped margin has headroom."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    x1, x2 = synth.random_points(8, 10000, 3).to(cuda), synth.random_points(8, 10000, 4).to(cuda)
    want = g.chamfer_forward(x1, x2, tensor=False)
    from geometrics_b200 import _lib

    g.lib().geom_debug_set(_lib.GEOM_DEBUG_TC_EPS_FACTOR, 0.125)
    try:
        got = g.chamfer_forward(x1, x2, tensor=True)
        torch.cuda.synchronize()
    finally:
        g.lib().geom_debug_set(_lib.GEOM_DEBUG_TC_EPS_FACTOR, 0.0)
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("chunks", [2, 3, 4, 8])
def test_tensor_path_candidate_split(oracle, cuda, chunks):
    """Small batches split every row block's candidate sweep over several CTAs (results meet in 64-bit keys, the last
    CTA to arrive decodes): same bits as the oracle, tail semantics and lattice / duplicate ties included."""
    import geometrics_b200 as g
    from geometrics_b200 import _lib

    g.lib().geom_debug_set(_lib.GEOM_DEBUG_TC_CHUNKS, float(chunks))
    try:
        for case in [("random", 2, 2466, 3000, 6), ("random", 1, 300, 10001, 8), ("lattice", 2, 4100, 4099, 11),
                     ("dup", 2, 3000, 2466, 13), ("random", 2, 1500, 999, 7), ("cluster", 2, 1640, 1300, 15)]:
            x1, x2 = chamfer_inputs(*case)
            for exact in (False, True):
                got = _run(x1, x2, cuda, tensor=True, exact_tail=exact)
                want = oracle.chamfer_fwd(x1.numpy(), x2.numpy(), exact=exact)
                for name, g_, w_ in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
                    if name.startswith("dist"):
                        assert_same_f32(g_, w_, f"{case} {name}")
                    else:
                        assert (g_ == w_).all(), f"{case} {name}: {np.count_nonzero(g_ != w_)} of {w_.size} indices differ"
    finally:
        g.lib().geom_debug_set(_lib.GEOM_DEBUG_TC_CHUNKS, 0.0)


@pytest.mark.parametrize("b", [1, 4, 8])
def test_tensor_path_small_batches_full_size(cuda, b):
    """The per-GPU batches of the strong-scaling split (32 items over 4 / 8 GPUs) at full size: the library's own choice
    of the candidate split must match the CUDA-core kernel bit for bit."""
    from geometrics_b200 import synth
    import geometrics_b200 as g

    x1, x2 = synth.random_points(b, 10000, 51).to(cuda), synth.random_points(b, 10000, 52).to(cuda)
    want = g.chamfer_forward(x1, x2, tensor=False)
    got = g.chamfer_forward(x1, x2)
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("mode", [dict(), dict(tensor=False), dict(oneway=True), dict(tensor=True, exact_tail=True),
                                  dict(tensor=False, exact_tail=True)],
                         ids=["tensor", "cuda_cores", "oneway", "tensor_exact", "cuda_cores_exact"])
def test_nonfinite_inputs_follow_the_reference(oracle, cuda, mode):
    """NaN / Inf coordinates: the reference takes the first candidate of every 512-tile unconditionally
    (chamfer_distance.cu:39) and merges tiles with a strict '>' (:47), so a NaN distance at a tile start voids that tile
    and a NaN at tile 0 sticks.  Every path reproduces the oracle (= the reference kernel restated) bit for bit."""
    x1, x2 = chamfer_inputs("random", 2, 1300, 1538, 91)
    x2[0, 0, 1] = float("nan")        # candidate 0 of tile 0
    x2[0, 512, 0] = float("inf")      # first candidate of tile 1
    x2[1, 700] = float("nan")         # mid-tile: simply never the minimum
    x2[1, 1024, 2] = float("-inf")    # first candidate of tile 2
    x1[0, 3, 0] = float("nan")        # NaN query
    x1[1, 77] = float("inf")          # Inf query
    x1[1, 512, 1] = float("nan")      # also a tile start of the other direction
    got = _run(x1, x2, cuda, **mode)
    want = oracle.chamfer_fwd(x1.numpy(), x2.numpy(), exact=bool(mode.get("exact_tail")))
    for name, g_, w_ in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
        if name.startswith("dist"):
            assert_same_f32(g_, w_, name)
        else:
            assert (g_ == w_).all(), f"{name}: {np.count_nonzero(g_ != w_)} of {w_.size} indices differ"


def test_nonfinite_inputs_match_reference_cuda(ref_cuda, cuda):
    import geometrics_b200 as g

    x1, x2 = chamfer_inputs("random", 2, 1030, 2466, 92)
    x2[0, 0] = float("nan")
    x2[1, 512, 2] = float("nan")
    x1[0, 9, 1] = float("inf")
    x1d, x2d = x1.to(cuda), x2.to(cuda)
    want = ref_cuda.chamfer_forward_t(x1d, x2d)
    for kw in (dict(), dict(tensor=False), dict(oneway=True)):
        got = g.chamfer_forward(x1d, x2d, **kw)
        torch.cuda.synchronize()
        for a_, b_ in zip(got, want):
            assert_same_f32(a_.cpu().numpy().view(np.float32) if a_.dtype == torch.float32 else a_.cpu().numpy().astype(np.float32),
                            b_.cpu().numpy().view(np.float32) if b_.dtype == torch.float32 else b_.cpu().numpy().astype(np.float32), str(kw))


@pytest.mark.parametrize("mode", [dict(), dict(tensor=False)], ids=["tensor", "cuda_cores"])
def test_nan_only_input_is_detected(oracle, cuda, mode):
    """A single NaN coordinate (no Inf anywhere) must trigger the non-finite slow path too: fmaxf-style range scans drop
    NaNs, so the detection has to propagate them explicitly."""
    x1, x2 = chamfer_inputs("random", 1, 700, 1100, 93)
    x2[0, 512, 1] = float("nan")   # first candidate of the second 512-tile
    x1[0, 0, 2] = float("nan")
    got = _run(x1, x2, cuda, **mode)
    want = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
    for name, g_, w_ in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
        if name.startswith("dist"):
            assert_same_f32(g_, w_, name)
        else:
            assert (g_ == w_).all(), f"{name}: {np.count_nonzero(g_ != w_)} of {w_.size} indices differ"


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (100.0, 0.0), (1e-3, 0.0), (1.0, 50.0), (1e6, 0.0), (1e4, -3e4)])
def test_filter_error_bound_on_the_hardware(cuda, scale, offset):
    """The approximate distances dt read back from TMEM (geom_debug_tc_dump: same pack kernel, same tcgen05.mma, same
    tcgen05.ld as the search) against the exact squared distances in float64: |dt - d| <= eps / 3 with eps = 2.5e-4 R^2, the
    shipped margin (the search lists a block when its minimum is within 2 eps of the running minimum), and the tensor
    core's fp32 accumulation stays within 1.2e-5 R^2 of the ideally accumulated contraction of the same operands."""
    import geometrics_b200 as g
    from test_tc_filter_bound_cpu import forms

    rng = np.random.default_rng(11)
    n, m = 128, 1000
    p = ((rng.random((n, 3)) - 0.5) * scale + offset).astype(np.float32)
    q = ((rng.random((m, 3)) - 0.5) * scale + offset).astype(np.float32)
    q[:40] = p[:40] + (1e-4 * scale * rng.standard_normal((40, 3))).astype(np.float32)  # near-coincident pairs
    L = g.lib()
    pd, qd = torch.from_numpy(p).to(cuda), torch.from_numpy(q).to(cuda)
    mp = (m + 127) // 128 * 128
    dt = torch.empty(128, mp, device=cuda)
    rmax = torch.empty(1, device=cuda)
    wsb = L.geom_chamfer_tensor_workspace_bytes(1, 128, m)
    ws = torch.empty(wsb // 8 + 2, dtype=torch.int64, device=cuda)
    rc = L.geom_debug_tc_dump(n, pd.data_ptr(), m, qd.data_ptr(), dt.data_ptr(), rmax.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    r = float(rmax)
    assert r == float(max(np.abs(p).max(), np.abs(q).max()))
    got = dt[:, :m].cpu().numpy().astype(np.float64)
    d = ((p[:, None, :].astype(np.float64) - q[None].astype(np.float64)) ** 2).sum(-1)
    eps = 2.5e-4 * r * r
    err = np.abs(got - d).max()
    a, _ = forms(p)
    _, b = forms(q)
    ideal = a.astype(np.float64) @ b.astype(np.float64).T
    acc = np.abs(got - ideal).max()
    print(f"scale {scale} offset {offset}: max|dt-d| = {err / (r * r):.3e} R^2 = {err / eps:.3f} eps; accumulation {acc / (r * r):.3e} R^2")
    assert err <= eps / 3, (err / eps)   # measured on B200: 0.12 ... 0.27 eps over these ranges
    assert acc <= 1.2e-5 * r * r, acc / (r * r)   # measured: ~1e-6 R^2
    assert (dt[:, m:] > 1e29).all()  # padding candidates can never be a minimum
