"""CPU tests that pin the oracle: against the reference's own CPU code (oracle/_ref/libref_cpu.so,
built from my_lib.c unmodified), against committed golden vectors produced by that code, and
against hand-derived known answers for the reference kernels' quirks."""
import os

import numpy as np
import pytest

from helpers import CHAMFER_CASES, GOLDEN, chamfer_inputs, assert_same_f32
from golden.make_golden_cpu import CASES as GOLDEN_CASES, make_inputs as golden_inputs


def brute(x1, x2):
    """float64 numpy brute force (independent of both the oracle and the reference)."""
    d = ((x1[:, :, None, :].astype(np.float64) - x2[:, None, :, :].astype(np.float64)) ** 2).sum(-1)
    return d.min(-1), d.argmin(-1)


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_reference_cpu_golden(oracle, name):
    kind, b, n, m, seed = GOLDEN_CASES[name]
    x1, x2 = golden_inputs(kind, b, n, m, seed)
    gold = np.load(os.path.join(GOLDEN, f"chamfer_refcpu_{name}.npz"))
    d1, i1, d2, i2 = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
    for d, i, gd, gi, cand, q in ((d1, i1, gold["dist1"], gold["idx1"], x2, x1), (d2, i2, gold["dist2"], gold["idx2"], x1, x2)):
        diff = i != gi
        if kind in ("lattice", "dup") and kind == "lattice":
            assert not diff.any()  # exactly representable data: must agree bit for bit
            assert_same_f32(d, gd, name)
        # The reference CPU code sums without fma; the GPU kernel (and this oracle) use fma.  They may
        # pick different candidates only when those candidates are tied to within rounding.
        np.testing.assert_allclose(d, gd, rtol=3e-7, atol=1e-12)
        assert diff.mean() < 1e-3
    # gradients: my restatement of nnd_backward vs the reference's nnd_backward on its own indices
    g1, g2 = oracle.chamfer_loss_bwd(x1.numpy(), x2.numpy(), gold["idx1"], gold["idx2"], 0.25, 0.5)
    np.testing.assert_allclose(g1, gold["grad1"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(g2, gold["grad2"], rtol=1e-5, atol=1e-7)


def test_oracle_vs_live_reference_cpu(oracle, ref_cpu):
    for kind, b, n, m, seed in CHAMFER_CASES:
        if (m % 512) % 4 or (n % 512) % 4:
            continue  # tail-skip sizes: CPU reference has no tail skip (my_lib.c:12)
        x1, x2 = chamfer_inputs(kind, b, n, m, seed)
        d1, i1, d2, i2 = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
        rd1, ri1, rd2, ri2 = ref_cpu.nnd_forward(x1.numpy(), x2.numpy())
        np.testing.assert_allclose(d1, rd1, rtol=3e-7, atol=1e-12)
        np.testing.assert_allclose(d2, rd2, rtol=3e-7, atol=1e-12)
        assert (i1 != ri1).mean() < 1e-3 and (i2 != ri2).mean() < 1e-3
    # the port of nnsearch is the same arithmetic as the original: bit-exact
    x1, x2 = chamfer_inputs("random", 2, 500, 731, 99)
    d, i = oracle.nnsearch(x1.numpy(), x2.numpy())
    rd, ri = ref_cpu.nnsearch(x1.numpy(), x2.numpy())
    assert (i == ri).all()
    assert_same_f32(d, rd)


@pytest.mark.parametrize("case", CHAMFER_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}x{c[3]}")
def test_oracle_exact_mode_vs_numpy(oracle, case):
    kind, b, n, m, seed = case
    x1, x2 = chamfer_inputs(kind, b, n, m, seed)
    x1, x2 = x1.numpy(), x2.numpy()
    d, i = oracle.chamfer_dir(x1, x2, exact=True)
    bd, bi = brute(x1, x2)
    np.testing.assert_allclose(d, bd, rtol=1e-5, atol=1e-10)
    chosen = ((x1.astype(np.float64) - np.take_along_axis(x2, i[..., None].astype(np.int64), 1)) ** 2).sum(-1)
    np.testing.assert_allclose(chosen, bd, rtol=1e-5, atol=1e-10)
    if kind == "lattice":  # exact arithmetic: lowest index among ties == numpy argmin
        assert (i == bi).all()


def test_tail_skip_known_answers(oracle):
    """chamfer_distance.cu:31-33: the last (m mod 512) mod 4 candidates are never examined."""
    q = np.zeros((1, 2, 3), np.float32)
    q[0, 1] = 0.3
    cand = np.full((1, 6, 3), 0.2, np.float32)  # m=6 -> candidates 4,5 skipped
    cand[0, 5] = 0.0                            # the true nearest of q0, but skipped
    d, i = oracle.chamfer_dir(q, cand)
    assert i.tolist() == [[0, 0]]
    de, ie = oracle.chamfer_dir(q, cand, exact=True)
    assert ie.tolist() == [[5, 0]] and de[0, 0] == 0.0
    # m = 513: the final tile holds one candidate, its scan is empty -> (best=0, idx=0) overwrites
    # every positive result (chamfer_distance.cu:29-30,47)
    rng = np.random.default_rng(0)
    cand = rng.random((1, 513, 3), dtype=np.float32)
    q = rng.random((1, 7, 3), dtype=np.float32)
    q[0, 3] = cand[0, 100]  # an exact hit keeps its result: 0 > 0 is false
    d, i = oracle.chamfer_dir(q, cand)
    assert (d == 0).all()
    assert i[0].tolist() == [0, 0, 0, 100, 0, 0, 0]
    # m < 4: nothing is scanned at all
    d, i = oracle.chamfer_dir(q, cand[:, :3])
    assert (d == 0).all() and (i == 0).all()


def test_tie_rule_lowest_index(oracle):
    """chamfer_distance.cu:39 strict '<' in a tile, :47 strict '>' across tiles."""
    cand = np.zeros((1, 1536, 3), np.float32)
    cand[0, :, 0] = 1.0
    for k in (7, 8, 600, 1400):
        cand[0, k] = (0.5, 0.0, 0.0)  # four exactly tied nearest candidates, in three tiles
    q = np.zeros((1, 1, 3), np.float32)
    d, i = oracle.chamfer_dir(q, cand)
    assert i[0, 0] == 7 and d[0, 0] == 0.25


def test_loss_sums_and_grad_consistency(oracle):
    x1, x2 = chamfer_inputs("random", 2, 300, 411, 5)
    x1, x2 = x1.numpy(), x2.numpy()
    _, i1, _, i2 = oracle.chamfer_fwd(x1, x2, exact=True)
    s1, s2 = oracle.chamfer_loss_sums(x1, x2, i1, i2)
    w1, w2 = 0.7, 1.3
    g1, g2 = oracle.chamfer_loss_bwd(x1, x2, i1, i2, w1, w2)
    eps = 1e-3
    rng = np.random.default_rng(1)
    dirn = rng.standard_normal(x1.shape).astype(np.float32)
    s1p, s2p = oracle.chamfer_loss_sums(x1 + eps * dirn, x2, i1, i2)
    s1m, s2m = oracle.chamfer_loss_sums(x1 - eps * dirn, x2, i1, i2)
    fd = (w1 * (s1p - s1m) + w2 * (s2p - s2m)) / (2 * eps)
    assert abs(fd - float((g1 * dirn).sum())) < 1e-3 * max(1.0, abs(fd))
    assert g2.shape == x2.shape
