"""CPU: the error bound behind the tensor-core Chamfer filter (csrc/chamfer_tc.cu).

pack_kernel's operand construction is restated in numpy (bf16 round-to-nearest-even splits of the coordinates and of
the squared norms, the K=16 row layouts) and the contraction dt = A.B is compared with the exact squared distance:
|dt - d| must stay below the shipped filter margin eps = 2.5e-4 * R^2 (R = max |coordinate|) with room to spare, on
several coordinate ranges.  The products of two bf16 values are exact in fp32, so the only error sources are the
dropped lo*lo terms, the bf16 rounding of the lo parts / norm residuals, and the fp32 accumulation (bounded
separately below)."""
import numpy as np
import pytest

K_EPS_SCALE = 2.5e-4  # chamfer_tc.cu kEpsScale


def bf16(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def forms(p):
    """(A rows, B rows) of pack_kernel for points p (N,3) float32: two (N,16) float32 arrays of bf16 values."""
    p = p.astype(np.float32)
    h = bf16(p)
    lo = bf16(p - h)                     # x - xh is exact in fp32; rounded to bf16 when packed
    nn = (p[:, 2] * p[:, 2] + (p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1])).astype(np.float32)
    n1 = bf16(nn)
    n2 = bf16(nn - n1)
    n3 = bf16(nn - n1 - n2)
    one, zero = np.ones_like(nn), np.zeros_like(nn)
    m2 = np.float32(-2.0)
    a = np.stack([m2 * h[:, 0], m2 * h[:, 1], m2 * h[:, 2], bf16(m2 * (p - h)[:, 0]), bf16(m2 * (p - h)[:, 1]), bf16(m2 * (p - h)[:, 2]),
                  m2 * h[:, 0], m2 * h[:, 1], m2 * h[:, 2], n1, n2, n3, one, one, one, zero], axis=1)
    b = np.stack([h[:, 0], h[:, 1], h[:, 2], h[:, 0], h[:, 1], h[:, 2], lo[:, 0], lo[:, 1],
                  lo[:, 2], one, one, one, n1, n2, n3, zero], axis=1)
    return a.astype(np.float32), b.astype(np.float32)


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (100.0, 0.0), (1e-3, 0.0), (1.0, 50.0), (1e6, 0.0), (1.0, -0.37)])
def test_bf16_split_contraction_error_is_inside_the_filter_margin(scale, offset):
    rng = np.random.default_rng(7)
    n = 1500
    p = ((rng.random((n, 3)) - 0.5) * scale + offset).astype(np.float32)
    q = ((rng.random((n, 3)) - 0.5) * scale + offset).astype(np.float32)
    q[:50] = p[:50] + (1e-4 * scale * rng.standard_normal((50, 3))).astype(np.float32)  # near-coincident pairs
    a, _ = forms(p)
    _, b = forms(q)
    dt = a.astype(np.float64) @ b.astype(np.float64).T                 # exact products, ideal accumulation
    d = ((p[:, None, :].astype(np.float64) - q[None].astype(np.float64)) ** 2).sum(-1)
    r = float(max(np.abs(p).max(), np.abs(q).max()))
    err = np.abs(dt - d).max() / (r * r)
    # fp32 accumulation of 15 terms each <= 4 R^2 in magnitude: <= 15 * 2^-24 * 3 * 4 R^2 ~ 1.1e-5 R^2 on top
    accumulate = 15 * 2.0 ** -24 * 12
    assert err + accumulate <= 0.5 * K_EPS_SCALE, (err, accumulate)
    assert err <= 1.0e-4                                                  # the derivation's 7.9e-5 R^2


def test_operand_rows_match_the_documented_layout():
    p = np.array([[0.3, -0.2, 0.45]], np.float32)
    a, b = forms(p)
    # dt(p, p) ~ 0: |p|^2 + |p|^2 - 2 p.p
    assert abs(float(a[0].astype(np.float64) @ b[0].astype(np.float64))) <= 1e-4 * 0.45 ** 2
    assert a[0, 15] == 0 and b[0, 15] == 0 and (a[0, 12:15] == 1).all() and (b[0, 9:12] == 1).all()
