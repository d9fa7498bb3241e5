"""CPU: a numpy model of the tensor-core Chamfer scan's survivor logic (csrc/chamfer_tc.cu scan_kernel)
checked against brute force.

The model follows the kernel's control flow for one query row: 128-candidate tiles of four 32-candidate blocks, the
running minimum mt of the APPROXIMATE distances, a tile is listed when its minimum is <= mt + 2 eps, the list holds
kEvCap tiles; a full list is first COMPACTED (tiles whose minimum left the window are dropped, nothing is evaluated)
and only if more than half of it survives the thread evaluates what it has; after the sweep each listed block is
re-checked against the final mt + 2 eps and the survivors are evaluated exactly, ascending index with strict '<'
(the kernel does that last step with the whole warp per block; the winner is the lowest lane of the minimum).  The approximate distances are the exact ones plus an
ADVERSARIAL error of magnitude <= eps (random, all +eps, all -eps, alternating), so the test exercises the claim that
any |dt - d| <= eps keeps the result exact, including the lowest-index rule on exact ties and list overflows."""
import zlib

import numpy as np
import pytest

K_EV_CAP = 16


def scan_row(d_exact, dt, eps):
    """-> (best distance, index) for one query; d_exact (m,) float32 exact distances, dt (m,) approximations."""
    m = len(dt)
    pad = (-m) % 128
    dtp = np.concatenate([dt, np.full(pad, 1e30, np.float32)])
    mt, best, bidx = np.float32(np.inf), np.float32(np.inf), 0
    events = []                       # (tile, [bm0..bm3])
    evaluated = 0

    def flush():
        nonlocal best, bidx, events, evaluated
        thr = mt + 2 * eps
        blocks = [t * 4 + i for t, bms in events for i in range(4) if not (bms[i] > thr)]
        for blk in blocks:            # ascending: events are appended in tile order, blocks in block order
            for j in range(blk * 32, min(blk * 32 + 32, m)):
                evaluated += 1
                if d_exact[j] < best:
                    best, bidx = d_exact[j], j
        events = []

    def make_room():
        nonlocal events
        thr = mt + 2 * eps
        events = [(t, bms) for t, bms in events if not (bms.min() > thr)]
        if len(events) > K_EV_CAP // 2:
            flush()

    for t in range(len(dtp) // 128):
        bms = dtp[t * 128:(t + 1) * 128].reshape(4, 32).min(1)
        mt = min(mt, bms.min())
        if not (bms.min() > mt + 2 * eps):
            events.append((t, bms))
        if len(events) >= K_EV_CAP:
            make_room()
    flush()
    return best, bidx, evaluated


def brute(d_exact):
    j = int(np.argmin(d_exact))      # numpy returns the first (lowest-index) minimum
    return d_exact[j], j


@pytest.mark.parametrize("order", ["random", "descending", "ascending", "ties"])
@pytest.mark.parametrize("noise", ["random", "plus", "minus", "alternating"])
def test_survivor_logic_is_exact_for_any_bounded_error(order, noise):
    rng = np.random.default_rng(zlib.crc32(f"{order}/{noise}".encode()))
    m, eps = 5000 + 37, np.float32(6.25e-5)
    for trial in range(6):
        q = rng.random(3) - 0.5
        c = rng.random((m, 3)) - 0.5
        if order == "ties":
            c = np.round(c * 16) / 16           # lattice: many exactly equal distances
            q = np.round(q * 16) / 16
        d = ((c - q) ** 2).sum(1).astype(np.float32)
        if order == "descending":
            d = np.sort(d)[::-1].copy()          # every tile sets a new minimum: the event list overflows repeatedly
        elif order == "ascending":
            d = np.sort(d)
        e = {"random": rng.uniform(-1, 1, m), "plus": np.ones(m), "minus": -np.ones(m),
             "alternating": np.where(np.arange(m) % 2 == 0, 1.0, -1.0)}[noise]
        dt = (d.astype(np.float64) + 0.999 * float(eps) * e).astype(np.float32)
        got = scan_row(d, dt, eps)
        want = brute(d)
        assert (got[0], got[1]) == want, (order, noise, trial, got, want)
        if order == "random":
            assert got[2] <= 0.05 * m             # the filter actually filters: a few blocks are evaluated exactly


def test_infinite_margin_degenerates_to_brute_force():
    rng = np.random.default_rng(3)
    d = rng.random(777).astype(np.float32)
    got = scan_row(d, np.zeros_like(d), np.float32(np.inf))   # useless approximation, eps = inf: everything evaluated
    assert (got[0], got[1]) == brute(d) and got[2] == 777
