#!/usr/bin/env python
"""Chamfer tensor-core scan: time at C2 and (CH_TIMING builds) the per-warp phase timers of a development variant.
usage: ch_timing.py build_variants/libgeom_x.so"""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometrics_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
import geometrics_b200 as g
from geometrics_b200 import synth
dev = torch.device("cuda:0")
out = {"lib": os.path.basename(_lib.LIB_PATH)}
for b in (32, 4):
    sets = [(synth.random_points(b, 10000, 41 + i).to(dev), synth.random_points(b, 10000, 141 + i).to(dev)) for i in range(max(2, 96 // b))]
    f = lambda i: g.chamfer_forward(*sets[i % len(sets)], want_dist=False)
    for i in range(3): f(i)
    _lib.prof_begin()
    for i in range(20): f(i)
    k = _lib.prof_end()
    want = g.chamfer_forward(*sets[0], tensor=False)
    got = g.chamfer_forward(*sets[0])
    out[f"b{b}_scan_ms"] = round(k["geom::chamfer_tc::scan_kernel"][0] / 20, 4)
    out[f"b{b}_exact"] = all(bool(torch.equal(a, c)) for a, c in zip(got, want))
print(json.dumps(out), flush=True)
L = g.lib()
if hasattr(L, "geom_debug_read_ch_timing"):
    sets = [(synth.random_points(32, 10000, 41).to(dev), synth.random_points(32, 10000, 141).to(dev))]
    g.chamfer_forward(*sets[0], want_dist=False)
    torch.cuda.synchronize()
    L.geom_debug_read_ch_timing.argtypes = [C.c_void_p, C.c_size_t]
    n = 256 * 20 * 8
    buf = np.zeros(n, np.uint64)
    assert L.geom_debug_read_ch_timing(buf.ctypes.data, n) == 0
    t = buf.reshape(256, 20, 8).astype(np.float64) / 79.0   # cycles per tile (C2: 79 tiles)
    nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 16   # scanner warps per CTA
    sc, pr = t[:, :nsw], t[:, nsw:nsw + nsw // 4]
    names_s = ["wait full0", "ld0", "min0", "wait full1", "ld1", "min1+tail", "flush: list", "flush: evaluate"]
    names_p = ["refill", "wait b_full", "wait e0", "mma0", "wait e1", "mma1"]
    print("  scanner cycles/tile :", "  ".join(f"{nm} {sc[:, :, i].mean():.0f}" for i, nm in enumerate(names_s)), f"  sum {sc[:, :, :6].sum(2).mean():.0f}")
    print("  producer cycles/tile:", "  ".join(f"{nm} {pr[:, :, i].mean():.0f}" for i, nm in enumerate(names_p)), f"  sum {pr[:, :, :6].sum(2).mean():.0f}")
    print("  producer g=0        :", "  ".join(f"{nm} {pr[:, 0, i].mean():.0f}" for i, nm in enumerate(names_p)))
    if hasattr(L, "geom_debug_read_ch_stamps"):
        L.geom_debug_read_ch_stamps.argtypes = [C.c_void_p, C.c_size_t]
        sb = np.zeros(2048 * 8, np.uint64)
        assert L.geom_debug_read_ch_stamps(sb.ctypes.data, sb.size) == 0
        st = sb.reshape(2048, 8)[:1280].astype(np.int64)
        names = ["entry -> barriers/TMEM ready", "-> first tile ready", "-> sweep done", "-> re-evaluation done", "-> results written", "-> TMEM freed (exit)"]
        d = np.diff(st[:, :7], axis=1) / 1000.0
        print("  CTA time line (us, mean / median over 1280 CTAs):", "  ".join(f"{n} {d[:, i].mean():.2f}/{np.median(d[:, i]):.2f}" for i, n in enumerate(names)), f"  total {(st[:, 6] - st[:, 0]).mean() / 1000.0:.2f}")
        # gap between consecutive CTAs on the same SM
        gaps = []
        for sm in np.unique(st[:, 7]):
            rows = st[st[:, 7] == sm]
            rows = rows[np.argsort(rows[:, 0])]
            gaps += list((rows[1:, 0] - rows[:-1, 6]) / 1000.0)
        print(f"  gap between a CTA's exit and the next CTA's entry on the same SM: mean {np.mean(gaps):.2f} us, median {np.median(gaps):.2f} us;  kernel span {(st[:, 6].max() - st[:, 0].min()) / 1000.0:.1f} us")
