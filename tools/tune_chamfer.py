#!/usr/bin/env python
"""Sweep geom_chamfer_fwd CUDA-core kernel variants ("variant,strip,tile_cols", applied through the library's
geom_debug_set knobs) on the BASELINE config-2 workload; one subprocess per setting.

    python tools/tune_chamfer.py [variant,strip,tile ...]      # parent: sweeps
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import geometrics_b200 as g
    from geometrics_b200 import _lib, synth

    dev = torch.device("cuda:0")
    B, N, M = 32, 10000, 10000
    sets = [(synth.random_points(B, N, 41 + i).to(dev), synth.random_points(B, M, 141 + i).to(dev)) for i in range(8)]
    L = g.lib()
    tune = [int(x) for x in os.environ.get("_TUNE_SETTING", "").split(",") if x]
    if len(tune) >= 1:
        L.geom_debug_set(_lib.GEOM_DEBUG_CHAMFER_VARIANT, float(tune[0] + 1))
    if len(tune) >= 2:
        L.geom_debug_set(_lib.GEOM_DEBUG_CHAMFER_STRIP, float(tune[1]))
    if len(tune) >= 3:
        L.geom_debug_set(_lib.GEOM_DEBUG_CHAMFER_TILE_COLS, float(tune[2]))
    wsb = max(L.geom_chamfer_workspace_bytes(B, N, M), L.geom_chamfer_tensor_workspace_bytes(B, N, M))
    ws = torch.empty(wsb // 8 + 1, dtype=torch.int64, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
    i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream

    TENSOR = 4 if os.environ.get("GEOM_BENCH_TENSOR") else 0

    def step(k, flags=TENSOR):
        a, b = sets[k % len(sets)]
        rc = L.geom_chamfer_fwd(B, N, a.data_ptr(), M, b.data_ptr(), None, i1.data_ptr(), None, i2.data_ptr(),
                                ws.data_ptr(), wsb, flags, sp)
        assert rc == 0, rc

    # correctness against the simple one-way kernel
    step(0, 2)
    r1, r2 = i1.clone(), i2.clone()
    step(0)
    torch.cuda.synchronize()
    ok = bool(torch.equal(r1, i1) and torch.equal(r2, i2))
    for k in range(5):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 40
    e0.record()
    for k in range(K):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / K
    _lib.prof_begin()
    for k in range(K):
        step(k)
    kern = _lib.prof_end()
    print(json.dumps({"tune": os.environ.get("_TUNE_SETTING", ""), "ok": ok, "step_ms": round(total, 4),
                      **{k.split("::")[-1]: round(v[0] / v[1], 4) for k, v in kern.items()}}), flush=True)


if __name__ == "__main__":
    if os.environ.get("_TUNE_CHILD"):
        child()
    else:
        settings = sys.argv[1:] or ["0,1", "0,2", "0,4", "1,2", "2,2", "3,2", "4,2", "0,2,128", "2,2,128", "4,2,128"]
        for s in settings:
            env = dict(os.environ, _TUNE_CHILD="1", _TUNE_SETTING=s)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
