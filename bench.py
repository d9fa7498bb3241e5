#!/usr/bin/env python
"""bench.py -- headline benchmark: Chamfer point-pairs/s on 32 x (10000 vs 10000) fp32 (BASELINE
config 2), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]

A "step" is one ChamferDistance nearest-neighbour search (both directions, the reference's
ChamferDistanceKernelLauncher, chamfer_distance.cu:57-73) followed by the Chamfer loss sums on its
indices (geom_chamfer_loss_fwd, utils.py:402-417) and, when the batch is sharded, ONE NCCL all-reduce of
the two partial sums.  Default --scaling strong: the SAME 32 point-set pairs split
32/16/8/4 per GPU at N = 1/2/4/8 (BASELINE metric: "32x10k vs 10k at 1/2/4/8 B200"); --scaling weak keeps
32 per GPU.  point-pairs per step = global batch x n x m.  The all-reduce of step k is asynchronous: it
overlaps the searches of steps k+1, k+2 and is waited for two steps later, inside the timed region.

Output: one JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = the same metric through the public API with pinned HOST buffers, H2D/D2H copies inside the
timed region; `roofline` = algorithmic HBM bytes of the dominant kernel / its measured duration vs
MEASURED_PEAKS.json (the search is FP32-issue bound, see DESIGN.md: the FP32 fraction is reported
beside it); `cpu_baseline` = the reference's own CPU nnsearch (oracle/_ref/libref_cpu.so, else the
oracle port) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, N_PTS, M_PTS, SEED = 32, 10000, 10000, 41
METRIC = "chamfer_point_pairs_per_sec"
UNIT = "point-pairs/s"
WORKLOAD = "ChamferDistance 32x(10000 vs 10000) fp32 NN search, both directions (BASELINE configs[1])"
RING = 24  # distinct input/output sets cycled through so the working set (24 x 15.4 MB) exceeds L2


def _config(n_gpus, scaling="strong"):
    per_gpu = B // n_gpus if scaling == "strong" else B
    return {"workload": WORKLOAD, "batch_per_gpu": per_gpu, "n": N_PTS, "m": M_PTS, "seed": SEED,
            "global_batch": per_gpu * n_gpus,
            "parallelism": f"batch axis sharded x{n_gpus} ({scaling} scaling), one all-reduce of the 2 loss sums per step"
                           if n_gpus > 1 else "single GPU",
            "step": "NN search (both directions) + Chamfer loss sums" + (" + NCCL all-reduce of the sums" if n_gpus > 1 else ""),
            "l2_policy": f"ring of {RING} distinct input/output/workspace sets (> 126 MB L2 at N=1), no reuse between consecutive steps"}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's nnsearch (old_GEOMetrics/chamfer_distance/src/my_lib.c:4-26)
# ------------------------------------------------------------------------------------------------
class CpuBaseline:
    """Times the reference CPU path.  This (and smoke()/tests) is the only place oracle/ is used."""

    def __init__(self):
        from oracle import Oracle, RefCpu, have_ref_cpu

        if have_ref_cpu():
            self.kind, self._search = "reference", RefCpu().nnsearch
        else:
            orc = Oracle()
            self.kind, self._search = "port", lambda a, b, d=None, i=None: orc.nnsearch(a, b)
        self.threads = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))

    def _item(self, x1, x2):
        # both directions, as nnd_forward does (my_lib.c:42-43); one batch item per call
        self._search(x1, x2)
        self._search(x2, x1)

    def run(self, x1, x2, items, threads):
        """Search `items` batch items (cycled from x1/x2) on `threads` threads; -> seconds."""
        import numpy as np

        pairs = [(np.ascontiguousarray(x1[k % len(x1)][None]), np.ascontiguousarray(x2[k % len(x2)][None]))
                 for k in range(items)]
        nxt = [0]
        lock = threading.Lock()

        def work():
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= items:
                    return
                self._item(*pairs[k])  # ctypes releases the GIL

        t0 = time.perf_counter()
        ts = [threading.Thread(target=work) for _ in range(min(threads, items))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0


def _host_inputs(seed):
    from geometrics_b200 import synth

    return synth.random_points(B, N_PTS, seed), synth.random_points(B, M_PTS, seed + 1000)


def cpu_baseline_block(budget_s=12.0):
    cb = CpuBaseline()
    x1, x2 = _host_inputs(SEED)
    x1, x2 = x1.numpy(), x2.numpy()
    t1 = cb.run(x1, x2, 1, 1)  # one item, one thread: how the reference itself runs it
    items = int(max(cb.threads, min(4 * B, budget_s / max(t1, 1e-3) * cb.threads)))
    t = cb.run(x1, x2, items, cb.threads)
    return {
        "value": items * N_PTS * M_PTS / t, "unit": UNIT, "cores": cb.threads, "kind": cb.kind,
        "sample": f"{items} batch items of the workload (each 10000 vs 10000, both directions) over {cb.threads} threads, {t:.2f} s",
        "single_thread_value": N_PTS * M_PTS / t1, "single_thread_sample": f"1 batch item, 1 thread, {t1:.3f} s",
        "source": "old_GEOMetrics/chamfer_distance/src/my_lib.c nnsearch compiled unmodified (gcc -O3)" if cb.kind == "reference"
                  else "oracle/chamfer_oracle.c oracle_nnsearch (restatement of my_lib.c:4-26)",
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = CpuBaseline()
    x1, x2 = _host_inputs(SEED)
    x1, x2 = x1.numpy(), x2.numpy()
    t1 = cb.run(x1, x2, 1, 1)
    # size the per-step sample so that the whole run stays within ~2 minutes
    per_step = 120.0 / max(1, args.steps + args.warmup)
    items = int(max(1, min(B, per_step / max(t1, 1e-3) * cb.threads)))
    for _ in range(args.warmup):
        cb.run(x1, x2, items, cb.threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb.run(x1, x2, items, cb.threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = items * N_PTS * M_PTS / dt
    sample = f"{items} of the {B} batch items per step over {cb.threads} host threads (each 10000 vs 10000, both directions)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cb.threads, "kind": cb.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (pynvml thread; nvidia-smi fallback)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._loop, daemon=True)

    def sample(self):
        """One reading (SM clock + throttle reasons) from the calling thread."""
        if self._h is None:
            return
        nv = self._nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(0.005)

    def start(self):
        if self._h is not None:
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._h is not None and self._t.is_alive():
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def _note(msg):
    """Progress lines on stderr (stdout carries exactly one JSON line): a stuck multi-GPU run shows where it stopped."""
    sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:7.2f}s] {msg}\n")
    sys.stderr.flush()


_T0 = time.perf_counter()


def run_product(args):
    import faulthandler

    import torch
    import torch.distributed as dist

    faulthandler.dump_traceback_later(args.watchdog_s, exit=True)  # a hung collective must not eat the GPU lease

    import geometrics_b200 as g
    from geometrics_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Libraries (NCCL's version banner, for one) write to stdout; the contract is ONE JSON line there, so
    # everything before the final print goes to stderr.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    L = g.lib()
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    # ---- inputs: a ring of distinct sets, resident in HBM.  strong scaling: rank r holds items [r*Bl, (r+1)*Bl) of the
    #      same global batch of 32; weak: 32 items of its own per rank ----
    global B
    B = args.global_batch
    strong = args.scaling == "strong"
    if strong and B % world:
        raise SystemExit(f"bench.py: --scaling strong needs a GPU count that divides the batch of {B}")
    Bl = B // world if strong else B
    ring = []
    fp32_path = args.path == "fp32"
    flags = _lib.GEOM_FLAG_FP32 if fp32_path else 0
    # the tensor-core filtered search is the library's choice whenever the workspace is large enough for it
    ws_bytes = (L.geom_chamfer_workspace_bytes if fp32_path else L.geom_chamfer_tensor_workspace_bytes)(Bl, N_PTS, M_PTS)
    lws_bytes = L.geom_chamfer_loss_workspace_bytes(Bl, N_PTS, M_PTS)
    for r in range(RING):
        if strong:
            x1, x2 = _host_inputs(SEED + 7 * r)
            x1, x2 = x1[rank * Bl:(rank + 1) * Bl].contiguous(), x2[rank * Bl:(rank + 1) * Bl].contiguous()
        else:
            x1, x2 = _host_inputs(SEED + 7 * r + 100000 * rank)
        ring.append(dict(
            x1=x1.to(dev), x2=x2.to(dev),
            i1=torch.empty(Bl, N_PTS, dtype=torch.int32, device=dev), i2=torch.empty(Bl, M_PTS, dtype=torch.int32, device=dev),
            ws=torch.empty((ws_bytes + 15) // 16 * 2, dtype=torch.int64, device=dev),
            lws=torch.empty(max(lws_bytes // 16, 1) * 2, dtype=torch.float64, device=dev),
            sums=torch.zeros(2, dtype=torch.float32, device=dev)))

    def step_kernels(s, sp_):
        # indices only, as chamfer_distance.py:31 returns (dist pointers NULL); then the loss sums on them
        st = L.geom_chamfer_fwd(Bl, N_PTS, s["x1"].data_ptr(), M_PTS, s["x2"].data_ptr(), None, s["i1"].data_ptr(),
                                None, s["i2"].data_ptr(), s["ws"].data_ptr(), ws_bytes, flags, sp_)
        if st != 0:
            _lib.check(st, "geom_chamfer_fwd")
        st = L.geom_chamfer_loss_fwd(Bl, N_PTS, s["x1"].data_ptr(), M_PTS, s["x2"].data_ptr(), s["i1"].data_ptr(),
                                     s["i2"].data_ptr(), None, None, s["sums"].data_ptr(), s["lws"].data_ptr(),
                                     s["lws"].numel() * 8, sp_)
        if st != 0:
            _lib.check(st, "geom_chamfer_loss_fwd")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # One CUDA graph per ring slot (search + loss: memsets, 2 pack launches, scan, 2 loss launches): at 4 items per GPU a
    # step is ~100 us of device time, less than the host needs to launch it kernel by kernel.
    _note(f"inputs resident: {RING} sets of {Bl} items per rank")
    graphs = None
    if not args.no_graph:
        for k in range(2):
            step_kernels(ring[k], sp)  # first-use work (shared-memory opt-in) must happen outside a capture
        torch.cuda.synchronize(dev)
        graphs = []
        for s in ring:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                step_kernels(s, torch.cuda.current_stream(dev).cuda_stream)
            graphs.append(gph)

    # NCCL comes up only now: graph capture and a live NCCL watchdog thread do not mix (a capture in progress forbids
    # the event queries the watchdog issues from its own thread)
    _note("graphs captured" if graphs is not None else "direct launches")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        _note("NCCL process group up")

    pending = []

    def step(k):
        if graphs is not None:
            graphs[k % RING].replay()
        else:
            step_kernels(ring[k % RING], sp)
        if world > 1:
            # the ONE collective of the path: all-reduce of the two partial loss sums, asynchronous -- it overlaps the
            # following searches and is waited for (stream-side) two steps later: its kernel needs an SM on every GPU
            # at about the same time, and the search keeps all of them busy until a wave of CTAs retires
            pending.append(dist.all_reduce(ring[k % RING]["sums"], async_op=True))
            if len(pending) > 2:
                pending.pop(0).wait()

    def drain():
        while pending:
            pending.pop(0).wait()

    for k in range(args.warmup):
        step(k)
    drain()
    barrier()
    _note("warm-up done")
    # Clocks of the timed region: read from THIS thread once every step is enqueued and the GPU is still working through
    # them (and every 5 ms after that until the region ends).  A sampling thread started beside the enqueue loop makes its
    # first NVML calls while the host is launching; with 8 ranks on one box that stalled the launches of a 2-ms region
    # for several ms (r02: 0.29 ms/step timed against 0.09 sustained).
    sampler = ClockSampler(_physical_gpu_index(local))
    sampler.sample()      # NVML paths warm before the region; this reading is dropped
    sampler.samples.clear()
    sampler.reasons.clear()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h0 = time.perf_counter()
    for k in range(args.steps):
        step(args.warmup + k)
    host_us = (time.perf_counter() - h0) / args.steps * 1e6  # host time to enqueue one step
    drain()
    e1.record(stream)
    sampler.sample()      # everything is enqueued, the device is inside the region
    while not e1.query():
        sampler.sample()
        time.sleep(0.005)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    _note(f"timed region done: {ms_total / args.steps:.4f} ms/step")

    # ---- the same step for >= 2 s: what the rate settles at once the part is under sustained load ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / max(ms_total / args.steps, 1e-3)) + 1)
        if world > 1:  # every rank must run the SAME number of steps (each step carries a collective): rank 0 decides
            nt = torch.tensor([n_sus], dtype=torch.int64, device=dev)
            dist.broadcast(nt, 0)
            n_sus = int(nt.item())
        sam2 = ClockSampler(_physical_gpu_index(local))
        sam2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for k in range(n_sus):
            step(k)
        drain()
        s1.record(stream)
        barrier()
        sustained = {"steps": n_sus, "ms_total": s0.elapsed_time(s1), "clocks": sam2.stop()}
        _note("sustained region done")

    # ---- second pass of the same K steps with the library's per-kernel event timing on (direct launches, no graph):
    #      average launch duration of each kernel, on the launch stream ----
    _lib.prof_begin()
    for k in range(args.steps):
        step_kernels(ring[(args.warmup + k) % RING], sp)
    kern = _lib.prof_end()  # {name: (total ms, launches)}
    barrier()
    dom_name = max(kern, key=lambda k_: kern[k_][0])
    kern_ms = kern[dom_name][0] / kern[dom_name][1]
    launches = sum(c for _, c in kern.values())
    _note("per-kernel pass done")

    # ---- end to end: pinned host buffers in, indices out, copies inside the timed region ----
    if strong:
        hx1, hx2 = _host_inputs(SEED)
        hx1, hx2 = hx1[rank * Bl:(rank + 1) * Bl].contiguous(), hx2[rank * Bl:(rank + 1) * Bl].contiguous()
    else:
        hx1, hx2 = _host_inputs(SEED + 100000 * rank)
    hx1, hx2 = hx1.pin_memory(), hx2.pin_memory()
    hi1 = torch.empty(Bl, N_PTS, dtype=torch.int32).pin_memory()
    hi2 = torch.empty(Bl, M_PTS, dtype=torch.int32).pin_memory()
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(2):
        g.chamfer_forward_host(hx1, hx2, hi1, hi2, device=dev, chunks=args.e2e_chunks, tensor=False if fp32_path else None)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # returns after the D2H copy completed
        g.chamfer_forward_host(hx1, hx2, hi1, hi2, device=dev, chunks=args.e2e_chunks, tensor=False if fp32_path else None)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    _note("end-to-end pass done")

    # ---- reduce over ranks (max time) ----
    t = torch.tensor([ms_total, e2e_s, kern_ms or 0.0, sustained["ms_total"] if sustained else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, kern_ms_max, sus_ms = (float(v) for v in t.tolist())
    ms_step = ms_total / args.steps
    pairs_step = Bl * N_PTS * M_PTS * world

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[dom_name]
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        B_ = Bl  # the dominant kernel's launch covers this rank's shard
        alg_bytes = 16 * B_ * (N_PTS + M_PTS)  # read 12 B/point, write one int32 index/point
        dom_ms = kern_ms_max
        ach = alg_bytes / (dom_ms * 1e-3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or 1965
        hbm = {"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "algorithmic_bytes": alg_bytes,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}
        if "chamfer_tc" in dom_name:
            # tensor-core filtered search: one M128 N64 K16 bf16 MMA per (128 query rows x 64 candidates), both
            # directions, over the padded extents (rows to 512, candidates to 128); minima on the ALU pipe
            pad = lambda v, q: (v + q - 1) // q * q
            mma_flops = 2 * 16 * B_ * (pad(N_PTS, 512) * pad(M_PTS, 128) + pad(M_PTS, 512) * pad(N_PTS, 128))
            tc_peak = float(peaks.get("bf16_tflops", 1650.0))
            tc_ach = mma_flops / (dom_ms * 1e-3) / 1e12
            alu_peak = 148 * 64 * sm_mhz * 1e6    # ALU pipe: 16 lanes x 4 SMSP per SM
            alu_ops = B_ * N_PTS * M_PTS          # 3-input minimum: one lane-op per two approximate distances, two directions
            alg_flops = 16 * B_ * N_PTS * M_PTS  # 8 flop per distance evaluation, 2*B*n*m evaluations
            alg_ach = alg_flops / (dom_ms * 1e-3) / 1e12
            # `achieved` / `frac` follow SURVEY.md 8d's ALGORITHMIC flops; the flops the tensor core actually issues (K padded
            # 3 -> 16 for the 15-term bf16 split, extents padded to 512 / 128) are reported beside them as issued_*
            roofline = {"bound": "tensor", "achieved": alg_ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": alg_ach / tc_peak,
                        "issued_tflops": tc_ach, "issued_frac": tc_ach / tc_peak,
                        "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1650 TFLOP/s",
                        "kernel": dom_name, "kernel_ms": dom_ms, "issued_flops_per_launch": mma_flops,
                        "algorithmic_flops_per_launch": alg_flops,
                        "note": "achieved = SURVEY 8d algorithmic flops (8 per distance evaluation) / kernel time; issued_* = the "
                                "tcgen05.mma shapes actually launched (K=16 carries the 15-term split distance); the kernel is bound "
                                "by the TMEM accumulator round trip and the ALU-pipe minima, not by HBM (see DESIGN.md 4.1)",
                        "hbm": hbm,
                        "alu": {"lane_ops_per_launch": alu_ops, "achieved_tlops": alu_ops / (dom_ms * 1e-3) / 1e12,
                                "peak_tlops": alu_peak / 1e12, "frac": alu_ops / (dom_ms * 1e-3) / alu_peak,
                                "peak_source": f"148 SM x 64 ALU lanes x {sm_mhz} MHz (median SM clock during the run)"}}
        else:
            fp32_peak = 148 * 128 * sm_mhz * 1e6  # FP32 lane-ops/s at the clock observed under load
            lane_ops = 6 * B_ * N_PTS * M_PTS     # 3 sub + 1 mul + 2 fma per pair, each pair evaluated once
            roofline = dict(hbm, bound="hbm", traffic=traffic, traffic_source=traffic_src, kernel=dom_name, kernel_ms=dom_ms,
                            note="brute-force NN is FP32-issue bound, not HBM bound (SURVEY.md 8d); see fp32 below",
                            fp32={"lane_ops_per_launch": lane_ops, "achieved_tlops": lane_ops / (dom_ms * 1e-3) / 1e12,
                                  "peak_tlops": fp32_peak / 1e12, "frac": lane_ops / (dom_ms * 1e-3) / fp32_peak,
                                  "peak_source": f"148 SM x 128 FP32 lanes x {sm_mhz} MHz (median SM clock during the run)"})
        cfg = _config(world, args.scaling)
        search_path = "CUDA-core packed-FP32 brute force (GEOM_FLAG_FP32)" if fp32_path else \
            "tensor-core filtered exact search (tcgen05 bf16 filter + exact fp32 re-evaluation; library default)"
        out = {
            "metric": METRIC, "value": pairs_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "search_path": search_path, "launch": "direct" if graphs is None else "CUDA graph per step (search + loss), NCCL all-reduce outside the graph",
            "clocks": clocks, "host_enqueue_us_per_step": host_us,
            "e2e": {"value": pairs_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": hx1.numel() * 4 + hx2.numel() * 4,
                    "d2h_bytes_per_step": hi1.numel() * 4 + hi2.numel() * 4, "ms_per_step": e2e_s * 1e3,
                    "api": "geom_chamfer_fwd_host via geometrics_b200.chamfer_forward_host (pinned host in/out, batch chunks pipelined)", "chunks": args.e2e_chunks, "steps": e2e_steps},
            "gpu_launches": launches,
            "kernels": {k_: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share_of_step": v[0] / args.steps / ms_step}
                        for k_, v in kern.items()},
            "roofline": roofline,
        }
        if traffic is not None:
            out["roofline"]["traffic_note"] = "DRAM bytes per launch from the committed ncu --set full capture named in traffic_source (a constant of that capture, not measured in this run)"
        if sustained:
            out["sustained"] = {"value": pairs_step / (sus_ms / sustained["steps"] * 1e-3), "unit": UNIT, "steps": sustained["steps"],
                                "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / sustained["steps"], "clocks": sustained["clocks"]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_block()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=5, help="batch chunks pipelined by geom_chamfer_fwd_host")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the same 32 items split over the GPUs (BASELINE metric); weak: 32 items per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of one CUDA graph per step")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="also time the step for at least this many seconds (0 = skip)")
    ap.add_argument("--global-batch", type=int, default=B, help="items in the global batch (the BASELINE config is 32)")
    ap.add_argument("--watchdog-s", type=float, default=240.0, help="dump every thread's stack and exit if the run takes longer")
    ap.add_argument("--path", default="auto", choices=["auto", "fp32"], help="auto = library default (tensor-core filtered search); fp32 = CUDA-core brute force")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "product" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
