"""Drop-in for the reference's ``chamfer_distance.ChamferDistance``.

Same call surface as ``chamfer_distance/chamfer_distance.py:9-38`` of the reference:
``ChamferDistance()(xyz1, xyz2) -> (idx1, idx2)`` with ``idx1`` int32 (B,n) = nearest xyz2 point of
every xyz1 point and ``idx2`` int32 (B,m) the converse, both on the inputs' CUDA device.  The search
itself is ``geom_chamfer_fwd`` in ``libgeom_b200.so`` (sm_100a); outputs are allocated on the device
(the reference built them on the CPU and copied, :16-26) and the kernels run on torch's current
stream (the reference used the legacy default stream).
"""
from __future__ import annotations

import torch

from . import _lib


def chamfer_forward(xyz1: torch.Tensor, xyz2: torch.Tensor, *, exact_tail: bool = False,
                    oneway: bool = False, want_dist: bool = True):
    """-> (dist1, idx1, dist2, idx2); squared distances are fp32 (B,n)/(B,m) or None."""
    _lib.require_cuda_f32("xyz1", xyz1, 3)
    _lib.require_cuda_f32("xyz2", xyz2, 3)
    if xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError(f"batch mismatch: xyz1 {tuple(xyz1.shape)} vs xyz2 {tuple(xyz2.shape)}")
    if xyz1.device != xyz2.device:
        raise RuntimeError(f"xyz1 on {xyz1.device} but xyz2 on {xyz2.device}")
    L = _lib.lib()
    xyz1 = xyz1.detach().contiguous()  # the reference also calls .contiguous() (:14-15)
    xyz2 = xyz2.detach().contiguous()
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    with torch.cuda.device(dev):
        idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        dist1 = torch.empty((b, n), dtype=torch.float32, device=dev) if want_dist else None
        dist2 = torch.empty((b, m), dtype=torch.float32, device=dev) if want_dist else None
        ws_bytes = L.geom_chamfer_workspace_bytes(b, n, m)
        ws = torch.empty((max(ws_bytes, 8) + 7) // 8, dtype=torch.int64, device=dev)
        flags = (_lib.GEOM_FLAG_EXACT_TAIL if exact_tail else 0) | (_lib.GEOM_FLAG_ONEWAY if oneway else 0)
        st = L.geom_chamfer_fwd(
            b, n, xyz1.data_ptr(), m, xyz2.data_ptr(),
            dist1.data_ptr() if want_dist else None, idx1.data_ptr(),
            dist2.data_ptr() if want_dist else None, idx2.data_ptr(),
            ws.data_ptr(), ws.numel() * 8, flags, _lib.current_stream_ptr(dev))
        _lib.check(st, "geom_chamfer_fwd")
    return dist1, idx1, dist2, idx2


class ChamferDistanceFunction(torch.autograd.Function):
    """Forward-only, like the reference (:9-31): integer outputs carry no gradient."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        _, idx1, _, idx2 = chamfer_forward(xyz1, xyz2, want_dist=False)
        ctx.mark_non_differentiable(idx1, idx2)
        return idx1, idx2

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover - outputs are non-differentiable
        return None, None


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)


# ------------------------------------------------------------------------------------------------
# Host-buffer entry: pinned host point sets in, indices out.
# ------------------------------------------------------------------------------------------------
_host_state = {}


def _host_pipe(dev, b, n, m, chunks):
    key = (dev.index, b, n, m, chunks)
    st = _host_state.get(key)
    if st is None:
        L = _lib.lib()
        per = (b + chunks - 1) // chunks
        st = dict(
            per=per,
            h2d=torch.cuda.Stream(dev), comp=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev),
            x1=torch.empty((b, n, 3), dtype=torch.float32, device=dev),
            x2=torch.empty((b, m, 3), dtype=torch.float32, device=dev),
            i1=torch.empty((b, n), dtype=torch.int32, device=dev),
            i2=torch.empty((b, m), dtype=torch.int32, device=dev),
            ws_bytes=L.geom_chamfer_workspace_bytes(per, n, m),
            ev_in=[torch.cuda.Event() for _ in range(chunks)],
            ev_out=[torch.cuda.Event() for _ in range(chunks)],
        )
        st["ws"] = [torch.empty(max(st["ws_bytes"] // 8, 1), dtype=torch.int64, device=dev) for _ in range(chunks)]
        _host_state.clear()  # keep one shape resident
        _host_state[key] = st
    return st


def chamfer_forward_host(xyz1: torch.Tensor, xyz2: torch.Tensor, idx1_out: torch.Tensor, idx2_out: torch.Tensor,
                         device=None, chunks: int = 4, exact_tail: bool = False):
    """ChamferDistance on HOST tensors: ``xyz1`` (B,n,3) / ``xyz2`` (B,m,3) fp32 CPU tensors (pinned
    memory for full copy/compute overlap) -> ``idx1_out`` (B,n) / ``idx2_out`` (B,m) int32 CPU tensors.

    The batch is cut into ``chunks`` slices; slice c+1 is copied to the device while slice c is
    searched and slice c-1's indices are copied back (three streams).  Returns when every index is in
    host memory.  This is the end-to-end path bench.py times (``e2e``)."""
    if xyz1.is_cuda or xyz2.is_cuda or idx1_out.is_cuda or idx2_out.is_cuda:
        raise RuntimeError("chamfer_forward_host takes CPU tensors; use chamfer_forward for device tensors")
    for name, t in (("xyz1", xyz1), ("xyz2", xyz2)):
        if t.dtype != torch.float32 or t.dim() != 3 or t.shape[-1] != 3 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous float32 (B,N,3) tensor")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz2.shape[0] != b:
        raise ValueError("batch mismatch")
    for name, t, cnt in (("idx1_out", idx1_out, n), ("idx2_out", idx2_out, m)):
        if t.dtype != torch.int32 or tuple(t.shape) != (b, cnt) or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous int32 ({b},{cnt}) tensor")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if b == 0:
        return idx1_out, idx2_out
    L = _lib.lib()
    chunks = max(1, min(chunks, b))
    st = _host_pipe(dev, b, n, m, chunks)
    per = st["per"]
    flags = _lib.GEOM_FLAG_EXACT_TAIL if exact_tail else 0
    with torch.cuda.device(dev):
        for c in range(chunks):
            lo, hi = c * per, min(b, (c + 1) * per)
            if lo >= hi:
                break
            with torch.cuda.stream(st["h2d"]):
                st["x1"][lo:hi].copy_(xyz1[lo:hi], non_blocking=True)
                st["x2"][lo:hi].copy_(xyz2[lo:hi], non_blocking=True)
                st["ev_in"][c].record(st["h2d"])
            st["comp"].wait_event(st["ev_in"][c])
            rc = L.geom_chamfer_fwd(hi - lo, n, st["x1"][lo:hi].data_ptr(), m, st["x2"][lo:hi].data_ptr(),
                                    None, st["i1"][lo:hi].data_ptr(), None, st["i2"][lo:hi].data_ptr(),
                                    st["ws"][c].data_ptr(), st["ws_bytes"], flags, st["comp"].cuda_stream)
            _lib.check(rc, "geom_chamfer_fwd")
            st["ev_out"][c].record(st["comp"])
            st["d2h"].wait_event(st["ev_out"][c])
            with torch.cuda.stream(st["d2h"]):
                idx1_out[lo:hi].copy_(st["i1"][lo:hi], non_blocking=True)
                idx2_out[lo:hi].copy_(st["i2"][lo:hi], non_blocking=True)
        st["d2h"].synchronize()
    return idx1_out, idx2_out
