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
                    oneway: bool = False, want_dist: bool = True, tensor=None):
    """-> (dist1, idx1, dist2, idx2); squared distances are fp32 (B,n)/(B,m) or None.

    ``tensor``: None (default) lets the library choose (the tensor-core filtered search whenever the sizes
    allow), True / False force the tensor-core / the CUDA-core search.  All paths return identical bits."""
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
    with _lib.on_device(dev):
        idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        dist1 = torch.empty((b, n), dtype=torch.float32, device=dev) if want_dist else None
        dist2 = torch.empty((b, m), dtype=torch.float32, device=dev) if want_dist else None
        fp32_only = tensor is False or oneway
        ws_bytes = (L.geom_chamfer_workspace_bytes if fp32_only else L.geom_chamfer_tensor_workspace_bytes)(b, n, m)
        ws = torch.empty((max(ws_bytes, 16) + 15) // 16 * 2, dtype=torch.int64, device=dev)
        flags = (_lib.GEOM_FLAG_EXACT_TAIL if exact_tail else 0) | (_lib.GEOM_FLAG_ONEWAY if oneway else 0) \
            | (_lib.GEOM_FLAG_TENSOR if tensor else 0) | (_lib.GEOM_FLAG_FP32 if tensor is False else 0)
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
# Host-buffer entry: pinned host point sets in, indices out (geom_chamfer_fwd_host).
# ------------------------------------------------------------------------------------------------
_host_ctx = {}


def _ctx_for(dev: torch.device):
    """One geom_host_ctx (streams + reusable device arena) per (device, thread)."""
    import ctypes
    import threading

    key = (dev.index, threading.get_ident())
    ctx = _host_ctx.get(key)
    if ctx is None:
        handle = ctypes.c_void_p()
        with _lib.on_device(dev):
            _lib.check(_lib.lib().geom_host_ctx_create(ctypes.byref(handle)), "geom_host_ctx_create")
        ctx = _host_ctx[key] = handle
    return ctx


def chamfer_forward_host(xyz1: torch.Tensor, xyz2: torch.Tensor, idx1_out: torch.Tensor, idx2_out: torch.Tensor,
                         device=None, chunks: int = 4, exact_tail: bool = False, dist1_out=None, dist2_out=None,
                         tensor=None):
    """ChamferDistance on HOST tensors: ``xyz1`` (B,n,3) / ``xyz2`` (B,m,3) fp32 CPU tensors (pinned
    memory for full copy/compute overlap) -> ``idx1_out`` (B,n) / ``idx2_out`` (B,m) int32 CPU tensors
    (and optionally the squared distances).

    One call into the C ABI (``geom_chamfer_fwd_host``): the batch is cut into ``chunks`` slices; slice
    c+1 is copied to the device while slice c is searched and slice c-1's indices are copied back
    (three streams owned by the library).  Returns when every result is in host memory.  This is the
    end-to-end path bench.py times (``e2e``)."""
    tensors = [("xyz1", xyz1), ("xyz2", xyz2), ("idx1_out", idx1_out), ("idx2_out", idx2_out)]
    tensors += [(nm, t) for nm, t in (("dist1_out", dist1_out), ("dist2_out", dist2_out)) if t is not None]
    for name, t in tensors:
        if t.is_cuda:
            raise RuntimeError(f"{name}: chamfer_forward_host takes CPU tensors; use chamfer_forward for device tensors")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    for name, t in (("xyz1", xyz1), ("xyz2", xyz2)):
        if t.dtype != torch.float32 or t.dim() != 3 or t.shape[-1] != 3:
            raise ValueError(f"{name} must be a float32 (B,N,3) tensor")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz2.shape[0] != b:
        raise ValueError("batch mismatch")
    for name, t, cnt, dt in (("idx1_out", idx1_out, n, torch.int32), ("idx2_out", idx2_out, m, torch.int32),
                             ("dist1_out", dist1_out, n, torch.float32), ("dist2_out", dist2_out, m, torch.float32)):
        if t is not None and (t.dtype != dt or tuple(t.shape) != (b, cnt)):
            raise ValueError(f"{name} must be a {dt} ({b},{cnt}) tensor")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device (geometrics_b200 has no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = _lib.lib()
    with _lib.on_device(dev):
        rc = L.geom_chamfer_fwd_host(_ctx_for(dev), b, n, xyz1.data_ptr(), m, xyz2.data_ptr(),
                                     dist1_out.data_ptr() if dist1_out is not None else None, idx1_out.data_ptr(),
                                     dist2_out.data_ptr() if dist2_out is not None else None, idx2_out.data_ptr(),
                                     int(chunks), (_lib.GEOM_FLAG_EXACT_TAIL if exact_tail else 0)
                                     | (_lib.GEOM_FLAG_TENSOR if tensor else 0) | (_lib.GEOM_FLAG_FP32 if tensor is False else 0))
    _lib.check(rc, "geom_chamfer_fwd_host")
    return idx1_out, idx2_out
