"""Drop-in for the reference's ``tri_distance.TriDistance``.

Same call surface as ``tri_distance/tri_distance.py:9-43`` of the reference:
``TriDistance()(xyz1, tri1, tri2, tri3) -> (dist, point, index)``, all (B,n) on the inputs' device:
squared distance fp32, region option int32 in 0..6, face index int32.  ``tri1/2/3`` are the per-face
vertices A/B/C (B,m,3) the caller gathers (``utils.py:467-469``).
"""
from __future__ import annotations

import torch

from . import _lib


def tri_forward(xyz1, tri1, tri2, tri3, *, exact_tail: bool = False, want_dist: bool = True, tensor=None):
    """``tensor``: None = the library picks (tensor-core filtered search from 2^24 pairs up), True / False force the
    tensor-core filtered search / the CUDA-core brute force; the results are bit-identical."""
    _lib.require_cuda_f32("xyz1", xyz1, 3)
    for name, t in (("tri1", tri1), ("tri2", tri2), ("tri3", tri3)):
        _lib.require_cuda_f32(name, t, 3)
        if t.shape != tri1.shape:
            raise ValueError(f"{name} shape {tuple(t.shape)} != tri1 shape {tuple(tri1.shape)}")
        if t.device != xyz1.device:
            raise RuntimeError(f"{name} on {t.device} but xyz1 on {xyz1.device}")
    if tri1.shape[0] != xyz1.shape[0]:
        raise ValueError(f"batch mismatch: xyz1 {tuple(xyz1.shape)} vs tri1 {tuple(tri1.shape)}")
    L = _lib.lib()
    xyz1, tri1, tri2, tri3 = (t.detach().contiguous() for t in (xyz1, tri1, tri2, tri3))  # :15-18
    b, n, _ = xyz1.shape
    m = tri1.shape[1]
    dev = xyz1.device
    with _lib.on_device(dev):
        dist = torch.empty((b, n), dtype=torch.float32, device=dev) if want_dist else None
        point = torch.empty((b, n), dtype=torch.int32, device=dev)
        index = torch.empty((b, n), dtype=torch.int32, device=dev)
        flags = _lib.GEOM_FLAG_EXACT_TAIL if exact_tail else 0
        if tensor is None:
            big = b * n * m >= (1 << 24)
            ws_bytes = L.geom_tri_tensor_workspace_bytes(b, n, m) if big else L.geom_tri_workspace_bytes(b, n)
        elif tensor:
            flags |= _lib.GEOM_FLAG_TENSOR
            ws_bytes = L.geom_tri_tensor_workspace_bytes(b, n, m)
        else:
            flags |= _lib.GEOM_FLAG_FP32
            ws_bytes = L.geom_tri_workspace_bytes(b, n)
        ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.int64, device=dev)  # the caching allocator aligns to 512 bytes
        st = L.geom_tri_fwd(b, n, xyz1.data_ptr(), m, tri1.data_ptr(), tri2.data_ptr(), tri3.data_ptr(),
                            dist.data_ptr() if want_dist else None, point.data_ptr(), index.data_ptr(),
                            ws.data_ptr(), ws.numel() * 8, flags, _lib.current_stream_ptr(dev))
        _lib.check(st, "geom_tri_fwd")
    return dist, point, index


class TriDistanceFunction(torch.autograd.Function):
    """Forward-only, like the reference (:9-36)."""

    @staticmethod
    def forward(ctx, xyz1, tri1, tri2, tri3):
        dist, point, index = tri_forward(xyz1, tri1, tri2, tri3)
        ctx.mark_non_differentiable(dist, point, index)
        return dist, point, index

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        return None, None, None, None


class TriDistance(torch.nn.Module):
    def forward(self, xyz1, tri1, tri2, tri3):
        return TriDistanceFunction.apply(xyz1, tri1, tri2, tri3)
