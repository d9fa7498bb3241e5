"""0N-GCN layers with the reference's class names, constructor arguments, parameter names / shapes
(state-dict compatible with the shipped checkpoints) and ``forward(input, adj, activation)``
signature (reference ``layers.py``).  ``adj`` is still the dense (V,V) tensor the reference's callers
hold (``adj_info['adj']``); it is converted once to CSR on the device and cached, and the forward is
``geom_zeron_gcn_fwd`` (feature GEMM with fused bias / concat + CSR SpMM over the first ``out//split``
columns) instead of two cuBLAS calls against the dense matrix.  CUDA only.
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.parameter import Parameter

from . import _lib

# ------------------------------------------------------------------------------------------------
# dense adjacency -> CSR (and CSR of the transpose for the backward), cached per adjacency tensor
# ------------------------------------------------------------------------------------------------
_csr_cache: dict = {}
_CSR_CACHE_MAX = 16


class CsrAdjacency:
    __slots__ = ("V", "nnz", "row_ptr", "col_idx", "val", "t_row_ptr", "t_col_idx", "t_val")


def _build_csr(adj: torch.Tensor, transpose: bool):
    L = _lib.lib()
    V = adj.shape[0]
    dev = adj.device
    sp = _lib.current_stream_ptr(dev)
    row_cnt = torch.empty(max(V, 1), dtype=torch.int32, device=dev)
    row_ptr = torch.empty(V + 1, dtype=torch.int32, device=dev)
    _lib.check(L.geom_csr_count(V, adj.data_ptr(), int(transpose), row_cnt.data_ptr(), row_ptr.data_ptr(), sp), "geom_csr_count")
    nnz = int(row_ptr[V].item())  # one-time setup sync per adjacency
    col_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    _lib.check(L.geom_csr_fill(V, adj.data_ptr(), int(transpose), row_ptr.data_ptr(), col_idx.data_ptr(), val.data_ptr(), sp), "geom_csr_fill")
    return nnz, row_ptr, col_idx, val


def csr_from_dense(adj: torch.Tensor) -> CsrAdjacency:
    """CSR (+ transposed CSR) of a dense (V,V) fp32 CUDA adjacency; cached on (storage pointer, version) for as
    long as the tensor object the entry was built from is alive (a dead owner means the allocator may have handed
    the same address to a different matrix, so the entry is dropped and rebuilt)."""
    if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
        raise ValueError(f"adj must be a square (V,V) matrix, got {tuple(adj.shape)}")
    if not adj.is_cuda:
        raise RuntimeError("adj must be a CUDA tensor (geometrics_b200 has no CPU path)")
    if adj.dtype != torch.float32:
        raise TypeError(f"adj must be float32, got {adj.dtype}")
    key = (adj.data_ptr(), adj._version, adj.shape[0], adj.device.index, adj.is_contiguous())
    hit = _csr_cache.get(key)
    if hit is not None:
        owner, cached = hit
        if owner() is not None:  # the owner keeps the storage alive: same pointer + version == same contents
            return cached
        del _csr_cache[key]
    a = adj.detach().contiguous()
    with _lib.on_device(a.device):
        c = CsrAdjacency()
        c.V = a.shape[0]
        c.nnz, c.row_ptr, c.col_idx, c.val = _build_csr(a, False)
        _, c.t_row_ptr, c.t_col_idx, c.t_val = _build_csr(a, True)
    if len(_csr_cache) >= _CSR_CACHE_MAX:
        _csr_cache.pop(next(iter(_csr_cache)))
    _csr_cache[key] = (weakref.ref(adj), c)
    return c


class CsrMatmulFunction(torch.autograd.Function):
    """Y = A X for the cached CSR of a dense (V,V) matrix and X (B,V,C); backward uses the CSR of A^T."""

    @staticmethod
    def forward(ctx, x, csr):
        L = _lib.lib()
        xc = x.detach().contiguous()
        b, v, c = xc.shape
        y = torch.empty_like(xc)
        with _lib.on_device(xc.device):
            _lib.check(L.geom_csr_spmm(b, v, c, csr.row_ptr.data_ptr(), csr.col_idx.data_ptr(), csr.val.data_ptr(),
                                       xc.data_ptr(), y.data_ptr(), _lib.current_stream_ptr(xc.device)), "geom_csr_spmm")
        ctx.csr = csr
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        csr = ctx.csr
        dy = dy.contiguous()
        b, v, c = dy.shape
        dx = torch.empty_like(dy)
        with _lib.on_device(dy.device):
            _lib.check(L.geom_csr_spmm(b, v, c, csr.t_row_ptr.data_ptr(), csr.t_col_idx.data_ptr(), csr.t_val.data_ptr(),
                                       dy.data_ptr(), dx.data_ptr(), _lib.current_stream_ptr(dy.device)), "geom_csr_spmm")
        return dx, None


def csr_matmul(adj, x):
    """``torch.matmul(adj, x)`` for a dense (V,V) ``adj`` (converted once to CSR) and x (B,V,C) or (V,C)."""
    _lib.require_cuda_f32("x", x)
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    csr = adj if isinstance(adj, CsrAdjacency) else csr_from_dense(adj)
    if csr.V != x.shape[1]:
        raise ValueError(f"matrix is {csr.V}x{csr.V} but x has {x.shape[1]} rows")
    y = CsrMatmulFunction.apply(x, csr)
    return y.squeeze(0) if squeeze else y


# ------------------------------------------------------------------------------------------------
# the fused layer
# ------------------------------------------------------------------------------------------------
class ZeroNGCNFunction(torch.autograd.Function):
    """Y = [adj . (XW)[:, :k] | (XW)[:, k:]] + bias for X (B,V,in), W (in,out), k = ksplit."""

    @staticmethod
    def forward(ctx, x, weight, bias, csr, ksplit, relu, tensor=False):
        L = _lib.lib()
        xc, wc = x.detach().contiguous(), weight.detach().contiguous()
        bc = bias.detach().contiguous() if bias is not None else None
        b, v, fin = xc.shape
        fout = wc.shape[1]
        dev = xc.device
        with _lib.on_device(dev):
            y = torch.empty((b, v, fout), dtype=torch.float32, device=dev)
            wsb = (L.geom_zeron_gcn_tensor_workspace_bytes(b, v, fin, fout, ksplit) if tensor
                   else L.geom_zeron_gcn_workspace_bytes(b, v, ksplit))
            ws = torch.empty(max(wsb // 4, 4), dtype=torch.float32, device=dev)
            _lib.check(L.geom_zeron_gcn_fwd(b, v, fin, fout, ksplit, xc.data_ptr(), wc.data_ptr(),
                                            bc.data_ptr() if bc is not None else None, csr.row_ptr.data_ptr(),
                                            csr.col_idx.data_ptr(), csr.val.data_ptr(), y.data_ptr(), int(relu) | (2 if tensor else 0),
                                            ws.data_ptr(), ws.numel() * 4, _lib.current_stream_ptr(dev)),
                       "geom_zeron_gcn_fwd")
        ctx.save_for_backward(xc, wc, y if relu else None)
        ctx.csr, ctx.ksplit, ctx.has_bias, ctx.tensor = csr, ksplit, bias is not None, bool(tensor)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, wc, y = ctx.saved_tensors
        csr, ksplit = ctx.csr, ctx.ksplit
        L = _lib.lib()
        dy = dy.contiguous()
        if y is not None:
            dy = dy * (y > 0)
        b, v, fout = dy.shape
        fin = xc.shape[2]
        dev = dy.device
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        with _lib.on_device(dev):
            dx = torch.empty_like(xc) if need_x else None
            dw = torch.empty_like(wc) if need_w else None
            db = torch.empty(fout, dtype=torch.float32, device=dev) if need_b else None
            wsb = L.geom_zeron_gcn_bwd_workspace_bytes(b, v, fin, fout)
            ws = torch.empty(wsb // 4 + 4, dtype=torch.float32, device=dev)
            # dS = adj^T-smoothed dY, db = column sums, dW = X^T dS, dX = dS W^T: all in this library (gcn_bwd.cu)
            _lib.check(L.geom_zeron_gcn_bwd(b, v, fin, fout, ksplit, xc.data_ptr(), wc.data_ptr(), dy.data_ptr(),
                                            csr.t_row_ptr.data_ptr(), csr.t_col_idx.data_ptr(), csr.t_val.data_ptr(),
                                            dx.data_ptr() if need_x else None, dw.data_ptr() if need_w else None,
                                            db.data_ptr() if need_b else None, 2 if ctx.tensor else 0,
                                            ws.data_ptr(), ws.numel() * 4, _lib.current_stream_ptr(dev)), "geom_zeron_gcn_bwd")
        return dx, dw, db, None, None, None, None


class BnReluResFunction(torch.autograd.Function):
    """out = relu(batch_norm_over_vertices(x)) or (res + that) / 2 -- ``geom_bn_relu_res_fwd / _bwd`` (bn.cu): the
    BatchNorm1d(V) + ReLU + averaged-residual stage of BatchMeshDeformationBlock (models.py:241-294) in one launch."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, running_mean, running_var, batches_tracked, training, momentum, eps):
        L = _lib.lib()
        xc = x.detach().contiguous()
        b, v, c = xc.shape
        dev = xc.device
        ld_res = 0
        if res is not None:
            if res.shape != xc.shape:
                raise ValueError(f"residual {tuple(res.shape)} does not match x {tuple(xc.shape)}")
            if res.stride(2) != 1 or res.stride(0) != v * res.stride(1):
                res = res.contiguous()
            ld_res = res.stride(1)
        with _lib.on_device(dev):
            out = torch.empty_like(xc)
            stats = torch.empty(2, v, dtype=torch.float32, device=dev)
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            _lib.check(L.geom_bn_relu_res_fwd(b, v, c, xc.data_ptr(), ptr(res), ld_res, ptr(gamma), ptr(beta), ptr(running_mean),
                                              ptr(running_var), ptr(batches_tracked), int(training), float(momentum), float(eps),
                                              stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(),
                                              _lib.current_stream_ptr(dev)), "geom_bn_relu_res_fwd")
        ctx.save_for_backward(xc, gamma, beta, stats)
        ctx.training, ctx.has_res = bool(training), res is not None
        return out  # (the running statistics are plain buffers updated in place by the kernel, as F.batch_norm does)

    @staticmethod
    def backward(ctx, dout):
        xc, gamma, beta, stats = ctx.saved_tensors
        L = _lib.lib()
        dout = dout.contiguous()
        b, v, c = xc.shape
        dev = xc.device
        need_res = ctx.has_res and ctx.needs_input_grad[1]
        with _lib.on_device(dev):
            dx = torch.empty_like(xc)
            dres = torch.empty_like(xc) if need_res else None
            dg = torch.empty(v, dtype=torch.float32, device=dev) if gamma is not None else None
            dbt = torch.empty(v, dtype=torch.float32, device=dev) if beta is not None else None
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            _lib.check(L.geom_bn_relu_res_bwd(b, v, c, xc.data_ptr(), dout.data_ptr(), ptr(gamma), ptr(beta), stats[0].data_ptr(),
                                              stats[1].data_ptr(), int(ctx.training), int(ctx.has_res), dx.data_ptr(), ptr(dres),
                                              ptr(dg), ptr(dbt), _lib.current_stream_ptr(dev)), "geom_bn_relu_res_bwd")
        return dx, dres, dg, dbt, None, None, None, None, None, None


_ws_cache: dict = {}


def _workspace(dev, nbytes: int, slot: str):
    """A reusable scratch buffer per (device, slot): the kernels only need it for the duration of the call they are
    enqueued by and every user runs on the current stream, so stream order makes reuse safe (a per-call torch.empty costs
    a few microseconds of host time, tens of times per deformation block)."""
    key = (dev.index, slot, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() * 4 < nbytes:
        buf = _ws_cache[key] = torch.empty(max(nbytes // 4 + 64, 1024), dtype=torch.float32, device=dev)
    return buf


class GcnBnStageFunction(torch.autograd.Function):
    """One stage of BatchMeshDeformationBlock (models.py:241-294) as ONE autograd node:
        y = [adj . (xW)[:, :k] | (xW)[:, k:]] + b;   out = relu(BatchNorm1d_V(y))  [; out = (res + out) / 2]
    forward = geom_zeron_gcn_fwd + geom_bn_relu_res_fwd, backward = geom_bn_relu_res_bwd + geom_zeron_gcn_bwd.  The same
    kernels as ZeroNGCNFunction / BnReluResFunction; merging the nodes halves the host-side cost of the block, which is
    launch-bound at the reference's shapes (B=16, V=482)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, res, csr, ksplit, tensor, running_mean, running_var, batches_tracked,
                training, momentum, eps):
        L = _lib.lib()
        xc, wc = x.detach().contiguous(), weight.detach().contiguous()
        b, v, fin = xc.shape
        fout = wc.shape[1]
        dev = xc.device
        ld_res = 0
        if res is not None:
            if tuple(res.shape) != (b, v, fout):
                raise ValueError(f"residual {tuple(res.shape)} does not match the stage output {(b, v, fout)}")
            if res.stride(2) != 1 or res.stride(0) != v * res.stride(1):
                res = res.contiguous()
            ld_res = res.stride(1)
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        with _lib.on_device(dev):
            sp = _lib.current_stream_ptr(dev)
            y = torch.empty((b, v, fout), dtype=torch.float32, device=dev)
            out = torch.empty_like(y)
            stats = torch.empty(2, v, dtype=torch.float32, device=dev)
            wsb = (L.geom_zeron_gcn_tensor_workspace_bytes(b, v, fin, fout, ksplit) if tensor
                   else L.geom_zeron_gcn_workspace_bytes(b, v, ksplit))
            ws = _workspace(dev, wsb, "gcn_fwd")
            _lib.check(L.geom_zeron_gcn_fwd(b, v, fin, fout, ksplit, xc.data_ptr(), wc.data_ptr(), ptr(bias), csr.row_ptr.data_ptr(),
                                            csr.col_idx.data_ptr(), csr.val.data_ptr(), y.data_ptr(), 2 if tensor else 0,
                                            ws.data_ptr(), ws.numel() * 4, sp), "geom_zeron_gcn_fwd")
            _lib.check(L.geom_bn_relu_res_fwd(b, v, fout, y.data_ptr(), ptr(res), ld_res, ptr(gamma), ptr(beta), ptr(running_mean),
                                              ptr(running_var), ptr(batches_tracked), int(training), float(momentum), float(eps),
                                              stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(), sp), "geom_bn_relu_res_fwd")
        ctx.save_for_backward(xc, wc, y, gamma, beta, stats)
        ctx.csr, ctx.ksplit, ctx.tensor, ctx.has_bias = csr, ksplit, bool(tensor), bias is not None
        ctx.training, ctx.has_res = bool(training), res is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, wc, y, gamma, beta, stats = ctx.saved_tensors
        csr, ksplit = ctx.csr, ctx.ksplit
        L = _lib.lib()
        dout = dout.contiguous()
        b, v, fin = xc.shape
        fout = wc.shape[1]
        dev = xc.device
        need_x, need_w, need_b, need_res = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2],
                                            ctx.has_res and ctx.needs_input_grad[5])
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        with _lib.on_device(dev):
            sp = _lib.current_stream_ptr(dev)
            dy = torch.empty_like(y)
            dres = torch.empty_like(y) if need_res else None
            dg = torch.empty(v, dtype=torch.float32, device=dev) if gamma is not None else None
            dbt = torch.empty(v, dtype=torch.float32, device=dev) if beta is not None else None
            _lib.check(L.geom_bn_relu_res_bwd(b, v, fout, y.data_ptr(), dout.data_ptr(), ptr(gamma), ptr(beta), stats[0].data_ptr(),
                                              stats[1].data_ptr(), int(ctx.training), int(ctx.has_res), dy.data_ptr(), ptr(dres),
                                              ptr(dg), ptr(dbt), sp), "geom_bn_relu_res_bwd")
            dx = torch.empty_like(xc) if need_x else None
            dw = torch.empty_like(wc) if need_w else None
            db = torch.empty(fout, dtype=torch.float32, device=dev) if need_b else None
            ws = _workspace(dev, L.geom_zeron_gcn_bwd_workspace_bytes(b, v, fin, fout), "gcn_bwd")
            _lib.check(L.geom_zeron_gcn_bwd(b, v, fin, fout, ksplit, xc.data_ptr(), wc.data_ptr(), dy.data_ptr(),
                                            csr.t_row_ptr.data_ptr(), csr.t_col_idx.data_ptr(), csr.t_val.data_ptr(), ptr(dx), ptr(dw),
                                            ptr(db), 2 if ctx.tensor else 0, ws.data_ptr(), ws.numel() * 4, sp), "geom_zeron_gcn_bwd")
        return dx, dw, db, dg, dbt, dres, None, None, None, None, None, None, None, None, None


def gcn_bn_stage(x, layer, bn: nn.BatchNorm1d, csr, res=None, tensor=None):
    """``F.relu(bn(layer(x, adj, identity)))`` [then ``(res + .) / 2``] for a Batch_Image_ZERON_GCNGCN ``layer`` and
    ``bn = nn.BatchNorm1d(V)`` as one autograd node (see GcnBnStageFunction); BatchNorm semantics as in bn_relu_res."""
    _lib.require_cuda_f32("input", x, None)
    weight = layer.weight1[0] if layer.weight1.dim() == 3 else layer.weight1
    if x.dim() != 3 or x.shape[1] != bn.num_features or x.shape[1] != csr.V or x.shape[2] != weight.shape[0]:
        raise ValueError(f"input {tuple(x.shape)} does not fit layer {tuple(weight.shape)} / BatchNorm1d({bn.num_features}) / adjacency {csr.V}")
    use_batch = bn.training or bn.running_mean is None
    momentum = bn.momentum
    if momentum is None:
        momentum = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.num_batches_tracked is not None else 0.0
    track = bn.training and bn.track_running_stats
    keep = track or not use_batch
    return GcnBnStageFunction.apply(x, weight, layer.bias, bn.weight, bn.bias, res, csr, weight.shape[1] // layer.split,
                                    TENSOR_GEMM if tensor is None else bool(tensor), bn.running_mean if keep else None,
                                    bn.running_var if keep else None, bn.num_batches_tracked if track else None, use_batch,
                                    momentum, bn.eps)


def bn_relu_res(x, bn: nn.BatchNorm1d, res=None):
    """``F.relu(bn(x))`` for x (B,V,C) and ``bn = nn.BatchNorm1d(V)`` [then ``(res + .) / 2``], fused.  The module only
    supplies its parameters and buffers (state-dict compatible); semantics follow torch's BatchNorm1d exactly: batch
    statistics in training mode (or when it tracks no running statistics), running statistics updated in place."""
    _lib.require_cuda_f32("x", x)
    if x.dim() != 3 or x.shape[1] != bn.num_features:
        raise ValueError(f"expected (B, {bn.num_features}, C), got {tuple(x.shape)}")
    use_batch = bn.training or bn.running_mean is None
    momentum = bn.momentum
    if momentum is None:  # cumulative moving average
        momentum = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.num_batches_tracked is not None else 0.0
    track = bn.training and bn.track_running_stats
    return BnReluResFunction.apply(x, res, bn.weight, bn.bias, bn.running_mean if (track or not use_batch) else None,
                                   bn.running_var if (track or not use_batch) else None,
                                   bn.num_batches_tracked if track else None, use_batch, momentum, bn.eps)


TENSOR_GEMM = True  # feature GEMM on the tensor cores (3xTF32 tcgen05) whenever the shape allows (out >= 16, <= 256)


def zeron_gcn(x, weight, bias, adj, split: int, activation=None, tensor=None):
    """Functional form.  ``x`` (B,V,in) or (V,in); ``adj`` dense (V,V) or a CsrAdjacency; the first
    ``out // split`` output columns are neighbour-averaged (split = 3: image GCN, 10: encoder GCN)."""
    _lib.require_cuda_f32("input", x)
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError(f"input must be (B,V,in) or (V,in), got {tuple(x.shape)}")
    if weight.dim() == 3:  # Batch_Image_ZERON_GCNGCN.weight1 is (1,in,out)
        weight = weight[0]
    csr = adj if isinstance(adj, CsrAdjacency) else csr_from_dense(adj)
    if csr.V != x.shape[1]:
        raise ValueError(f"adjacency is {csr.V}x{csr.V} but input has {x.shape[1]} vertices")
    for name, t in (("weight", weight), ("bias", bias)):
        if t is not None and (t.device != x.device or t.dtype != torch.float32):
            raise RuntimeError(f"{name} must be float32 on {x.device}, got {t.dtype} on {t.device}")
    if isinstance(adj, torch.Tensor) and adj.device != x.device:
        raise RuntimeError(f"adj on {adj.device} but input on {x.device}")
    if weight.shape[0] != x.shape[2]:
        raise ValueError(f"weight is {tuple(weight.shape)} but input has {x.shape[2]} features")
    fuse_relu = activation is F.relu or activation is torch.relu
    y = ZeroNGCNFunction.apply(x, weight, bias, csr, weight.shape[1] // split, fuse_relu,
                               TENSOR_GEMM if tensor is None else bool(tensor))
    if squeeze:
        y = y.squeeze(0)
    if activation is None or fuse_relu:
        return y
    return activation(y)


class ZERON_GCN(nn.Module):
    """layers.py:14-41: un-batched encoder layer, weight (in,out), split out//10."""

    split = 10

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.Tensor(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 6. / math.sqrt(self.weight.size(1) + self.weight.size(0))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-0, 0)

    def forward(self, input, adj, activation):
        return zeron_gcn(input, self.weight, self.bias, adj, self.split, activation)


class BatchZERON_GCN(ZERON_GCN):
    """layers.py:123-152: batched encoder layer (same parameters, input (B,V,in))."""


class Batch_Image_ZERON_GCNGCN(nn.Module):
    """layers.py:84-116: deformation-block layer, weight1 (1,in,out), split out//3."""

    split = 3

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight1 = Parameter(torch.Tensor(1, in_features, out_features))
        if bias:
            self.bias = Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 6. / math.sqrt((self.weight1.size(1) + self.weight1.size(0)))
        stdv *= .3
        self.weight1.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-.1, .1)

    def forward(self, input, adj, activation):
        return zeron_gcn(input, self.weight1, self.bias, adj, self.split, activation)


class GCNMax(nn.Module):
    """layers.py:43-79: encoder head; max over vertices of the ACTIVATED layer output (:76-78)."""

    pool_activated = True

    def __init__(self, in_features, print_length):
        super().__init__()
        self.in_features = in_features
        self.print_length = print_length
        self.weight_Ws = nn.ParameterList(Parameter(torch.Tensor(in_features, print_length)) for i in range(1))
        self.weight_Bs = nn.ParameterList(Parameter(torch.Tensor(print_length)) for i in range(1))
        self.reset_parameters()

    def reset_parameters(self):
        for i in range(1):
            stdv = 6. / math.sqrt(self.weight_Bs[i].size(0))
            self.weight_Bs[i].data.uniform_(-stdv, stdv)
            stdv = 6. / math.sqrt(self.weight_Ws[i].size(0) + self.weight_Ws[i].size(1))
            self.weight_Ws[i].data.uniform_(-stdv, stdv)

    def forward(self, r_s, adj, activation):
        v_s = zeron_gcn(r_s, self.weight_Ws[0], self.weight_Bs[0], adj, 10, None)
        pooled = activation(v_s) if self.pool_activated else v_s
        return torch.max(pooled, dim=-2)[0]


class BatchGCNMax(GCNMax):
    """layers.py:157-189: batched head; pools the PRE-activation tensor (:186-187), unlike GCNMax."""

    pool_activated = False


ZeroNGCN = ZERON_GCN  # the name BASELINE.json's north_star uses


class Image_ZERON_GCNGCN(nn.Module):
    """old_GEOMetrics/layers.py:106-141: the un-batched deformation-block layer that
    ``MeshDeformationBlock`` (models.py:138-201) instantiates; weight1 (in,out), split out//3.  The
    legacy caller passes the adjacency dict and the layer reads ``adj['adj']``; both forms work."""

    split = 3

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight1 = Parameter(torch.Tensor(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 6. / math.sqrt((self.weight1.size(1) + self.weight1.size(0)))
        stdv *= .6
        self.weight1.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-.1, .1)

    def forward(self, input, adj, activation):
        if isinstance(adj, dict):
            adj = adj["adj"]
        return zeron_gcn(input, self.weight1, self.bias, adj, self.split, activation)
