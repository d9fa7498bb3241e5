"""Graph models of the hot path with the reference's class names, constructor arguments, attribute
names (``gc1..gc13, gc15, bn1..bn14``; ``h1..h11, reduce``) and forward signatures, so that the
shipped checkpoints (``checkpoint/pretrained/model_{1,2,3}``) load with ``load_state_dict``.
Reference: models.py:138-201 (MeshDeformationBlock), :203-297 (BatchMeshDeformationBlock),
:299-348 (MeshEncoder), :386-435 (BatchMeshEncoder).  The image encoder (VGG) and the voxel decoder
are stock cuDNN modules and are not part of this package."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .layers import (BatchGCNMax, BatchZERON_GCN, Batch_Image_ZERON_GCNGCN, GCNMax, Image_ZERON_GCNGCN,
                     ZERON_GCN, bn_relu_res, csr_from_dense, gcn_bn_stage)

_identity = None  # activation=None means identity in layers.zeron_gcn


def _as_csr(adj):
    """Convert once per forward; every layer of the stack then shares the CSR."""
    if isinstance(adj, dict):
        adj = adj["adj"]
    return csr_from_dense(adj) if isinstance(adj, torch.Tensor) else adj


class _DeformationStack(nn.Module):
    """gc1..gc13 + gc15 with seven averaged residual stages (models.py:156-201 / 237-297)."""

    layer_cls = None
    feature_dim = -1

    def _gcn(self, i, x, adj, res=None):
        """Layer i with its normalisation / activation; with ``res`` also the averaged residual (res + .) / 2."""
        raise NotImplementedError

    def forward(self, features, pooled, adj):
        adj = _as_csr(adj)
        full_features = torch.cat((features, pooled), dim=self.feature_dim)
        x = self._gcn(1, full_features, adj)
        features = self._gcn(2, x, adj, res=full_features[..., :self.hidden])   # models.py:246-248
        for i in (3, 5, 7, 9, 11):
            x = self._gcn(i, features, adj)
            features = self._gcn(i + 1, x, adj, res=features)                   # :255-256 ...
        features = self._gcn(13, features, adj, res=features)                   # :290-294
        coords = self.gc15(features, adj, _identity)
        return features, coords


class BatchMeshDeformationBlock(_DeformationStack):
    """models.py:203-297: [gc -> BatchNorm1d(verts) -> ReLU] stages on (B,V,C) tensors.  Note that
    BatchNorm1d(verts) normalises per VERTEX over (batch, channels), as in the reference."""

    def __init__(self, input_features, verts, hidden=192, output_features=3):
        super().__init__()
        dims = [input_features] + [hidden] * 13
        for i in range(1, 14):
            setattr(self, f"gc{i}", Batch_Image_ZERON_GCNGCN(dims[i - 1], hidden))
        self.gc15 = Batch_Image_ZERON_GCNGCN(hidden, output_features)
        self.hidden = hidden
        for i in range(1, 15):
            setattr(self, f"bn{i}", nn.BatchNorm1d(verts))

    FUSED_BN = True     # BatchNorm1d(V) + ReLU + residual average in one kernel (bn.cu); False: the torch modules
    FUSED_STAGE = True  # ... and layer + normalisation as ONE autograd node (layers.GcnBnStageFunction)

    def _gcn(self, i, x, adj, res=None):
        bn = getattr(self, f"bn{i}")
        # the fused kernels implement nn.BatchNorm1d itself; anything else (e.g. SyncBatchNorm after conversion for
        # multi-GPU training) keeps its own module semantics
        fused = self.FUSED_BN and type(bn) is nn.BatchNorm1d and x.is_cuda and x.dim() == 3
        if fused and self.FUSED_STAGE:
            return gcn_bn_stage(x, getattr(self, f"gc{i}"), bn, adj, res)
        x = getattr(self, f"gc{i}")(x, adj, _identity)
        if fused:
            return bn_relu_res(x, bn, res)
        x = F.relu(bn(x))
        return x if res is None else (res + x) / 2


class MeshDeformationBlock(_DeformationStack):
    """models.py:138-201 (un-batched, (V,C) tensors, ReLU fused into each layer).  The reference class
    names ``Image_ZERON_GCNGCN``, which only exists in old_GEOMetrics/layers.py:106; it is provided
    here so the block works."""

    def __init__(self, input_features, hidden=192, output_features=3):
        super().__init__()
        dims = [input_features] + [hidden] * 13
        for i in range(1, 14):
            setattr(self, f"gc{i}", Image_ZERON_GCNGCN(dims[i - 1], hidden))
        self.gc15 = Image_ZERON_GCNGCN(hidden, output_features)
        self.hidden = hidden

    def _gcn(self, i, x, adj, res=None):
        x = getattr(self, f"gc{i}")(x, adj, F.relu)
        return x if res is None else (res + x) / 2


_ENCODER_DIMS = [("h1", 3, 60), ("h21", 60, 60), ("h22", 60, 60), ("h23", 60, 60), ("h24", 60, 120),
                 ("h3", 120, 120), ("h4", 120, 120), ("h41", 120, 150), ("h5", 150, 200), ("h6", 200, 210),
                 ("h7", 210, 250), ("h8", 250, 300), ("h81", 300, 300), ("h9", 300, 300), ("h10", 300, 300),
                 ("h11", 300, 300)]


class _EncoderStack(nn.Module):
    layer_cls = None
    head_cls = None

    def __init__(self, latent_length):
        super().__init__()
        for name, fin, fout in _ENCODER_DIMS:
            setattr(self, name, self.layer_cls(fin, fout))
        self.reduce = self.head_cls(300, latent_length)

    def forward(self, positions, adj, play=False):
        adj = _as_csr(adj)
        features = positions
        for name, _, _ in _ENCODER_DIMS:
            features = getattr(self, name)(features, adj, F.elu)
        return self.reduce(features, adj, F.elu)


class MeshEncoder(_EncoderStack):
    """models.py:299-348: 16 ZERON_GCN (3 -> 300) + GCNMax -> latent."""

    layer_cls = ZERON_GCN
    head_cls = GCNMax


class BatchMeshEncoder(_EncoderStack):
    """models.py:386-435: batched form (BatchZERON_GCN + BatchGCNMax)."""

    layer_cls = BatchZERON_GCN
    head_cls = BatchGCNMax
