"""geometrics_b200 -- B200 (sm_100a) implementation of the GEOMetrics geometric hot path.

Host side mirrors the reference's Python call surface; compute is hand-written CUDA behind the C ABI
in ``include/geom_b200.h`` (``libgeom_b200.so``).  No CPU fallback exists.
"""
from ._lib import GeomLibraryError, LIB_PATH, lib  # noqa: F401
from .chamfer_distance import (ChamferDistance, ChamferDistanceFunction, chamfer_forward,  # noqa: F401
                               chamfer_forward_host)
from .tri_distance import TriDistance, TriDistanceFunction, tri_forward  # noqa: F401

from . import layers, models, utils  # noqa: F401,E402
from .layers import (BatchGCNMax, BatchZERON_GCN, Batch_Image_ZERON_GCNGCN, GCNMax, Image_ZERON_GCNGCN,  # noqa: F401,E402
                     ZERON_GCN, ZeroNGCN, csr_from_dense, zeron_gcn)
from .models import BatchMeshDeformationBlock, BatchMeshEncoder, MeshDeformationBlock, MeshEncoder  # noqa: F401,E402
from .utils import (adj_init, batch_point_to_point, batch_point_to_surface, batch_sample, calc_adj,  # noqa: F401,E402
                    face_areas, normalize_adj)

__all__ = [
    "ZERON_GCN", "BatchZERON_GCN", "Batch_Image_ZERON_GCNGCN", "Image_ZERON_GCNGCN", "GCNMax", "BatchGCNMax", "ZeroNGCN",
    "zeron_gcn", "csr_from_dense", "BatchMeshDeformationBlock", "MeshDeformationBlock", "BatchMeshEncoder", "MeshEncoder",
    "batch_sample", "batch_point_to_point", "batch_point_to_surface", "face_areas", "calc_adj", "normalize_adj", "adj_init",
    "ChamferDistance", "ChamferDistanceFunction", "chamfer_forward", "chamfer_forward_host",
    "TriDistance", "TriDistanceFunction", "tri_forward",
    "GeomLibraryError", "LIB_PATH", "lib",
]
