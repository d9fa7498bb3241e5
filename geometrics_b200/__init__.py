"""geometrics_b200 -- B200 (sm_100a) implementation of the GEOMetrics geometric hot path.

Host side mirrors the reference's Python call surface; compute is hand-written CUDA behind the C ABI
in ``include/geom_b200.h`` (``libgeom_b200.so``).  No CPU fallback exists.
"""
from ._lib import GeomLibraryError, LIB_PATH, lib  # noqa: F401
from .chamfer_distance import (ChamferDistance, ChamferDistanceFunction, chamfer_forward,  # noqa: F401
                               chamfer_forward_host)
from .tri_distance import TriDistance, TriDistanceFunction, tri_forward  # noqa: F401

__all__ = [
    "ChamferDistance", "ChamferDistanceFunction", "chamfer_forward", "chamfer_forward_host",
    "TriDistance", "TriDistanceFunction", "tri_forward",
    "GeomLibraryError", "LIB_PATH", "lib",
]
