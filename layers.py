"""Import shim: ``from layers import *`` (reference ``models.py:4``, ``GEOMetrics.py``) resolves to the
B200 implementation when this repository precedes the reference on ``sys.path``."""
from geometrics_b200.layers import *  # noqa: F401,F403
from geometrics_b200.layers import (BatchGCNMax, BatchZERON_GCN, Batch_Image_ZERON_GCNGCN, GCNMax,  # noqa: F401
                                    Image_ZERON_GCNGCN, ZERON_GCN, ZeroNGCN)
