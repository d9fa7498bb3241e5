"""Import shim: ``from chamfer_distance import ChamferDistance`` (reference ``utils.py:31``) resolves
to the B200 implementation when this repository precedes the reference on ``sys.path``."""
from geometrics_b200.chamfer_distance import ChamferDistance, ChamferDistanceFunction  # noqa: F401
