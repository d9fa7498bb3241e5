"""Import shim: ``from tri_distance import TriDistance`` (reference ``utils.py:34``) resolves to the
B200 implementation when this repository precedes the reference on ``sys.path``."""
from geometrics_b200.tri_distance import TriDistance, TriDistanceFunction  # noqa: F401
