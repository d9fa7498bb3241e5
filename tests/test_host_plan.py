"""CPU: the slice plan of geom_chamfer_fwd_host (pure host logic inside the C-ABI library, no device access)."""
import ctypes as C

import pytest

from geometrics_b200 import _lib


def plan(b, n, m, chunks, flags=0, sms=148):
    lo = (C.c_int * 9)()
    k = _lib.lib().geom_chamfer_host_plan(b, n, m, chunks, flags, sms, lo)
    return [lo[i] for i in range(k + 1)]


def test_tensor_search_slices_are_whole_waves():
    # config 2: 40 CTAs per item (2 directions x 10240 / 512 rows), 148 SMs -> 3 items fill one wave, 7 two, 11 three
    assert plan(32, 10000, 10000, 5) == [0, 3, 10, 21, 32]
    assert plan(32, 10000, 10000, 3) == [0, 3, 10, 32]
    sizes = [b - a for a, b in zip(plan(64, 10000, 10000, 8), plan(64, 10000, 10000, 8)[1:])]
    assert sum(sizes) == 64 and sizes[:3] == [3, 7, 11] and all(s > 0 for s in sizes)


@pytest.mark.parametrize("b,n,m,chunks,flags", [(1, 100, 100, 4, 0), (5, 1500, 1030, 8, 0), (32, 10000, 10000, 0, 0), (7, 3000, 3000, 2, 0),
                                                (32, 10000, 10000, 4, _lib.GEOM_FLAG_FP32), (3, 50000, 50000, 8, 0), (9, 10, 10, 3, 0)])
def test_plan_is_a_partition(b, n, m, chunks, flags):
    lo = plan(b, n, m, chunks, flags)
    assert lo[0] == 0 and lo[-1] == b and 1 <= len(lo) - 1 <= 8
    assert all(y > x for x, y in zip(lo, lo[1:]))


def test_cuda_core_search_keeps_small_edge_slices():
    assert plan(32, 10000, 10000, 3, _lib.GEOM_FLAG_FP32) == [0, 4, 28, 32]
    assert plan(32, 10000, 10000, 1, _lib.GEOM_FLAG_FP32) == [0, 32]


def test_invalid_arguments():
    lo = (C.c_int * 9)()
    assert _lib.lib().geom_chamfer_host_plan(0, 10, 10, 4, 0, 148, lo) == 0
    assert _lib.lib().geom_chamfer_host_plan(4, 10, 10, 4, 0, 148, None) == 0
