"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/geom_b200.h declares; the Python surface mirrors the reference's names; and the product path
has no CPU fallback and never touches oracle/."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "geom_b200.h")).read()
    return sorted(set(re.findall(r"GEOM_API[^;]*?\b(geom_\w+)\s*\(", text)))


def test_header_constants_mirrored():
    """Numeric constants the ctypes layer duplicates must match the header."""
    import re
    from geometrics_b200 import _lib

    text = open(os.path.join(ROOT, "include", "geom_b200.h")).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (GEOM_[A-Z0-9_]+) (\d+)u?\b", text)}
    for name in ("GEOM_PROF_KERNELS", "GEOM_FLAG_EXACT_TAIL", "GEOM_FLAG_ONEWAY", "GEOM_FLAG_TENSOR", "GEOM_FLAG_FP32"):
        assert getattr(_lib, name) == consts[name], name


def test_header_symbols_exported():
    from geometrics_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no GEOM_API declarations parsed"
    assert sorted(_lib.SYMBOLS) == declared, "geometrics_b200/_lib.py SYMBOLS out of sync with the header"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in geom_b200.h but not exported"


def test_library_loads_and_reports_version():
    from geometrics_b200 import _lib

    L = _lib.lib()
    assert L.geom_abi_version() == _lib.GEOM_ABI_VERSION
    assert L.geom_status_string(0) == b"ok"
    assert b"workspace" in L.geom_status_string(-2)
    assert L.geom_chamfer_workspace_bytes(32, 10000, 10000) == 8 * 32 * 20000 + 16


def test_reference_call_surface_names():
    import chamfer_distance
    import tri_distance
    import geometrics_b200 as g

    assert chamfer_distance.ChamferDistance is g.ChamferDistance
    assert tri_distance.TriDistance is g.TriDistance
    import torch

    assert issubclass(g.ChamferDistance, torch.nn.Module)
    assert issubclass(g.TriDistanceFunction, torch.autograd.Function)


def test_no_cpu_fallback():
    import torch
    import geometrics_b200 as g

    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g.ChamferDistance()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g.TriDistance()(torch.zeros(1, 4, 3), *(torch.zeros(1, 2, 3),) * 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "geometrics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle"
                assert "liboracle" not in text and "libref_" not in text, f"{f} references an oracle library"
