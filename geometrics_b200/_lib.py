"""ctypes binding of ``libgeom_b200.so`` (C ABI declared in ``include/geom_b200.h``).

There is deliberately no fallback: if the CUDA library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C geometrics_b200/csrc``) every operator raises
:class:`GeomLibraryError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeom_b200.so")

GEOM_ABI_VERSION = 1
GEOM_FLAG_NONE = 0
GEOM_FLAG_EXACT_TAIL = 1
GEOM_FLAG_ONEWAY = 2
GEOM_FLAG_TENSOR = 4
GEOM_FLAG_FP32 = 8
# geom_debug_set knobs (tests / tuning only)
GEOM_DEBUG_TC_EPS_FACTOR = 0
GEOM_DEBUG_TRI_PR = 1
GEOM_DEBUG_TRI_CHUNK = 2
GEOM_DEBUG_TRI_TC_EPS_FACTOR = 3
GEOM_DEBUG_CHAMFER_VARIANT = 4
GEOM_DEBUG_CHAMFER_STRIP = 5
GEOM_DEBUG_CHAMFER_TILE_COLS = 6
GEOM_DEBUG_TC_CHUNKS = 7

# every symbol include/geom_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u, _sz = C.c_void_p, C.c_int, C.c_uint, C.c_size_t
SYMBOLS = {
    "geom_abi_version": (_i, []),
    "geom_status_string": (C.c_char_p, [_i]),
    "geom_device_info": (_i, [C.POINTER(_i)] * 3),
    "geom_prof_begin": (_i, []),
    "geom_prof_end": (_i, [C.POINTER(C.c_float), C.POINTER(_i)]),
    "geom_prof_kernel_name": (C.c_char_p, [_i]),
    "geom_chamfer_workspace_bytes": (_sz, [_i, _i, _i]),
    "geom_chamfer_tensor_workspace_bytes": (_sz, [_i, _i, _i]),
    "geom_chamfer_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp]),
    "geom_host_ctx_create": (_i, [C.POINTER(_vp)]),
    "geom_host_ctx_destroy": (_i, [_vp]),
    "geom_chamfer_host_plan": (_i, [_i, _i, _i, _i, _u, _i, C.POINTER(_i)]),
    "geom_chamfer_fwd_host": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _u]),
    "geom_chamfer_loss_workspace_bytes": (_sz, [_i, _i, _i]),
    "geom_chamfer_loss_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "geom_chamfer_loss_f1_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _sz, _vp]),
    "geom_chamfer_loss_bwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_debug_set": (_i, [_i, C.c_float]),
    "geom_debug_tc_dump": (_i, [_i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "geom_tri_workspace_bytes": (_sz, [_i, _i]),
    "geom_tri_tensor_workspace_bytes": (_sz, [_i, _i, _i]),
    "geom_tri_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp]),
    "geom_tri_loss_workspace_bytes": (_sz, [_i, _i]),
    "geom_tri_loss_fwd": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "geom_tri_loss_bwd": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_face_areas": (_i, [_i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "geom_sample_fwd": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "geom_sample_bwd": (_i, [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_edge_loss_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp]),
    "geom_edge_loss_bwd": (_i, [_i, _i, _vp, _i, _vp, _vp, C.c_float, _vp, _vp]),
    "geom_project_fwd": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "geom_project_bwd": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_pool_fwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "geom_pool_bwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "geom_csr_count": (_i, [_i, _vp, _i, _vp, _vp, _vp]),
    "geom_csr_fill": (_i, [_i, _vp, _i, _vp, _vp, _vp, _vp]),
    "geom_csr_spmm": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_zeron_gcn_workspace_bytes": (_sz, [_i, _i, _i]),
    "geom_zeron_gcn_tensor_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "geom_zeron_gcn_fwd": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "geom_zeron_gcn_bwd_smooth": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geom_zeron_gcn_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "geom_zeron_gcn_bwd": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "geom_bn_relu_res_fwd": (_i, [_i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    "geom_bn_relu_res_bwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
}


class GeomLibraryError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the CUDA library; raise loudly when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GeomLibraryError(
                f"{LIB_PATH} not found: the sm_100a CUDA library is not built and geometrics_b200 has "
                "no CPU fallback. Run `make -C geometrics_b200/csrc` (or __graft_entry__.build())."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(L, name)
            except AttributeError as exc:  # pragma: no cover - stale build
                raise GeomLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from exc
            fn.restype = res
            fn.argtypes = args
        if L.geom_abi_version() != GEOM_ABI_VERSION:
            raise GeomLibraryError(f"{LIB_PATH} has ABI {L.geom_abi_version()}, expected {GEOM_ABI_VERSION}; rebuild it")
        _lib = L
    return _lib


GEOM_PROF_KERNELS = 26


def prof_begin() -> None:
    check(lib().geom_prof_begin(), "geom_prof_begin")


def prof_end() -> dict:
    """-> {kernel name: (total device ms, launches)} for kernels launched since prof_begin()."""
    ms = (C.c_float * GEOM_PROF_KERNELS)()
    cnt = (_i * GEOM_PROF_KERNELS)()
    check(lib().geom_prof_end(ms, cnt), "geom_prof_end")
    return {lib().geom_prof_kernel_name(k).decode(): (float(ms[k]), int(cnt[k]))
            for k in range(GEOM_PROF_KERNELS) if cnt[k]}


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().geom_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def require_cuda_f32(name: str, t, last=None):
    """Argument validation the reference never did: fp32, CUDA, optional last dim."""
    import torch

    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (geometrics_b200 has no CPU path); got device {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if last is not None and (t.dim() != 3 or t.shape[-1] != last):
        raise ValueError(f"{name} must have shape (B, N, {last}), got {tuple(t.shape)}")


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """``with on_device(dev):`` -- torch.cuda.device(dev) only when dev is not already current (the context manager costs
    ~10 us of host time per use, which adds up over the ~30 small launches of a deformation block)."""
    import torch

    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def current_stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
