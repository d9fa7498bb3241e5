"""ctypes loaders for the checkers -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* :class:`Oracle`   -- ``liboracle.so``: C restatements in chamfer_oracle.c / tri_oracle.c.
* :class:`RefCpu`   -- ``_ref/libref_cpu.so``: the reference's
  ``old_GEOMetrics/chamfer_distance/src/my_lib.c`` compiled unmodified (``nnsearch`` :4-26,
  ``nnd_forward`` :28-46, ``nnd_backward`` :49-111) against ``th_shim/TH/TH.h``.
* :class:`RefCuda`  -- ``_ref/libref_cuda.so``: the reference's ``chamfer_distance.cu`` and
  ``tri_distance.cu`` compiled unmodified for sm_100a, behind ``ref_cuda_shim.cu``.

All array arguments are numpy (CPU) except :class:`RefCuda`, which takes raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_CPU_SO = os.path.join(_HERE, "_ref", "libref_cpu.so")
_REF_CUDA_SO = os.path.join(_HERE, "_ref", "libref_cuda.so")

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build_oracle(ref: bool | None = None) -> None:
    """(Re)build liboracle.so with gcc; also the _ref libraries when /root/reference exists."""
    targets = ["all"]
    if ref is None:
        ref = os.path.isdir("/root/reference")
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def have_ref_cpu() -> bool:
    return os.path.exists(_REF_CPU_SO)


def have_ref_cuda() -> bool:
    return os.path.exists(_REF_CUDA_SO)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class Oracle:
    """C restatements (liboracle.so)."""

    def __init__(self):
        if not os.path.exists(_ORACLE_SO):
            build_oracle(ref=False)
        self.lib = L = C.CDLL(_ORACLE_SO)
        for name in ("oracle_chamfer_dir", "oracle_chamfer_dir_exact"):
            getattr(L, name).argtypes = [C.c_int, C.c_int, _fp, C.c_int, _fp, _fp, _ip]
            getattr(L, name).restype = None
        L.oracle_chamfer_fwd.argtypes = [C.c_int, C.c_int, _fp, C.c_int, _fp, _fp, _ip, _fp, _ip]
        L.oracle_chamfer_fwd.restype = None
        L.oracle_nnsearch.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _ip]
        L.oracle_nnsearch.restype = None
        L.oracle_chamfer_loss_sums.argtypes = [C.c_int, C.c_int, _fp, C.c_int, _fp, _ip, _ip,
                                               C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_chamfer_loss_sums.restype = None
        L.oracle_chamfer_loss_bwd.argtypes = [C.c_int, C.c_int, _fp, C.c_int, _fp, _ip, _ip,
                                              C.c_float, C.c_float, _fp, _fp]
        L.oracle_chamfer_loss_bwd.restype = None
        for name in ("oracle_tri_fwd", "oracle_tri_fwd_exact"):
            getattr(L, name).argtypes = [C.c_int, C.c_int, _fp, C.c_int, _fp, _fp, _fp, _fp, _ip, _ip]
            getattr(L, name).restype = None
        L.oracle_tri_pair.argtypes = [_fp, _fp, _fp, _fp, _ip]
        L.oracle_tri_pair.restype = C.c_float

    # -- Chamfer -------------------------------------------------------------------------
    def chamfer_dir(self, xyz, xyz2, exact=False):
        xyz, xyz2 = _f32(xyz), _f32(xyz2)
        b, n, _ = xyz.shape
        m = xyz2.shape[1]
        dist = np.empty((b, n), np.float32)
        idx = np.empty((b, n), np.int32)
        fn = self.lib.oracle_chamfer_dir_exact if exact else self.lib.oracle_chamfer_dir
        fn(b, n, _p(xyz, _fp), m, _p(xyz2, _fp), _p(dist, _fp), _p(idx, _ip))
        return dist, idx

    def chamfer_fwd(self, xyz1, xyz2, exact=False):
        """-> dist1 (B,n), idx1 (B,n), dist2 (B,m), idx2 (B,m) as chamfer_distance.cu:57-68."""
        d1, i1 = self.chamfer_dir(xyz1, xyz2, exact)
        d2, i2 = self.chamfer_dir(xyz2, xyz1, exact)
        return d1, i1, d2, i2

    def nnsearch(self, xyz1, xyz2):
        xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dist = np.empty((b, n), np.float32)
        idx = np.empty((b, n), np.int32)
        self.lib.oracle_nnsearch(b, n, m, _p(xyz1, _fp), _p(xyz2, _fp), _p(dist, _fp), _p(idx, _ip))
        return dist, idx

    def chamfer_loss_sums(self, xyz1, xyz2, idx1, idx2):
        xyz1, xyz2, idx1, idx2 = _f32(xyz1), _f32(xyz2), _i32(idx1), _i32(idx2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        s1, s2 = C.c_double(), C.c_double()
        self.lib.oracle_chamfer_loss_sums(b, n, _p(xyz1, _fp), m, _p(xyz2, _fp), _p(idx1, _ip),
                                          _p(idx2, _ip), C.byref(s1), C.byref(s2))
        return s1.value, s2.value

    def chamfer_loss_bwd(self, xyz1, xyz2, idx1, idx2, w1, w2):
        xyz1, xyz2, idx1, idx2 = _f32(xyz1), _f32(xyz2), _i32(idx1), _i32(idx2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        g1, g2 = np.empty_like(xyz1), np.empty_like(xyz2)
        self.lib.oracle_chamfer_loss_bwd(b, n, _p(xyz1, _fp), m, _p(xyz2, _fp), _p(idx1, _ip),
                                         _p(idx2, _ip), w1, w2, _p(g1, _fp), _p(g2, _fp))
        return g1, g2

    # -- TriDistance ---------------------------------------------------------------------
    def tri_fwd(self, xyz, tri1, tri2, tri3, exact=False):
        """-> dist (B,n) f32, point (B,n) i32 in 0..6, index (B,n) i32 (tri_distance.cu:94-228)."""
        xyz, tri1, tri2, tri3 = _f32(xyz), _f32(tri1), _f32(tri2), _f32(tri3)
        b, n, _ = xyz.shape
        m = tri1.shape[1]
        dist = np.empty((b, n), np.float32)
        point = np.empty((b, n), np.int32)
        index = np.empty((b, n), np.int32)
        fn = self.lib.oracle_tri_fwd_exact if exact else self.lib.oracle_tri_fwd
        fn(b, n, _p(xyz, _fp), m, _p(tri1, _fp), _p(tri2, _fp), _p(tri3, _fp),
           _p(dist, _fp), _p(point, _ip), _p(index, _ip))
        return dist, point, index

    def tri_pair(self, p, a, b, c):
        p, a, b, c = _f32(p), _f32(a), _f32(b), _f32(c)
        opt = C.c_int()
        d = self.lib.oracle_tri_pair(_p(p, _fp), _p(a, _fp), _p(b, _fp), _p(c, _fp), C.byref(opt))
        return float(d), opt.value


class _THFloat(C.Structure):
    _fields_ = [("size", C.c_long * 4), ("data", _fp)]


class _THInt(C.Structure):
    _fields_ = [("size", C.c_long * 4), ("data", _ip)]


def _thf(a: np.ndarray) -> _THFloat:
    t = _THFloat()
    for k, s in enumerate(a.shape):
        t.size[k] = s
    t.data = _p(a, _fp)
    return t


def _thi(a: np.ndarray) -> _THInt:
    t = _THInt()
    for k, s in enumerate(a.shape):
        t.size[k] = s
    t.data = _p(a, _ip)
    return t


class RefCpu:
    """The reference's CPU code, unmodified (my_lib.c)."""

    def __init__(self):
        if not have_ref_cpu():
            raise FileNotFoundError(_REF_CPU_SO + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = L = C.CDLL(_REF_CPU_SO)
        L.nnsearch.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _ip]
        L.nnsearch.restype = None
        L.nnd_forward.argtypes = [C.POINTER(_THFloat)] * 4 + [C.POINTER(_THInt)] * 2
        L.nnd_forward.restype = C.c_int
        L.nnd_backward.argtypes = [C.POINTER(_THFloat)] * 6 + [C.POINTER(_THInt)] * 2
        L.nnd_backward.restype = C.c_int

    def nnsearch(self, xyz1, xyz2, dist=None, idx=None):
        xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        if dist is None:
            dist = np.empty((b, n), np.float32)
            idx = np.empty((b, n), np.int32)
        self.lib.nnsearch(b, n, m, _p(xyz1, _fp), _p(xyz2, _fp), _p(dist, _fp), _p(idx, _ip))
        return dist, idx

    def nnd_forward(self, xyz1, xyz2):
        xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        d1, d2 = np.empty((b, n), np.float32), np.empty((b, m), np.float32)
        i1, i2 = np.empty((b, n), np.int32), np.empty((b, m), np.int32)
        self.lib.nnd_forward(C.byref(_thf(xyz1)), C.byref(_thf(xyz2)), C.byref(_thf(d1)),
                             C.byref(_thf(d2)), C.byref(_thi(i1)), C.byref(_thi(i2)))
        return d1, i1, d2, i2

    def nnd_backward(self, xyz1, xyz2, graddist1, graddist2, idx1, idx2):
        xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
        gd1, gd2, idx1, idx2 = _f32(graddist1), _f32(graddist2), _i32(idx1), _i32(idx2)
        g1, g2 = np.empty_like(xyz1), np.empty_like(xyz2)
        self.lib.nnd_backward(C.byref(_thf(xyz1)), C.byref(_thf(xyz2)), C.byref(_thf(g1)),
                              C.byref(_thf(g2)), C.byref(_thf(gd1)), C.byref(_thf(gd2)),
                              C.byref(_thi(idx1)), C.byref(_thi(idx2)))
        return g1, g2


class RefCuda:
    """The reference's GPU kernels, unmodified, behind oracle/ref_cuda_shim.cu (device pointers)."""

    def __init__(self):
        if not have_ref_cuda():
            raise FileNotFoundError(_REF_CUDA_SO + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = L = C.CDLL(_REF_CUDA_SO)
        vp = C.c_void_p
        L.ref_chamfer_forward.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int]
        L.ref_chamfer_forward.restype = C.c_int
        L.ref_tri_forward.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        L.ref_tri_forward.restype = C.c_int

    def chamfer_forward_t(self, xyz1, xyz2, sync=True):
        """torch CUDA tensors in -> (dist1, idx1, dist2, idx2) torch CUDA tensors."""
        import torch

        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        d1 = torch.zeros(b, n, device=xyz1.device)
        d2 = torch.zeros(b, m, device=xyz1.device)
        i1 = torch.zeros(b, n, dtype=torch.int32, device=xyz1.device)
        i2 = torch.zeros(b, m, dtype=torch.int32, device=xyz1.device)
        torch.cuda.current_stream().synchronize()
        rc = self.lib.ref_chamfer_forward(b, n, xyz1.data_ptr(), m, xyz2.data_ptr(), d1.data_ptr(),
                                          i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), int(sync))
        if rc:
            raise RuntimeError(f"reference chamfer kernel failed: cudaError {rc}")
        return d1, i1, d2, i2

    def tri_forward_t(self, xyz, tri1, tri2, tri3, sync=True):
        import torch

        xyz, tri1, tri2, tri3 = (t.contiguous() for t in (xyz, tri1, tri2, tri3))
        b, n, _ = xyz.shape
        m = tri1.shape[1]
        dist = torch.zeros(b, n, device=xyz.device)
        point = torch.zeros(b, n, dtype=torch.int32, device=xyz.device)
        index = torch.zeros(b, n, dtype=torch.int32, device=xyz.device)
        torch.cuda.current_stream().synchronize()
        rc = self.lib.ref_tri_forward(b, n, xyz.data_ptr(), m, tri1.data_ptr(), tri2.data_ptr(),
                                      tri3.data_ptr(), dist.data_ptr(), point.data_ptr(),
                                      index.data_ptr(), int(sync))
        if rc:
            raise RuntimeError(f"reference tri kernel failed: cudaError {rc}")
        return dist, point, index
