"""Mirror of the reference's `chamfer3D` extension module
(OSF/assets/cuda/chamfer3D/chamfer3D_cuda.cpp:32-35: forward, backward) on libhimo_b200.so."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

#: finest search-cell edge in metres used by forward(); <=0 selects the library default (0.25 m)
CELL_SIZE = float(os.environ.get("HIMO_NN_CELL", "0"))


def _chk(t, name, dtype, cols=None):
    _lib.require_cuda(t, name)
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous {dtype} tensor")
    if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
        raise RuntimeError(f"{name} must have shape [N,{cols}]")


def forward(pc0, pc1, dist0, dist1, idx0, idx1) -> int:
    """chamfer3D.forward (chamfer3D.cu:85-105): fills dist0/dist1/idx0/idx1 in place, returns 1."""
    _chk(pc0, "pc0", torch.float32, 3)
    _chk(pc1, "pc1", torch.float32, 3)
    _chk(dist0, "dist0", torch.float32)
    _chk(dist1, "dist1", torch.float32)
    _chk(idx0, "idx0", torch.int32)
    _chk(idx1, "idx1", torch.int32)
    n0, n1 = pc0.shape[0], pc1.shape[0]
    if dist0.numel() != n0 or idx0.numel() != n0 or dist1.numel() != n1 or idx1.numel() != n1:
        raise RuntimeError("output sizes do not match the clouds")
    L = _lib.lib()
    dev = pc0.device
    ws_bytes = L.himo_chamfer_workspace_bytes(n0, n1)
    with _lib.on_device(dev):
        ws = _lib.workspace.get(ws_bytes, dev)
        st = L.himo_chamfer_forward(_lib.ptr(pc0), n0, _lib.ptr(pc1), n1, _lib.ptr(dist0),
                                    _lib.ptr(dist1), _lib.ptr(idx0), _lib.ptr(idx1),
                                    ctypes.c_float(CELL_SIZE), _lib.ptr(ws),
                                    ctypes.c_size_t(ws.numel()), _lib.stream_ptr(dev))
    _lib.check(st, "chamfer3D.forward")
    return 1


def forward_radius(pc0, pc1, dist0, dist1, idx0, idx1, radius: float) -> int:
    """Radius-limited forward: exact nearest neighbour where dist^2 <= radius^2, else (1e20, -1).
    Equivalent to chamfer3D.forward followed by the reference's truncation masks
    (chamfer3D/__init__.py:64-82) for every point the mask keeps."""
    _chk(pc0, "pc0", torch.float32, 3)
    _chk(pc1, "pc1", torch.float32, 3)
    _chk(dist0, "dist0", torch.float32)
    _chk(dist1, "dist1", torch.float32)
    _chk(idx0, "idx0", torch.int32)
    _chk(idx1, "idx1", torch.int32)
    n0, n1 = pc0.shape[0], pc1.shape[0]
    if dist0.numel() != n0 or idx0.numel() != n0 or dist1.numel() != n1 or idx1.numel() != n1:
        raise RuntimeError("output sizes do not match the clouds")
    if not radius > 0:
        raise RuntimeError("radius must be positive")
    L = _lib.lib()
    dev = pc0.device
    ws_bytes = L.himo_chamfer_workspace_bytes(n0, n1)
    with _lib.on_device(dev):
        ws = _lib.workspace.get(ws_bytes, dev)
        st = L.himo_chamfer_forward_radius(_lib.ptr(pc0), n0, _lib.ptr(pc1), n1, _lib.ptr(dist0),
                                           _lib.ptr(dist1), _lib.ptr(idx0), _lib.ptr(idx1),
                                           ctypes.c_float(radius), _lib.ptr(ws),
                                           ctypes.c_size_t(ws.numel()), _lib.stream_ptr(dev))
    _lib.check(st, "chamfer3D.forward_radius")
    return 1


def backward(pc0, pc1, idx0, idx1, grad_dist0, grad_dist1, grad_pc0, grad_pc1) -> int:
    """chamfer3D.backward (chamfer3D.cu:131-154): accumulates into grad_pc0 / grad_pc1."""
    _chk(pc0, "pc0", torch.float32, 3)
    _chk(pc1, "pc1", torch.float32, 3)
    _chk(idx0, "idx0", torch.int32)
    _chk(idx1, "idx1", torch.int32)
    _chk(grad_dist0, "grad_dist0", torch.float32)
    _chk(grad_dist1, "grad_dist1", torch.float32)
    _chk(grad_pc0, "grad_pc0", torch.float32, 3)
    _chk(grad_pc1, "grad_pc1", torch.float32, 3)
    dev = pc0.device
    with _lib.on_device(dev):
        st = _lib.lib().himo_chamfer_backward(
            _lib.ptr(pc0), pc0.shape[0], _lib.ptr(pc1), pc1.shape[0], _lib.ptr(idx0), _lib.ptr(idx1),
            _lib.ptr(grad_dist0), _lib.ptr(grad_dist1), _lib.ptr(grad_pc0), _lib.ptr(grad_pc1),
            _lib.stream_ptr(dev))
    _lib.check(st, "chamfer3D.backward")
    return 1
