"""Mirror of the reference's `mmcv` extension module (the 4 pybind functions of
OSF/assets/cuda/mmcv/pybind.cpp:33-49) on top of libhimo_b200.so.

Same names, argument meaning and error behaviour, so the reference's own Python wrappers
(OSF/assets/cuda/mmcv/voxelize.py:12-29, scatter_points.py:12-29) load this module through
their `importlib.import_module('mmcv')` + hasattr check and run unmodified
(see himo_b200.dropin.install and INTEGRATION.md).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def dynamic_voxelize_forward(points: torch.Tensor, voxel_size: torch.Tensor,
                             coors_range: torch.Tensor, coors: torch.Tensor, NDim: int = 3) -> None:
    """mmcv.dynamic_voxelize_forward (voxelization.cpp:62-74): writes `coors` in place."""
    _lib.require_cuda(points, "points")
    _lib.require_cuda(coors, "coors")
    if NDim != 3 or coors.dtype != torch.int32 or coors.shape != (points.shape[0], 3):
        raise RuntimeError("coors must be int32 [N,3] (NDim=3)")
    if points.dtype != torch.float32:
        raise RuntimeError("points must be float32")
    pts = points if points.is_contiguous() else points.contiguous()
    if not coors.is_contiguous():
        raise RuntimeError("coors must be contiguous")
    vs, cr = _host_f32(voxel_size), _host_f32(coors_range)
    with _lib.on_device(pts.device):
        st = _lib.lib().himo_dynamic_voxelize_forward(
            _lib.ptr(pts), pts.shape[0], pts.shape[1], _lib.ptr(vs), _lib.ptr(cr),
            _lib.ptr(coors), _lib.stream_ptr(pts.device))
    _lib.check(st, "dynamic_voxelize_forward")


def _host_f32(t: torch.Tensor) -> torch.Tensor:
    """voxel_size / coors_range as contiguous float32 HOST tensors (they are, in the reference's call: voxelize.py:78-85);
    anything else is converted."""
    if t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous() and not t.requires_grad:
        return t
    return t.detach().to("cpu", torch.float32).contiguous()


def hard_voxelize_forward(*args, **kwargs):
    """Present only so the reference loader's hasattr check passes (voxelize.py:27-29);
    hard voxelization is not on the hot path."""
    raise NotImplementedError("hard_voxelize_forward is out of scope for himo_b200")


def dynamic_point_to_voxel_forward(feats: torch.Tensor, coors: torch.Tensor, reduce_type: str):
    """mmcv.dynamic_point_to_voxel_forward (scatter_points.cpp:36-41) ->
    [voxel_feats [M,C], voxel_coors [M,3], point2voxel_map [N] i32, voxel_points_count [M] i32]."""
    if reduce_type not in _REDUCE:
        # scatter_points.cpp:24-34 raises on anything else
        raise RuntimeError("do not support reduce type " + str(reduce_type))
    _lib.require_cuda(feats, "feats")
    _lib.require_cuda(coors, "coors")
    n, c = feats.shape
    if n == 0:
        # scatter_points_cuda.cu:15-18
        return [feats.clone().detach(), coors.clone().detach(),
                coors.new_empty((0,), dtype=torch.int32), coors.new_empty((0,), dtype=torch.int32)]
    if feats.dtype != torch.float32:
        raise RuntimeError("feats must be float32")
    if coors.dtype not in (torch.int32, torch.int64) or coors.shape != (n, 3):
        raise RuntimeError("coors must be int32/int64 [N,3]")
    feats_c = feats.contiguous()
    coors_c = coors.contiguous()
    dev = feats.device
    # Grid extent: the reference learns it implicitly from its sort; one tiny reduction here.
    # (The reference syncs at the same place: out_coors[0][0].lt(0).item(), scatter_points_cuda.cu:29.)
    dims_t = (coors_c.max(dim=0).values.clamp(min=0) + 1).to("cpu", torch.int32).contiguous()
    L = _lib.lib()
    ws_bytes = L.himo_dynamic_point_to_voxel_workspace_bytes(n, c, _lib.ptr(dims_t))
    if ws_bytes == 0:
        raise RuntimeError("dynamic_point_to_voxel_forward: voxel grid too large")
    with _lib.on_device(dev):
        ws = _lib.workspace.get(ws_bytes, dev)
        voxel_feats = torch.empty((n, c), dtype=torch.float32, device=dev)
        voxel_coors = torch.empty((n, 3), dtype=coors.dtype, device=dev)
        p2v = torch.empty((n,), dtype=torch.int32, device=dev)
        cnt = torch.empty((n,), dtype=torch.int32, device=dev)
        m_dev = torch.empty((1,), dtype=torch.int32, device=dev)
        st = L.himo_dynamic_point_to_voxel_forward(
            _lib.ptr(feats_c), _lib.ptr(coors_c), int(coors.dtype == torch.int64), n, c,
            _REDUCE[reduce_type], _lib.ptr(dims_t), _lib.ptr(voxel_feats), _lib.ptr(voxel_coors),
            _lib.ptr(p2v), _lib.ptr(cnt), _lib.ptr(m_dev), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
            _lib.stream_ptr(dev))
    _lib.check(st, "dynamic_point_to_voxel_forward")
    m = int(m_dev.item())  # data-dependent output size, as in the reference (unique_dim)
    return [voxel_feats[:m], voxel_coors[:m], p2v, cnt[:m]]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats,
                                    coors_idx, reduce_count, reduce_type):
    """mmcv.dynamic_point_to_voxel_backward (pybind.cpp:36-40) -- training only, outside the
    inference hot path; present for the loader's hasattr check (scatter_points.py:27-29)."""
    raise NotImplementedError("dynamic_point_to_voxel_backward is out of scope for himo_b200")
