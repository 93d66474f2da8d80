"""oracle/leaf.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy-facing wrappers of oracle/leaf_ops.c plus small pure-numpy twins that the tests
use to cross-check the C restatement itself.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from . import lib

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


REDUCE = {"sum": 0, "mean": 1, "max": 2}


# --------------------------------------------------------------------------- voxelize
def dynamic_voxelize(points: np.ndarray, voxel_size, coors_range) -> np.ndarray:
    """mmcv.dynamic_voxelize_forward (OSF/assets/cuda/mmcv/voxelization_cuda_kernel.cuh:13-50).
    points [N,F>=3] f32 -> coors [N,3] int32 in (z,y,x) order with the partial -1 marks."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, f = pts.shape
    coors = np.zeros((n, 3), dtype=np.int32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    if n:
        lib().oracle_dynamic_voxelize(_p(pts, _f32p), n, f, _p(vs, _f32p), _p(cr, _f32p),
                                      _p(coors, _i32p))
    return coors


def dynamic_voxelize_np(points: np.ndarray, voxel_size, coors_range) -> np.ndarray:
    """Pure-numpy twin of dynamic_voxelize (same reference lines)."""
    pts = np.asarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    grid = np.round((cr[3:] - cr[:3]) / vs).astype(np.int64)
    with np.errstate(invalid="ignore", over="ignore"):
        c = np.floor((pts[:, :3] - cr[:3]) / vs)
    c = np.nan_to_num(c, nan=0.0, posinf=2.0**31 - 1, neginf=-2.0**31)
    c = np.clip(c, -2.0**31, 2.0**31 - 1).astype(np.int64)
    ok = (c >= 0) & (c < grid)
    okx, oky, okz = ok[:, 0], ok[:, 1], ok[:, 2]
    out = np.zeros((pts.shape[0], 3), dtype=np.int32)
    bad_x = ~okx
    bad_y = okx & ~oky
    bad_z = okx & oky & ~okz
    good = okx & oky & okz
    out[bad_x, 0] = -1
    out[bad_y, 0] = -1
    out[bad_y, 1] = -1
    out[bad_z] = -1
    out[good, 0] = c[good, 2]
    out[good, 1] = c[good, 1]
    out[good, 2] = c[good, 0]
    return out


# --------------------------------------------------------------------------- scatter
def dynamic_point_to_voxel(feats: np.ndarray, coors: np.ndarray, reduce_type: str = "mean",
                           accum: str = "f32_seq"
                           ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """mmcv.dynamic_point_to_voxel_forward (OSF/assets/cuda/mmcv/scatter_points_cuda.cu:9-66).
    Returns (voxel_feats [M,C] f32, voxel_coors [M,3] (dtype of coors), point2voxel [N] i32,
    voxel_points_count [M] i32).  accum: 'f32_seq' = fp32 sums in ascending point order,
    'exact' = double sums rounded once (order-free limit of the device's unordered atomics)."""
    f = np.ascontiguousarray(feats, dtype=np.float32)
    n, c = f.shape
    co = np.ascontiguousarray(coors).astype(np.int64)
    if n == 0:
        return (f.copy(), np.asarray(coors).copy(), np.zeros(0, np.int32), np.zeros(0, np.int32))
    vf = np.empty((n, c), np.float32)
    vc = np.empty((n, 3), np.int64)
    p2v = np.empty(n, np.int32)
    cnt = np.empty(n, np.int32)
    m = lib().oracle_dynamic_point_to_voxel(_p(f, _f32p), _p(co, _i64p), n, c,
                                            REDUCE[reduce_type], 1 if accum == "exact" else 0,
                                            _p(vf, _f32p), _p(vc, _i64p), _p(p2v, _i32p),
                                            _p(cnt, _i32p))
    return vf[:m].copy(), vc[:m].astype(np.asarray(coors).dtype), p2v, cnt[:m].copy()


def dynamic_point_to_voxel_np(feats, coors, reduce_type="mean"):
    """Pure-numpy twin (np.unique(axis=0) is the same sorted-unique as at::unique_dim)."""
    f = np.asarray(feats, dtype=np.float32)
    co = np.asarray(coors).astype(np.int64).copy()
    if f.shape[0] == 0:
        return f.copy(), np.asarray(coors).copy(), np.zeros(0, np.int32), np.zeros(0, np.int32)
    co[(co < 0).any(axis=1)] = -1
    uniq, inv, cnt = np.unique(co, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    if uniq[0, 0] < 0:
        uniq, cnt, inv = uniq[1:], cnt[1:], inv - 1
    m = uniq.shape[0]
    valid = inv >= 0
    if reduce_type == "max":
        out = np.full((m, f.shape[1]), -np.inf, np.float32)
        np.maximum.at(out, inv[valid], f[valid])
    else:
        out = np.zeros((m, f.shape[1]), np.float32)
        np.add.at(out, inv[valid], f[valid])
        if reduce_type == "mean":
            out = out / cnt.astype(np.float32)[:, None]
    return out, uniq.astype(np.asarray(coors).dtype), inv.astype(np.int32), cnt.astype(np.int32)


# --------------------------------------------------------------------------- chamfer
def nn_bruteforce(query: np.ndarray, ref: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """One direction of chamfer3D.forward (OSF/assets/cuda/chamfer3D/chamfer3D.cu:33-83):
    squared distance to, and index of, the nearest ref point for every query point."""
    q = np.ascontiguousarray(query[:, :3], dtype=np.float32)
    r = np.ascontiguousarray(ref[:, :3], dtype=np.float32)
    dist = np.empty(q.shape[0], np.float32)
    idx = np.empty(q.shape[0], np.int32)
    if q.shape[0]:
        lib().oracle_nn_bruteforce(_p(q, _f32p), q.shape[0], _p(r, _f32p), r.shape[0],
                                   _p(dist, _f32p), _p(idx, _i32p))
    return dist, idx


def chamfer_forward(pc0: np.ndarray, pc1: np.ndarray):
    """chamfer3D.forward (chamfer3D.cu:85-105): (dist0, dist1, idx0, idx1)."""
    d0, i0 = nn_bruteforce(pc0, pc1)
    d1, i1 = nn_bruteforce(pc1, pc0)
    return d0, d1, i0, i1


def nn_bruteforce_np(query, ref):
    """Pure-numpy twin for small inputs; emulates the device's FMA chain in float64
    (exact for the products, rounded to fp32 after each fused step)."""
    q = np.asarray(query[:, :3], np.float32)
    r = np.asarray(ref[:, :3], np.float32)
    if r.shape[0] == 0:
        return np.full(q.shape[0], 1e20, np.float32), np.full(q.shape[0], -1, np.int32)
    d = (r[None, :, :] - q[:, None, :]).astype(np.float32).astype(np.float64)
    t = (d[..., 1] * d[..., 1]).astype(np.float32).astype(np.float64)          # dy*dy rounded, then fma dx, fma dz
    t = (d[..., 0] * d[..., 0] + t).astype(np.float32).astype(np.float64)
    t = (d[..., 2] * d[..., 2] + t).astype(np.float32)
    idx = np.argmin(t, axis=1).astype(np.int32)  # first minimiser = lowest index
    return t[np.arange(q.shape[0]), idx], idx


def chamfer_backward(pc0, pc1, idx0, idx1, g0, g1):
    """chamfer3D.backward (chamfer3D.cu:107-154) -> (grad_pc0, grad_pc1)."""
    a = np.ascontiguousarray(pc0, np.float32)
    b = np.ascontiguousarray(pc1, np.float32)
    i0 = np.ascontiguousarray(idx0, np.int32)
    i1 = np.ascontiguousarray(idx1, np.int32)
    g0 = np.ascontiguousarray(g0, np.float32)
    g1 = np.ascontiguousarray(g1, np.float32)
    ga = np.zeros_like(a)
    gb = np.zeros_like(b)
    lib().oracle_chamfer_backward(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0],
                                  _p(i0, _i32p), _p(i1, _i32p), _p(g0, _f32p), _p(g1, _f32p),
                                  _p(ga, _f32p), _p(gb, _f32p))
    return ga, gb


# --------------------------------------------------------------------------- FastGeodis DT
def geodesic3d_euclid(mask_dist: np.ndarray, spacing, iterations: int = 1) -> np.ndarray:
    """FastGeodis.generalised_geodesic3d with lamb=0 on `mask_dist` = v*softmask
    ([H,W,D] f32).  PARITY UNPINNED third-party restatement, see leaf_ops.c."""
    d = np.ascontiguousarray(mask_dist, dtype=np.float32).copy()
    sp = np.asarray(spacing, np.float32)
    lib().oracle_geodesic3d_euclid(_p(d, _f32p), d.shape[0], d.shape[1], d.shape[2],
                                   _p(sp, _f32p), int(iterations))
    return d
