"""GPU: our operators against the reference's OWN CUDA kernels, compiled for sm_100a from /root/reference into
oracle/_ref/{chamfer3D,mmcv}.so by oracle/build_ref.py (the build container does that; the .so files travel with the
snapshot).  Same inputs, same pybind signatures on both sides.  This is the pin the CPU restatement cannot give:
voxel indices, scatter maps, nearest-neighbour indices and squared distances are required to be bit-equal to what the
reference computes on the same GPU.  Skipped when the reference build is not present.

Measured on a B200 (profiles/r01_ref_kernels_vs_ours.json, profiles/r01_ref_kernels_tests.log): voxel coordinates,
scatter maps / counts / voxel order and ALL nearest-neighbour indices (376 k queries) bit-equal; voxel means differ by
2.4e-7 (atomics order); the squared distances differed in the last bit for a share of the points because the kernel and
oracle/leaf_ops.c then rounded fma(dz,dz, fma(dy,dy, dx*dx)) where the reference binary (SASS of
oracle/_ref/chamfer3D.so) computes fma(dz,dz, fma(dx,dx, dy*dy)).  Both were changed to the binary's sequence; the
re-run at the start of round 2 (profiles/r02_ref_kernels_strict_tests.log, profiles/r02_ref_kernels_vs_ours.json) then
found every index AND every squared distance bit-equal to the reference binary (0 mismatches in 376 k queries), so the
Chamfer test demands strict equality and runs in the normal -m gpu suite."""
import numpy as np
import pytest
import torch

from himo_b200 import chamfer3d_ext, frames, mmcv_ext
from oracle import build_ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (build_ref.built("chamfer3D") and build_ref.built("mmcv")),
                                 reason="oracle/_ref/*.so not built (python -m oracle.build_ref)")]
VS, RNG = frames.VOXEL_SIZE, frames.POINT_CLOUD_RANGE


@pytest.fixture(scope="module")
def ref_chamfer():
    return build_ref.load("chamfer3D")


@pytest.fixture(scope="module")
def ref_mmcv():
    return build_ref.load("mmcv")


def _clouds(kind, n, seed, fixture_clouds):
    if kind == "fixture":
        return fixture_clouds[0], fixture_clouds[1]
    if kind == "uniform":
        return frames.uniform_frame(n, seed)[:, :3].copy(), frames.uniform_frame(n - 7, seed + 1)[:, :3].copy()
    tr = frames.lidar_triple(n, seed)
    return tr["pc0"][:, :3].copy(), tr["pc1"][:, :3].copy()


def _chamfer(mod, a, b):
    d0 = torch.zeros(a.shape[0], device="cuda"); d1 = torch.zeros(b.shape[0], device="cuda")
    i0 = torch.zeros(a.shape[0], dtype=torch.int32, device="cuda"); i1 = torch.zeros(b.shape[0], dtype=torch.int32, device="cuda")
    mod.forward(a, b, d0, d1, i0, i1)
    return d0, d1, i0, i1


@pytest.mark.parametrize("kind,n,seed", [("fixture", 0, 0), ("lidar", 30000, 61), ("uniform", 20000, 62), ("lidar", 257, 63)])
def test_chamfer_forward_backward_equal_reference_kernels(ref_chamfer, kind, n, seed, fixture_clouds):
    a_np, b_np = _clouds(kind, n, seed, fixture_clouds)
    a, b = torch.from_numpy(a_np).cuda().contiguous(), torch.from_numpy(b_np).cuda().contiguous()
    ours, ref = _chamfer(chamfer3d_ext, a, b), _chamfer(ref_chamfer, a, b)
    for name, x, y in zip(("idx0", "idx1", "dist0", "dist1"), ours[2:] + ours[:2], ref[2:] + ref[:2]):
        assert torch.equal(x, y), f"{name}: {(x != y).sum().item()} of {x.numel()} differ"
    g0, g1 = torch.rand_like(ours[0]), torch.rand_like(ours[1])
    grads = []
    for mod in (chamfer3d_ext, ref_chamfer):
        ga, gb = torch.zeros_like(a), torch.zeros_like(b)
        mod.backward(a, b, ref[2], ref[3], g0, g1, ga, gb)
        grads.append((ga, gb))
    for x, y in zip(*grads):       # both sides accumulate with fp32 atomics in arbitrary order
        assert (x - y).abs().max().item() <= 1e-5 * max(1.0, y.abs().max().item())


@pytest.mark.parametrize("kind,n,seed", [("fixture", 0, 0), ("uniform", 100000, 64), ("lidar", 50000, 65)])
def test_voxelize_and_scatter_equal_reference_kernels(ref_mmcv, kind, n, seed, fixture_clouds):
    pts_np = _clouds(kind, n, seed, fixture_clouds)[0]
    pts = torch.from_numpy(np.ascontiguousarray(pts_np)).cuda()
    vs, rg = torch.tensor(VS), torch.tensor(RNG)
    co_ours = pts.new_zeros((pts.shape[0], 3), dtype=torch.int32)
    co_ref = pts.new_zeros((pts.shape[0], 3), dtype=torch.int32)
    mmcv_ext.dynamic_voxelize_forward(pts, vs, rg, co_ours, 3)
    ref_mmcv.dynamic_voxelize_forward(pts, vs, rg, co_ref, 3)
    assert torch.equal(co_ours, co_ref), f"{(co_ours != co_ref).any(1).sum().item()} rows differ"
    rng = np.random.default_rng(seed)
    for C, reduce in ((3, "mean"), (32, "mean"), (5, "max"), (4, "sum")):
        feats = torch.from_numpy(rng.normal(size=(pts.shape[0], C)).astype(np.float32)).cuda()
        o = mmcv_ext.dynamic_point_to_voxel_forward(feats, co_ref, reduce)
        r = ref_mmcv.dynamic_point_to_voxel_forward(feats, co_ref, reduce)
        assert o[1].dtype == r[1].dtype and torch.equal(o[1], r[1]), "voxel_coors"
        assert torch.equal(o[2], r[2].to(o[2].dtype)), "point2voxel_map"
        assert torch.equal(o[3], r[3].to(o[3].dtype)), "voxel_points_count"
        if reduce == "max":
            assert torch.equal(o[0], r[0])
        else:                       # the reference sums with fp32 atomics in arbitrary order
            assert (o[0] - r[0]).abs().max().item() <= 1e-5 * max(1.0, r[0].abs().max().item())
