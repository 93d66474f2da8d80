"""CPU: the C restatement of the leaf operators against its numpy twins, hand-computed cases and
the reference's known answer."""
import numpy as np
import pytest

from himo_b200 import frames
from oracle import leaf

VS, RNG = frames.VOXEL_SIZE, frames.POINT_CLOUD_RANGE


def test_voxelize_hand_cases():
    pts = np.array([
        [-51.2, -51.2, -3.0],      # exactly at min -> (0,0,0)
        [51.2, 0.0, 0.0],          # exactly at x max -> out in x
        [0.0, 51.2, 0.0],          # out in y
        [0.0, 0.0, 3.0],           # out in z (z == max)
        [0.0, 0.0, -3.1],          # out in z
        [51.19, 51.19, 2.9],       # last cell
        [-0.1, 0.1, 0.0],
        [1e9, 0.0, 0.0], [-1e9, 0.0, 0.0], [np.inf, 0, 0],
    ], np.float32)
    c = leaf.dynamic_voxelize(pts, VS, RNG)
    assert c[0].tolist() == [0, 0, 0]
    assert c[1].tolist() == [-1, 0, 0]
    assert c[2].tolist() == [-1, -1, 0]
    assert c[3].tolist() == [-1, -1, -1]
    assert c[4].tolist() == [-1, -1, -1]
    assert c[5].tolist() == [0, 511, 511]
    assert c[6].tolist() == [0, 256, 255]
    assert c[7].tolist() == [-1, 0, 0] and c[8].tolist() == [-1, 0, 0] and c[9].tolist() == [-1, 0, 0]


@pytest.mark.parametrize("seed", [0, 1])
def test_voxelize_c_vs_numpy(seed):
    pts = frames.uniform_frame(20000, seed)
    assert (leaf.dynamic_voxelize(pts, VS, RNG) == leaf.dynamic_voxelize_np(pts, VS, RNG)).all()


def test_voxelize_empty():
    assert leaf.dynamic_voxelize(np.zeros((0, 3), np.float32), VS, RNG).shape == (0, 3)


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
def test_scatter_c_vs_numpy(reduce):
    pts = frames.uniform_frame(20000, 3)
    co = leaf.dynamic_voxelize(pts, VS, RNG)
    rng = np.random.default_rng(0)
    feats = rng.normal(size=(pts.shape[0], 5)).astype(np.float32)
    a = leaf.dynamic_point_to_voxel(feats, co, reduce)
    b = leaf.dynamic_point_to_voxel_np(feats, co, reduce)
    assert a[0].shape == b[0].shape
    np.testing.assert_allclose(a[0], b[0], rtol=1e-6, atol=1e-6)
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all() and (a[3] == b[3]).all()
    # invalid rows map to -1, voxels sorted lexicographically
    bad = (co < 0).any(1)
    assert (a[2][bad] == -1).all() and (a[2][~bad] >= 0).all()
    key = (a[1][:, 0].astype(np.int64) * 512 + a[1][:, 1]) * 512 + a[1][:, 2]
    assert (np.diff(key) > 0).all()


def test_scatter_exact_close_to_sequential():
    pts = frames.lidar_triple(4000, 5)["pc0"]
    co = leaf.dynamic_voxelize(pts, VS, RNG)
    a = leaf.dynamic_point_to_voxel(pts, co, "mean", "f32_seq")
    b = leaf.dynamic_point_to_voxel(pts, co, "mean", "exact")
    np.testing.assert_allclose(a[0], b[0], rtol=0, atol=2e-5)


def test_scatter_all_invalid_and_empty():
    feats = np.ones((4, 2), np.float32)
    co = -np.ones((4, 3), np.int32)
    vf, vc, p2v, cnt = leaf.dynamic_point_to_voxel(feats, co, "mean")
    assert vf.shape == (0, 2) and vc.shape == (0, 3) and (p2v == -1).all() and cnt.shape == (0,)
    e = leaf.dynamic_point_to_voxel(np.zeros((0, 2), np.float32), np.zeros((0, 3), np.int32), "sum")
    assert e[0].shape == (0, 2) and e[2].shape == (0,)


def test_nn_c_vs_numpy_and_ties():
    rng = np.random.default_rng(1)
    q = rng.normal(size=(300, 3)).astype(np.float32)
    r = rng.normal(size=(500, 3)).astype(np.float32)
    r[100] = r[7]          # duplicate reference point: the lower index must win
    q[0] = r[7]
    d, i = leaf.nn_bruteforce(q, r)
    d2, i2 = leaf.nn_bruteforce_np(q, r)
    assert (d == d2).all() and (i == i2).all()
    assert i[0] == 7 and d[0] == 0.0


def test_nn_empty_reference():
    d, i = leaf.nn_bruteforce(np.zeros((3, 3), np.float32), np.zeros((0, 3), np.float32))
    assert (d == np.float32(1e20)).all() and (i == -1).all()


def test_chamfer_known_answer_subsampled(fixture_clouds):
    """Reference known answer: Chamfer loss 0.1710 on its own fixture clouds
    (OSF/assets/tests/chamferdis_speed_test.py:113-126).  The full 88k x 88k brute force is the
    gpu test's job; on CPU we check a strided subset against scipy's exact KD-tree."""
    from scipy.spatial import cKDTree
    pc0, pc1, _ = fixture_clouds
    a, b = pc0[::6], pc1[::6]
    d0, d1, i0, i1 = leaf.chamfer_forward(a, b)
    t0, t1 = cKDTree(b.astype(np.float64)), cKDTree(a.astype(np.float64))
    k0, _ = t0.query(a.astype(np.float64))
    k1, _ = t1.query(b.astype(np.float64))
    np.testing.assert_allclose(d0, k0 ** 2, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(d1, k1 ** 2, rtol=1e-5, atol=1e-7)


@pytest.mark.slow
def test_chamfer_known_answer_full(fixture_clouds):
    pc0, pc1, known = fixture_clouds
    from scipy.spatial import cKDTree
    k0, _ = cKDTree(pc1.astype(np.float64)).query(pc0.astype(np.float64))
    k1, _ = cKDTree(pc0.astype(np.float64)).query(pc1.astype(np.float64))
    loss = float((k0 ** 2).mean() + (k1 ** 2).mean())
    assert abs(loss - known) < 5e-4      # recorded to 4 decimals: 0.1710


def test_chamfer_backward_matches_autograd():
    import torch
    rng = np.random.default_rng(2)
    a = rng.normal(size=(200, 3)).astype(np.float32)
    b = rng.normal(size=(150, 3)).astype(np.float32)
    d0, d1, i0, i1 = leaf.chamfer_forward(a, b)
    g0 = rng.normal(size=200).astype(np.float32)
    g1 = rng.normal(size=150).astype(np.float32)
    ga, gb = leaf.chamfer_backward(a, b, i0, i1, g0, g1)
    ta = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tb = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    l = (torch.tensor(g0, dtype=torch.float64) * ((ta - tb[i0.astype(np.int64)]) ** 2).sum(1)).sum() + \
        (torch.tensor(g1, dtype=torch.float64) * ((tb - ta[i1.astype(np.int64)]) ** 2).sum(1)).sum()
    l.backward()
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-4, atol=1e-5)
