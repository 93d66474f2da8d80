"""GPU: the autograd face of the 1-NN kernels (himo_b200/chamfer3d.py ≙ OSF/assets/cuda/chamfer3D/__init__.py) and the
SeFlow / SeFlow++ losses on top of it, against the brute-force oracle and the reference's own golden values."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import chamfer3d, frames, lossfuncs
from oracle import leaf
from test_lossfuncs import TERMS, run_loss, synth_loss_frame

pytestmark = pytest.mark.gpu


def _pair(n, seed):
    tr = frames.lidar_triple(n, seed)
    return tr["pc0"][:, :3].copy(), tr["pc1"][:, :3].copy()


def _oracle_truncated(pc0, pc1, mode, t):
    """value and d/dpc0, d/dpc1 of the three reductions of chamfer3D/__init__.py from brute-force distances."""
    d0, d1, i0, i1 = leaf.chamfer_forward(pc0, pc1)
    if mode == "plain":
        k0, k1 = np.ones_like(d0, bool), np.ones_like(d1, bool)
        n0, n1 = d0.size, d1.size
    elif mode == "forward":                       # mean over the kept entries
        k0, k1 = d0 <= t, d1 <= t
        n0, n1 = k0.sum(), k1.sum()
    else:                                         # truncated_dis: dropped entries stay in the mean as zeros
        k0, k1 = d0 < t, d1 < t
        n0, n1 = d0.size, d1.size
    val = d0[k0].astype(np.float64).sum() / n0 + d1[k1].astype(np.float64).sum() / n1
    g0 = (k0 / n0).astype(np.float32); g1 = (k1 / n1).astype(np.float32)
    ga, gb = leaf.chamfer_backward(pc0, pc1, i0, i1, g0, g1)
    return val, np.asarray(ga), np.asarray(gb)


@pytest.mark.parametrize("mode,t", [("plain", -1), ("forward", 4), ("forward", 0.25), ("nsfp", 2), ("nsfp", 0.09)])
def test_nnchamferdis_values_and_gradients(mode, t):
    pc0, pc1 = _pair(6000, 41)
    a = torch.from_numpy(pc0).cuda().requires_grad_(True)
    b = torch.from_numpy(pc1).cuda().requires_grad_(True)
    m = chamfer3d.nnChamferDis()
    loss = m.truncated_dis(a, b, truncate_dist=t) if mode == "nsfp" else m(a, b, truncate_dist=t)
    loss.backward()
    val, ga, gb = _oracle_truncated(pc0, pc1, mode, t)
    assert float(loss) == pytest.approx(val, rel=2e-6)
    np.testing.assert_allclose(a.grad.cpu().numpy(), ga, rtol=0, atol=2e-7)
    np.testing.assert_allclose(b.grad.cpu().numpy(), gb, rtol=0, atol=2e-7)     # atomics: order-dependent last bit


def test_dis_res_and_disid_res_are_bit_exact():
    pc0, pc1 = _pair(5000, 42)
    a, b = torch.from_numpy(pc0).cuda(), torch.from_numpy(pc1).cuda()
    m = chamfer3d.nnChamferDis()
    d0, d1, i0, i1 = m.disid_res(a, b)
    r = leaf.chamfer_forward(pc0, pc1)
    for got, ref in zip((d0, d1, i0, i1), r):
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
    e0, e1 = m.dis_res(a, b)
    assert torch.equal(e0, d0) and torch.equal(e1, d1)
    nn = chamfer3d.NearestNeighborDis()(a, b)
    assert float(nn) == pytest.approx(float(r[0][r[0] <= 2].astype(np.float64).mean()), rel=2e-6)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "seflow_loss_*.npz"))))
def test_seflow_losses_match_reference_golden(path):
    z = np.load(path)
    frame = {k: v.cuda() for k, v in synth_loss_frame(int(z["seed"]), int(z["n"]), dynamic_fraction=float(z["frac"])).items()}
    for name in ("seflowLoss", "seflowppLoss"):
        d = {k: v.clone() for k, v in frame.items()}
        d["est_flow"].requires_grad_(True)
        out = getattr(lossfuncs, name)(d)
        grad = torch.autograd.grad(sum(out[k] for k in TERMS), d["est_flow"])[0].cpu().numpy()
        for i, k in enumerate(TERMS):
            assert float(out[k]) == pytest.approx(float(z[name + "_terms"][i]), rel=5e-6, abs=1e-7), (name, k)
        np.testing.assert_allclose(grad, z[name + "_grad"], rtol=0, atol=2e-7)
