"""CPU: himo_b200/lossfuncs.py (segmented SeFlow / SeFlow++ losses) against the reference's OWN
OSF/src/lossfuncs/selfsupervise.py (live, through oracle/ref_shims.py) and against the golden values that script
produced (tests/golden/seflow_loss_*.npz), values and gradients w.r.t. the estimated flow."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import lossfuncs
from oracle import ref_shims
from _cpu_chamfer import CpuChamferDis

TERMS = ("chamfer_dis", "dynamic_chamfer_dis", "static_flow_loss", "cluster_based_pc0pc1")


def synth_loss_frame(seed: int, n: int = 2400, n_clusters: int = 9, dynamic_fraction: float = 0.35):
    """Three sweeps with clustered movers: label 0 static, 1 dynamic-but-unclustered, >= 2 cluster ids."""
    rng = np.random.default_rng(seed)
    lab0 = np.zeros(n, np.int64)
    n_dyn = int(n * dynamic_fraction)
    pc0 = rng.uniform([-30, -30, -1], [30, 30, 2], (n, 3))
    vel = np.zeros((n, 3))
    if n_dyn:
        centres = rng.uniform(-25, 25, (n_clusters, 2))
        who = rng.integers(0, n_clusters, n_dyn)
        pc0[:n_dyn, :2] = centres[who] + rng.normal(0, 0.8, (n_dyn, 2))
        lab0[:n_dyn] = who + 2
        lab0[:n_dyn][rng.random(n_dyn) < 0.1] = 1
        cv = rng.normal(0, 0.7, (n_clusters, 3)) * [1, 1, 0]
        vel[:n_dyn] = cv[who]
    jitter = lambda: rng.normal(0, 0.02, (n, 3))
    keep1, keeph = rng.random(n) < 0.93, rng.random(n) < 0.93
    pc1, lab1 = (pc0 + vel + jitter())[keep1], lab0[keep1].copy()
    pch1, labh = (pc0 - vel + jitter())[keeph], lab0[keeph].copy()
    lab1[rng.random(lab1.shape[0]) < 0.15] = 0            # label noise: some movers look static in the next sweep
    perm = rng.permutation(n)
    est = (vel + rng.normal(0, 0.15, (n, 3)))[perm]
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return {"pc0": f(pc0[perm]), "pc1": f(pc1), "pch1": f(pch1), "est_flow": f(est),
            "pc0_labels": torch.from_numpy(lab0[perm]), "pc1_labels": torch.from_numpy(lab1), "pch1_labels": torch.from_numpy(labh)}


CASES = [(31, 2400, 0.35), (32, 2400, 0.05), (33, 900, 0.5), (34, 2400, 0.0)]     # clustered / too few movers / small / none


def run_loss(fn, frame, **kw):
    d = {k: v.clone() for k, v in frame.items()}
    d["est_flow"].requires_grad_(True)
    out = fn(d, **kw)
    total = sum(out[k] for k in TERMS)
    grad = torch.autograd.grad(total, d["est_flow"], allow_unused=True)[0]
    return {k: float(out[k].detach()) for k in TERMS}, (torch.zeros_like(d["est_flow"]) if grad is None else grad).numpy()


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("seed,n,frac", CASES)
@pytest.mark.parametrize("name", ["seflowLoss", "seflowppLoss"])
def test_losses_match_live_reference(name, seed, n, frac):
    ref = ref_shims.import_lossfuncs()
    frame = synth_loss_frame(seed, n, dynamic_fraction=frac)
    rv, rg = run_loss(getattr(ref, name), frame)
    v, g = run_loss(getattr(lossfuncs, name), frame, chamfer=CpuChamferDis())
    for k in TERMS:
        assert v[k] == pytest.approx(rv[k], rel=2e-6, abs=1e-7), k
    np.testing.assert_allclose(g, rg, rtol=0, atol=1e-7)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("name", ["seflowLoss", "seflowppLoss"])
def test_losses_fallback_when_no_cluster_qualifies(name):
    """Enough movers but none of them clustered (all label 1): selfsupervise.py:99-100 falls back to the truncated raw
    Chamfer distance, which carries no gradient."""
    ref = ref_shims.import_lossfuncs()
    frame = synth_loss_frame(35, 2400, dynamic_fraction=0.4)
    frame["pc0_labels"][frame["pc0_labels"] > 1] = 1
    rv, rg = run_loss(getattr(ref, name), frame)
    v, g = run_loss(getattr(lossfuncs, name), frame, chamfer=CpuChamferDis())
    assert rv["cluster_based_pc0pc1"] > 0 and rv["dynamic_chamfer_dis"] > 0
    for k in TERMS:
        assert v[k] == pytest.approx(rv[k], rel=2e-6, abs=1e-7), k
    np.testing.assert_allclose(g, rg, rtol=0, atol=1e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "seflow_loss_*.npz"))))
def test_losses_match_reference_golden(path):
    z = np.load(path)
    frame = synth_loss_frame(int(z["seed"]), int(z["n"]), dynamic_fraction=float(z["frac"]))
    for name in ("seflowLoss", "seflowppLoss"):
        v, g = run_loss(getattr(lossfuncs, name), frame, chamfer=CpuChamferDis())
        for i, k in enumerate(TERMS):
            assert v[k] == pytest.approx(float(z[name + "_terms"][i]), rel=2e-6, abs=1e-7), (name, k)
        np.testing.assert_allclose(g, z[name + "_grad"], rtol=0, atol=1e-7)


def test_cluster_term_picks_farthest_member_with_dynamic_neighbour():
    """Hand case for Eq. 8-11: cluster 2 has three members; the farthest one's neighbour is labelled static, so the
    second farthest supplies the displacement; cluster 3 has no member with a dynamic neighbour and is skipped."""
    pc0 = torch.tensor([[0., 0, 0], [10, 0, 0], [20, 0, 0], [30, 0, 0], [40, 0, 0]])
    pc1 = torch.tensor([[0.5, 0, 0], [11.5, 0, 0], [23, 0, 0], [30.2, 0, 0], [40.1, 0, 0]])
    lab0 = torch.tensor([2, 2, 2, 3, 0]); lab1 = torch.tensor([1, 1, 0, 0, 0])
    est = torch.tensor([[1., 0, 0], [1, 0, 0], [1, 0, 0], [5, 0, 0], [0, 3, 4]], requires_grad=True)
    ch = CpuChamferDis()
    d0, d1, i0, _ = ch.disid_res(pc0, pc1)
    s, m = lossfuncs._static_and_cluster_terms(pc0, pc1, est, lab0, lab1, d0, d1, i0, True)
    assert float(s) == pytest.approx(5.0)
    assert float(m) == pytest.approx(0.5)        # |1 - 1.5| for the three members of cluster 2 only


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_losses_match_live_reference_on_random_label_structures():
    """Twenty seeded frames with irregular label sets (gaps in the cluster ids, clusters of one point, no static points,
    no dynamic neighbours at all, duplicated points giving exact distance ties inside a cluster)."""
    ref = ref_shims.import_lossfuncs()
    rng = np.random.default_rng(123)
    for trial in range(20):
        fr = synth_loss_frame(100 + trial, n=int(rng.integers(700, 1400)), n_clusters=int(rng.integers(1, 30)),
                              dynamic_fraction=float(rng.uniform(0.3, 0.9)))
        l0 = fr["pc0_labels"]
        if trial % 4 == 0:
            l0[l0 > 1] = l0[l0 > 1] * 7 + 3                  # gaps in the id space
        if trial % 5 == 1:
            l0[l0 == 0] = 1                                   # no static points
        if trial % 5 == 2:
            fr["pc1_labels"][:] = 0                           # nothing dynamic in pc1: every cluster is skipped
        if trial % 5 == 3:
            dyn = torch.nonzero(l0 > 1).squeeze(1)
            l0[dyn[:5]] = torch.arange(1000, 1005)            # five clusters of a single point
        if trial % 5 == 4:
            for c in torch.unique(l0[l0 > 1])[:6]:            # duplicated points: exact distance ties inside a cluster
                members = torch.nonzero(l0 == c).squeeze(1)
                if members.numel() >= 2:
                    fr["pc0"][members[1]] = fr["pc0"][members[0]]
        for name in ("seflowLoss", "seflowppLoss"):
            rv, rg = run_loss(getattr(ref, name), fr)
            v, g = run_loss(getattr(lossfuncs, name), fr, chamfer=CpuChamferDis())
            for k in TERMS:
                assert v[k] == pytest.approx(rv[k], rel=2e-6, abs=1e-7), (trial, name, k)
            np.testing.assert_allclose(g, rg, rtol=0, atol=1e-7, err_msg=f"{trial} {name}")
