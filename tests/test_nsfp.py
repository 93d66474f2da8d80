"""CPU: the NSFP optimisation loop (himo_b200/nsfp.py) against the reference's OWN `src.models.NSFP.optimize`
(OSF/src/models/nsfp.py:74-131) run live through oracle/ref_shims.py, both on the brute-force Chamfer;
the CUDA path is covered by tests/test_gpu_nsfp.py against the golden that class produced."""
import numpy as np
import pytest
import torch

from himo_b200 import frames, nsfp
from oracle import ref_shims
from _cpu_chamfer import CpuChamferDis


def nsfp_pair(n=1500, seed=51, half_extent=10.0):
    tr = frames.lidar_triple(n * 4, seed)
    crop = lambda a: np.ascontiguousarray(a[(np.abs(a[:, 0]) < half_extent) & (np.abs(a[:, 1]) < half_extent)][:n, :3])
    return torch.from_numpy(crop(tr["pc0"])), torch.from_numpy(crop(tr["pc1"]))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("itr_num,patience", [(4, 30), (12, 2)])
def test_nsfp_optimize_matches_live_reference(itr_num, patience):
    models = ref_shims.import_models()
    pc0, pc1 = nsfp_pair()
    torch.set_num_threads(1)
    ref = models.NSFP(itr_num=itr_num, early_patience=patience)
    torch.manual_seed(7)
    r = ref.optimize({"pc0": pc0.clone().requires_grad_(True), "pc1": pc1.clone().requires_grad_(True)})
    mine = nsfp.NSFP(itr_num=itr_num, early_patience=patience, chamfer=CpuChamferDis())
    torch.manual_seed(7)
    m = mine.optimize(pc0, pc1)
    assert m["loss"] == pytest.approx(r["loss"], rel=1e-5)
    np.testing.assert_allclose(m["flow"].numpy(), r["flow"].detach().numpy(), rtol=0, atol=2e-5)


def test_prior_state_dict_layout_and_early_stop():
    p = nsfp._Prior()
    keys = list(p.state_dict().keys())
    assert keys[:2] == ["nn_layers.0.0.weight", "nn_layers.0.0.bias"] and keys[-2:] == ["nn_layers.16.weight", "nn_layers.16.bias"]
    assert sum(v.numel() for v in p.state_dict().values()) == 3 * 128 + 128 + 7 * (128 * 128 + 128) + 128 * 3 + 3
    model = nsfp.NSFP(chamfer=CpuChamferDis())
    model.timer[12].start("One Scan"); model.timer[12].stop(); model.timer.print()      # dztimer-like, as the runner uses it
    es = nsfp._EarlyStop(patience=2, min_delta=0.1)
    assert [es.step(v) for v in (1.0, 0.95, 0.85, 0.84, 0.83)] == [False, False, False, False, True]
    assert nsfp._EarlyStop(patience=3, min_delta=0.0).step(float("nan")) is False       # first value only seeds `best`
    es = nsfp._EarlyStop(patience=3, min_delta=0.0); es.step(1.0)
    assert es.step(float("nan")) is True
    assert not any(nsfp._EarlyStop(patience=0, min_delta=0.0).step(v) for v in (1.0, 2.0, 3.0))


def golden_state_dict(z):
    return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}


@pytest.mark.parametrize("k", [3, 12])
def test_nsfp_optimize_matches_reference_golden(k):
    import glob, os
    from conftest import GOLDEN
    from himo_b200.deflowpp import cal_pose0to1
    (path,) = glob.glob(os.path.join(GOLDEN, f"nsfp_*_k{k}.npz"))
    z = np.load(path)
    torch.set_num_threads(1)
    model = nsfp.NSFP(itr_num=int(z["itr_num"]), early_patience=int(z["patience"]), chamfer=CpuChamferDis())
    pc0, pc1 = torch.from_numpy(z["pc0"]), torch.from_numpy(z["pc1"])
    sel0, rm0 = model.range_limit_(pc0)
    sel1, _ = model.range_limit_(pc1)
    T = cal_pose0to1(torch.from_numpy(z["pose0"]), torch.from_numpy(z["pose1"]))
    tr0 = sel0 @ T[:3, :3].T + T[:3, 3]
    np.testing.assert_allclose((tr0 - sel0).numpy(), z["pose_flow"], rtol=0, atol=1e-6)
    res = model.optimize(tr0, sel1, init_state_dict=golden_state_dict(z))
    assert res["iterations"] == k
    np.testing.assert_allclose(res["flow"].numpy(), z["flow"][rm0.numpy()], rtol=0, atol=2e-6 if k == 3 else 5e-5)
