"""GPU: himo_b200.nsfp.NSFP (≙ src.models.NSFP, OSF/src/models/nsfp.py) on the CUDA Chamfer path against the output of
the reference's own class (tests/golden/nsfp_*.npz, made by tests/golden/make_golden.py::nsfp)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import nsfp
from test_nsfp import golden_state_dict

pytestmark = pytest.mark.gpu


def _batch(z):
    c = lambda k: torch.from_numpy(z[k]).cuda()
    return {"pc0": [c("pc0")], "pc1": [c("pc1")], "pose0": [torch.from_numpy(z["pose0"])], "pose1": [torch.from_numpy(z["pose1"])]}


# fp32 tolerance of the north star is 1e-4 on the flow; the optimiser amplifies rounding differences between the CPU
# and GPU matrix products by about 10x every 3 iterations, so only the short horizon is held to it
@pytest.mark.parametrize("k,atol", [(3, 1e-4), (12, None)])
def test_nsfp_forward_matches_reference_golden(k, atol):
    (path,) = glob.glob(os.path.join(GOLDEN, f"nsfp_*_k{k}.npz"))
    z = np.load(path)
    model = nsfp.NSFP(itr_num=int(z["itr_num"]), early_patience=int(z["patience"]))
    out = model(_batch(z), init_state_dicts=[golden_state_dict(z)])
    assert model.last_info["iterations"] == k
    np.testing.assert_allclose(out["pose_flow"][0].cpu().numpy(), z["pose_flow"], rtol=0, atol=2e-6)
    flow = out["flow"][0].cpu().numpy()
    assert flow.shape == z["flow"].shape and np.isfinite(flow).all()
    if atol is not None:
        np.testing.assert_allclose(flow, z["flow"], rtol=0, atol=atol)
    else:       # a 1e-7 relative perturbation of the weights already moves single points by 3e-3 after 12 iterations
        assert np.abs(flow - z["flow"]).mean() < 3e-3 and np.abs(z["flow"]).mean() > 1e-2


def test_nsfp_loss_decreases_and_early_stop():
    """On this pair the first Adam steps overshoot: with patience 5 the run ends after 1 + 5 iterations holding the
    initial flow; with the default patience the loss gets below where it started (0.1527 -> 0.1465 after 45)."""
    (path,) = glob.glob(os.path.join(GOLDEN, "nsfp_*_k3.npz"))
    z = np.load(path)
    short = nsfp.NSFP(itr_num=1)
    short(_batch(z), init_state_dicts=[golden_state_dict(z)])
    assert short.last_info["loss"] == pytest.approx(0.152696, rel=1e-4)
    impatient = nsfp.NSFP(itr_num=60, early_patience=5)
    impatient(_batch(z), init_state_dicts=[golden_state_dict(z)])
    assert impatient.last_info["iterations"] == 6 and impatient.last_info["loss"] == short.last_info["loss"]
    longer = nsfp.NSFP(itr_num=45, early_patience=30)
    longer(_batch(z), init_state_dicts=[golden_state_dict(z)])
    assert longer.last_info["iterations"] > 6 and longer.last_info["loss"] < 0.1515
