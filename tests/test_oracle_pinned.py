"""CPU: pin the network-level oracle restatement to outputs of the reference's OWN classes
(tests/golden/*.npz, made by tests/golden/make_golden.py in the build container) and, when
/root/reference is present, to the live reference run."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import weights
from oracle import deflowpp_ref, fastnsf_ref, ref_shims


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "deflowpp_*.npz"))))
def test_deflowpp_oracle_matches_reference_golden(path):
    z = np.load(path)
    sd = weights.synth_deflowpp_state_dict(int(z["weight_seed"]))
    out = deflowpp_ref.deflowpp_forward(sd, z["pch1"], z["pc0"], z["pc1"], z["poseh1"], z["pose0"], z["pose1"])
    assert (out["pc0_valid_point_idxes"].numpy() == z["pc0_valid_point_idxes"]).all()
    assert (out["pc1_valid_point_idxes"].numpy() == z["pc1_valid_point_idxes"]).all()
    assert (out["pch1_valid_point_idxes"].numpy() == z["pch1_valid_point_idxes"]).all()
    np.testing.assert_array_equal(out["pose_flow"].numpy(), z["pose_flow"])
    # same torch CPU kernels on both sides; only the thread-dependent summation order differs
    np.testing.assert_allclose(out["flow"].numpy(), z["flow"], rtol=0, atol=5e-5)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_deflowpp_oracle_matches_live_reference():
    from himo_b200 import frames
    models = ref_shims.import_models()
    sd = weights.synth_deflowpp_state_dict(3)
    net = models.DeFlowPP().eval()
    net.load_state_dict(sd, strict=True)
    tr = frames.uniform_triple(1500, 21)
    batch = {k: torch.from_numpy(tr[k])[None] for k in ("pc0", "pc1", "pch1")}
    batch.update({k: [torch.from_numpy(tr[k])] for k in ("pose0", "pose1", "poseh1")})
    with torch.no_grad():
        ref = net(batch)
    out = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"], tr["pose1"])
    assert (ref["pc0_valid_point_idxes"][0] == out["pc0_valid_point_idxes"]).all()
    np.testing.assert_allclose(out["flow"].numpy(), ref["flow"][0].numpy(), rtol=0, atol=5e-5)


def _golden_prior_state_dict(z):
    return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fastnsf_*.npz"))))
def test_fastnsf_oracle_matches_reference_golden(path):
    """oracle/fastnsf_ref.py against the output of the reference's own `src.models.FastNSF` class
    (OSF/src/models/fastnsf.py:83-222) started from the very weights that class drew; both sides use the restated
    FastGeodis transform, so this pins everything of H3 except that transform."""
    z = np.load(path)
    torch.set_num_threads(1)
    out = fastnsf_ref.fastnsf_forward(_golden_prior_state_dict(z), z["pc0"], z["pc1"], z["pose0"], z["pose1"],
                                      itr_num=int(z["itr_num"]), patience=int(z["patience"]))
    np.testing.assert_array_equal(out["pose_flow"].numpy(), z["pose_flow"])
    assert out["iterations"] == int(z["itr_num"])
    np.testing.assert_allclose(out["final_flow"].numpy(), z["flow"], rtol=0, atol=2e-6)
