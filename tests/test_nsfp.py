"""CPU: host-side pieces of himo_b200/nsfp.py (state-dict layout of the initial network, the EarlyStopping rule, the
dztimer-like attribute).  The optimisation loop itself is CUDA only (MLP, Adam and the loop's bookkeeping run in
libhimo_b200.so): tests/test_gpu_nsfp.py holds it to the goldens produced by the reference's OWN `src.models.NSFP`
(OSF/src/models/nsfp.py:74-131, tests/golden/make_golden.py::nsfp)."""
import numpy as np
import pytest
import torch

from himo_b200 import nsfp
from _cpu_chamfer import CpuChamferDis


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
