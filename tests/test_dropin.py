"""CPU: himo_b200.dropin.install() against the reference tree -- after it, the reference's own import lines resolve to
the himo_b200 mirrors.  Runs in a subprocess: the registrations are process-wide."""
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_shims

SCRIPT = r"""
import sys
from oracle import ref_shims
ref_shims.install(); ref_shims._av2_modules()          # dztimer / av2 stand-ins and the reference on sys.path
import himo_b200.dropin as d
d.install(replace_models=True, replace_packages=True)
import himo_b200.mmcv_ext, himo_b200.chamfer3d_ext, himo_b200.chamfer3d as m
assert sys.modules["mmcv"] is himo_b200.mmcv_ext and sys.modules["chamfer3D"] is himo_b200.chamfer3d_ext
for name in ("dynamic_voxelize_forward", "hard_voxelize_forward", "dynamic_point_to_voxel_forward", "dynamic_point_to_voxel_backward"):
    assert hasattr(sys.modules["mmcv"], name)            # the loader's hasattr check, voxelize.py:12-29
import assets.cuda.chamfer3D as c                        # nsfp.py:25, selfsupervise.py:18, process.py:118
from assets.cuda.chamfer3D import nnChamferDis
assert c is m and nnChamferDis is m.nnChamferDis
import src.lossfuncs as L
from himo_b200 import lossfuncs
assert L.seflowLoss is lossfuncs.seflowLoss and L.seflowppLoss is lossfuncs.seflowppLoss
import src.models as M
from himo_b200.nsfp import NSFP
from himo_b200.fastnsf import FastNSF
from himo_b200.deflowpp import DeFlowPP
assert M.NSFP is NSFP and M.FastNSF is FastNSF and M.DeFlowPP is DeFlowPP
print("dropin ok")
"""


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_dropin_registers_mirrors_under_the_reference_names():
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
