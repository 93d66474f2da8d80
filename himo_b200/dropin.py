"""Register the himo_b200 mirrors under the module names the reference imports.

    import himo_b200.dropin; himo_b200.dropin.install()

After this, the reference's own Python runs unmodified on the B200 kernels:
  * `importlib.import_module('mmcv')` in OSF/assets/cuda/mmcv/voxelize.py:12-29 and
    scatter_points.py:12-29 finds the four pybind names in `himo_b200.mmcv_ext`;
  * `import chamfer3D` in OSF/assets/cuda/chamfer3D/__init__.py:18 finds forward/backward in
    `himo_b200.chamfer3d_ext`;
  * with `replace_packages=True`, `from assets.cuda.chamfer3D import nnChamferDis` (nsfp.py:25,
    selfsupervise.py:18, process.py:118) resolves to `himo_b200.chamfer3d` (radius-pruned truncated losses) and
    `src.lossfuncs.selfsupervise.{seflowLoss, seflowppLoss}` to the segmented forms in `himo_b200.lossfuncs`.
"""
from __future__ import annotations

import sys


def install(replace_models: bool = False, replace_packages: bool = False) -> None:
    from . import chamfer3d_ext, mmcv_ext
    sys.modules["mmcv"] = mmcv_ext
    sys.modules["chamfer3D"] = chamfer3d_ext
    if replace_packages:
        import importlib
        from . import chamfer3d, lossfuncs
        sys.modules["assets.cuda.chamfer3D"] = chamfer3d
        try:
            ss = importlib.import_module("src.lossfuncs.selfsupervise")
            ss.seflowLoss, ss.seflowppLoss = lossfuncs.seflowLoss, lossfuncs.seflowppLoss
            pkg = importlib.import_module("src.lossfuncs")
            pkg.seflowLoss, pkg.seflowppLoss = lossfuncs.seflowLoss, lossfuncs.seflowppLoss
        except ImportError:
            pass        # the reference tree is not on sys.path: only the module alias applies
    if replace_models:
        # model-level drop-in: `hydra.utils.instantiate(cfg.model.target)` with
        # `_target_: src.models.DeFlowPP` (OSF/conf/model/deflowpp.yaml:3-9) resolves to our class
        import importlib
        models = importlib.import_module("src.models")
        from .deflowpp import DeFlowPP
        from .fastnsf import FastNSF
        from .nsfp import NSFP
        models.DeFlowPP, models.FastNSF, models.NSFP = DeFlowPP, FastNSF, NSFP
