"""Register the himo_b200 mirrors under the module names the reference imports.

    import himo_b200.dropin; himo_b200.dropin.install()

After this, the reference's own Python runs unmodified on the B200 kernels:
  * `importlib.import_module('mmcv')` in OSF/assets/cuda/mmcv/voxelize.py:12-29 and
    scatter_points.py:12-29 finds the four pybind names in `himo_b200.mmcv_ext`;
  * `import chamfer3D` in OSF/assets/cuda/chamfer3D/__init__.py:18 finds forward/backward in
    `himo_b200.chamfer3d_ext`.
"""
from __future__ import annotations

import sys


def install(replace_models: bool = False) -> None:
    from . import chamfer3d_ext, mmcv_ext
    sys.modules["mmcv"] = mmcv_ext
    sys.modules["chamfer3D"] = chamfer3d_ext
    if replace_models:
        # model-level drop-in: `hydra.utils.instantiate(cfg.model.target)` with
        # `_target_: src.models.DeFlowPP` (OSF/conf/model/deflowpp.yaml:3-9) resolves to our class
        import importlib
        models = importlib.import_module("src.models")
        from .deflowpp import DeFlowPP
        models.DeFlowPP = DeFlowPP
