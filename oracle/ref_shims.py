"""oracle/ref_shims.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Lets the reference's OWN Python (OpenSceneFlow `src.models.*`, HiMo `utils`, `tools/test/score.py`)
import and run on CPU in the build container, where /root/reference exists.  Five leaf
dependencies are absent from the image and are shimmed through sys.modules:

  dztimer      no-op hierarchical timer (deflow.py:112-113, fastnsf.py:100-101 only start/stop it)
  mmcv         the 4 pybind functions, backed by oracle/leaf_ops.c on CPU tensors
               (the reference registers CUDA implementations only: cudabind.cpp:57-60,101-104)
  chamfer3D    forward / backward, backed by oracle/leaf_ops.c
  FastGeodis   generalised_geodesic3d, backed by the PARITY-UNPINNED raster restatement
  (h5py, hydra, lightning are not needed for the model classes)

Nothing here is reachable from himo_b200/.  The GPU box has no /root/reference: only
tests/golden/make_golden.py and the `-m "not gpu"` pinning tests call install().
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import leaf

REFERENCE_ROOT = os.environ.get("HIMO_REFERENCE_ROOT", "/root/reference")
OSF_ROOT = os.path.join(REFERENCE_ROOT, "OpenSceneFlow")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(OSF_ROOT, "src", "models"))


class _NoTimer:
    """dztimer.Timing stand-in: timer[i][j].start(name)/stop()/print() are all no-ops."""

    def __getitem__(self, _):
        return self

    def start(self, *_a, **_k):
        return None

    def stop(self, *_a, **_k):
        return None

    def print(self, *_a, **_k):
        return None


def _mmcv_module(accum: str) -> types.ModuleType:
    m = types.ModuleType("mmcv")

    def dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3):
        out = leaf.dynamic_voxelize(points.detach().cpu().numpy(), voxel_size.numpy(),
                                    coors_range.numpy())
        coors.copy_(torch.from_numpy(out))

    def hard_voxelize_forward(*a, **k):
        raise NotImplementedError

    def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
        vf, vc, p2v, cnt = leaf.dynamic_point_to_voxel(feats.detach().cpu().numpy(),
                                                       coors.cpu().numpy(), reduce_type, accum)
        return [torch.from_numpy(vf), torch.from_numpy(vc).to(coors.dtype), torch.from_numpy(p2v),
                torch.from_numpy(cnt)]

    def dynamic_point_to_voxel_backward(*a, **k):
        raise NotImplementedError

    m.dynamic_voxelize_forward = dynamic_voxelize_forward
    m.hard_voxelize_forward = hard_voxelize_forward
    m.dynamic_point_to_voxel_forward = dynamic_point_to_voxel_forward
    m.dynamic_point_to_voxel_backward = dynamic_point_to_voxel_backward
    return m


def _chamfer_module() -> types.ModuleType:
    m = types.ModuleType("chamfer3D")

    def forward(pc0, pc1, dist0, dist1, idx0, idx1):
        d0, d1, i0, i1 = leaf.chamfer_forward(pc0.detach().cpu().numpy(), pc1.detach().cpu().numpy())
        dist0.copy_(torch.from_numpy(d0)); dist1.copy_(torch.from_numpy(d1))
        idx0.copy_(torch.from_numpy(i0)); idx1.copy_(torch.from_numpy(i1))
        return 1

    def backward(pc0, pc1, idx0, idx1, g0, g1, gp0, gp1):
        a, b = leaf.chamfer_backward(pc0.detach().cpu().numpy(), pc1.detach().cpu().numpy(),
                                     idx0.cpu().numpy(), idx1.cpu().numpy(),
                                     g0.detach().cpu().numpy(), g1.detach().cpu().numpy())
        gp0.add_(torch.from_numpy(a)); gp1.add_(torch.from_numpy(b))
        return 1

    m.forward = forward
    m.backward = backward
    return m


def _fastgeodis_module() -> types.ModuleType:
    m = types.ModuleType("FastGeodis")

    def generalised_geodesic3d(image, softmask, spacing, v, lamb, iterations):
        if float(lamb) != 0.0:
            raise NotImplementedError("only the lamb=0 (Euclidean) call of fastnsf.py:55-57")
        d = (softmask[0, 0].detach().cpu().numpy().astype(np.float32) * np.float32(v))
        out = leaf.geodesic3d_euclid(d, spacing, int(iterations))
        return torch.from_numpy(out)[None, None].to(image.device)

    m.generalised_geodesic3d = generalised_geodesic3d
    return m


def install(accum: str = "f32_seq") -> None:
    """Register the shims and put the OpenSceneFlow root on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    dz = types.ModuleType("dztimer")
    dz.Timing = _NoTimer
    sys.modules["dztimer"] = dz
    sys.modules["mmcv"] = _mmcv_module(accum)
    sys.modules["chamfer3D"] = _chamfer_module()
    sys.modules["FastGeodis"] = _fastgeodis_module()
    if OSF_ROOT not in sys.path:
        sys.path.insert(0, OSF_ROOT)


def import_models():
    """-> the reference `src.models` package (DeFlowPP, FastNSF, NSFP ...)."""
    install()
    import importlib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module("src.models")


def import_lossfuncs():
    """-> the reference `src.lossfuncs.selfsupervise` module (seflowLoss, seflowppLoss) over the chamfer3D shim."""
    install()
    _av2_modules()          # src/lossfuncs/__init__.py pulls in supervise.py -> av2_eval.py
    import importlib
    return importlib.import_module("src.lossfuncs.selfsupervise")


def import_submit_writers():
    """-> (write_output_file, zip_res): the reference's AV2 leaderboard writers, OSF/src/utils/av2_eval.py:758-801 and
    OSF/src/utils/mics.py:312-344.  mics.py imports h5py at module level without using it here: an empty stand-in is
    placed in sys.modules for the duration of the import only."""
    install()
    _av2_modules()
    import importlib
    import contextlib
    import io
    fake = "h5py" not in sys.modules and importlib.util.find_spec("h5py") is None
    if fake:
        sys.modules["h5py"] = types.ModuleType("h5py")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ae = importlib.import_module("src.utils.av2_eval")
            mics = importlib.import_module("src.utils.mics")
    finally:
        if fake:
            del sys.modules["h5py"]
    return ae.write_output_file, mics.zip_res


def _av2_modules() -> None:
    """`av2` and `rich` are imported at module level by OSF/src/utils/av2_eval.py (:19,:26,:77-80) but only their
    category enum matters for the metric arithmetic; the enum is rebuilt from the table HiMo's scorer carries
    (tools/test/score.py:29-60, identical order)."""
    import enum
    import importlib.util
    spec = importlib.util.spec_from_file_location("himo_ref_score_for_av2", os.path.join(REFERENCE_ROOT, "tools", "test", "score.py"))
    score = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(score)
    names = [k for k, v in sorted(score.CATEGORY_TO_INDEX.items(), key=lambda kv: kv[1]) if k != "NONE"]
    cats = enum.Enum("AnnotationCategories", {n: n for n in names}, type=str)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    if not hasattr(np, "NaN"):
        np.NaN = np.nan          # eval_metric.py:143 predates NumPy 2.0
    mod("av2"); mod("av2.datasets"); mod("av2.datasets.sensor")
    mod("av2.datasets.sensor.constants", AnnotationCategories=cats)
    mod("av2.geometry"); mod("av2.geometry.geometry"); mod("av2.geometry.se3", SE3=object)
    mod("av2.utils"); mod("av2.utils.typing", NDArrayFloat=np.ndarray, NDArrayBool=np.ndarray, NDArrayInt=np.ndarray)
    mod("av2.utils.io", read_feather=None)
    mod("rich"); mod("rich.progress", track=lambda it, **_k: it)
    for name in ("torch",):
        pass
    # av2_eval.py annotates with BoolTensor (torch) further down; provide it through typing if missing
    import builtins
    if not hasattr(builtins, "BoolTensor"):
        builtins.BoolTensor = torch.BoolTensor


def import_eval_metric():
    """-> the reference `src.utils.eval_metric` module (evaluate_leaderboard*, evaluate_ssf, OfficialMetrics)."""
    install()
    _av2_modules()
    import importlib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module("src.utils.eval_metric")
