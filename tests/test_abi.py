"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/himo_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "himo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(himo_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    assert "himo_dynamic_voxelize_forward" in syms
    assert "himo_dynamic_point_to_voxel_forward" in syms
    assert "himo_chamfer_forward" in syms


def test_library_exports_every_declared_symbol():
    from himo_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m himo_b200.build`"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/himo_b200.h but not exported: {missing}"


def test_abi_version_and_status_strings():
    from himo_b200 import _lib
    L = _lib.lib()
    assert L.himo_abi_version() == 1
    assert L.himo_status_string(0) == b"ok"
    assert b"workspace" in L.himo_status_string(-2)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under himo_b200/ may import it."""
    pkg = os.path.join(ROOT, "himo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_cuda_required_error():
    import pytest
    import torch
    from himo_b200 import mmcv_ext
    pts = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        mmcv_ext.dynamic_voxelize_forward(pts, torch.tensor([0.2, 0.2, 6.0]),
                                          torch.tensor([-1.0, -1, -1, 1, 1, 1]),
                                          torch.zeros(4, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        mmcv_ext.dynamic_point_to_voxel_forward(pts, torch.zeros(4, 3, dtype=torch.int32), "median")
