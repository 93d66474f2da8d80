"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Compile the reference's OWN CUDA extensions for sm_100a from the sources where they lie under /root/reference
(nothing is copied into the repo) into oracle/_ref/:

    oracle/_ref/chamfer3D.so   <- OSF/assets/cuda/chamfer3D/{chamfer3D_cuda.cpp, chamfer3D.cu}   (pybind: forward, backward)
    oracle/_ref/mmcv.so        <- OSF/assets/cuda/mmcv/{pybind, cudabind, voxelization, scatter_points}.cpp + the two .cu

These are the kernels the reference itself runs on a GPU.  They have no CPU build, so here (no GPU) they only compile;
on the GPU box `tests/test_gpu_vs_reference_kernels.py` loads them and checks our kernels against them on the same
inputs -- the parity pin the CPU restatements in oracle/leaf_ops.c cannot give -- and `scripts/bench_ref_kernels.py`
times them beside ours.  oracle/_ref/ is git-ignored and travels to the GPU box with the snapshot.

    python -m oracle.build_ref            # needs /root/reference; ~3 min
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("HIMO_REFERENCE_ROOT", "/root/reference")
CUDA_DIR = os.path.join(REF, "OpenSceneFlow", "assets", "cuda")

EXTENSIONS = {
    "chamfer3D": ["chamfer3D/chamfer3D_cuda.cpp", "chamfer3D/chamfer3D.cu"],
    "mmcv": ["mmcv/pybind.cpp", "mmcv/cudabind.cpp", "mmcv/voxelization.cpp", "mmcv/scatter_points.cpp",
             "mmcv/voxelization_cuda.cu", "mmcv/scatter_points_cuda.cu"],
}
DEFINES = ["-DCCCL_IGNORE_DEPRECATED_CUDA_BELOW_12", "-DTHRUST_IGNORE_CUB_VERSION_CHECK"]      # the reference's setup.py


def available() -> bool:
    return os.path.isdir(CUDA_DIR)


def built(name: str) -> str:
    p = os.path.join(OUT, name + ".so")
    return p if os.path.exists(p) else ""


def build(names=None, verbose: bool = False) -> dict:
    from torch.utils import cpp_extension
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.makedirs(OUT, exist_ok=True)
    done = {}
    for name in names or EXTENSIONS:
        if built(name):
            done[name] = built(name)
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        extra = ["-DMMCV_WITH_CUDA"] if name == "mmcv" else []
        cpp_extension.load(name=name, sources=[os.path.join(CUDA_DIR, s) for s in EXTENSIONS[name]],
                           extra_cflags=DEFINES + extra, extra_cuda_cflags=DEFINES + extra + ["-lineinfo"],
                           extra_include_paths=[os.path.join(CUDA_DIR, name)], build_directory=bdir,
                           with_cuda=True, is_python_module=False, verbose=verbose)
        shutil.copy2(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
        shutil.rmtree(bdir, ignore_errors=True)
        done[name] = os.path.join(OUT, name + ".so")
    return done


def load(name: str):
    """Import oracle/_ref/<name>.so as a Python module (GPU box: the reference kernels themselves)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    path = built(name)
    if not path:
        raise FileNotFoundError(f"oracle/_ref/{name}.so not built (python -m oracle.build_ref)")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if not available():
        sys.exit("needs /root/reference")
    print(build(sys.argv[1:] or None, verbose=True))
