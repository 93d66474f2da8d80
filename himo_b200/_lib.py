"""ctypes loader for libhimo_b200.so (the C-ABI CUDA library).

There is no CPU fallback anywhere in this package: if the library is missing or a call
returns a non-zero status the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhimo_b200.so")
_lib = None


def _sig(lib, name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


def lib() -> ctypes.CDLL:
    """Load (once) and return the C-ABI library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m himo_b200.build` "
            "(himo_b200 has no CPU or PyTorch fallback path)")
    L = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    _sig(L, "himo_abi_version", c_int, [])
    _sig(L, "himo_status_string", c_char_p, [c_int])
    _sig(L, "himo_dynamic_voxelize_forward", c_int, [P, c_int, c_int, P, P, P, P])
    _sig(L, "himo_dynamic_point_to_voxel_workspace_bytes", c_size_t, [c_int, c_int, P])
    _sig(L, "himo_dynamic_point_to_voxel_forward", c_int,
         [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P])
    _sig(L, "himo_chamfer_workspace_bytes", c_size_t, [c_int, c_int])
    _sig(L, "himo_chamfer_forward", c_int,
         [P, c_int, P, c_int, P, P, P, P, c_float, P, c_size_t, P])
    _sig(L, "himo_chamfer_forward_radius", c_int,
         [P, c_int, P, c_int, P, P, P, P, c_float, P, c_size_t, P])
    _sig(L, "himo_chamfer_backward", c_int, [P, c_int, P, c_int, P, P, P, P, P, P, P])
    for name, restype, argtypes in _LATE_SIGS:
        if hasattr(L, name):
            _sig(L, name, restype, argtypes)
    _lib = L
    return L


# signatures registered by the modules that own the entry points (embed / conv / decoder / nsf)
_LATE_SIGS: list = []


def register(name, restype, argtypes):
    _LATE_SIGS.append((name, restype, argtypes))
    if _lib is not None and hasattr(_lib, name):
        _sig(_lib, name, restype, argtypes)


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().himo_status_string(int(status)).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def ptr(t):
    """Device (or host) pointer of a tensor as c_void_p; None -> NULL."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device=None) -> c_void_p:
    """The caller's current CUDA stream, as the reference's ops use
    (at::cuda::getCurrentCUDAStream(), e.g. OSF/assets/cuda/chamfer3D/chamfer3D.cu:88)."""
    if _raw_stream is not None:
        idx = device.index if isinstance(device, torch.device) else device
        if isinstance(idx, int):
            return c_void_p(_raw_stream(idx))          # ~0.3 us instead of ~2 us through the Stream object
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def on_device(device):
    """`with on_device(t.device):` -- the device guard of the reference's ops (CUDAGuard, voxelization_cuda.cu:253)
    without the ~4 us of torch.cuda.device when `device` is already current (the common, single-GPU-per-process case)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (himo_b200 implements CUDA only, "
                           "like the reference ops: OSF/assets/cuda/mmcv/cudabind.cpp:57-60)")


class Workspace:
    """Grow-only per-device scratch buffer handed to the C ABI as `workspace`."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


workspace = Workspace()
