"""himo_b200 -- B200 (sm_100a) engine for the HiMo / OpenSceneFlow per-frame-pair hot path.

Only what the path needs: csrc/ (CUDA kernels + C ABI), the ctypes
loader, and host-side mirrors of the reference's operator / model interfaces.
"""
__version__ = "0.1.0"
