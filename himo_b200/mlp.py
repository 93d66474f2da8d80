"""The 3 -> 128 x 8 -> 3 ReLU coordinate prior (`Neural_Prior`, OSF/src/models/basic/nsfp_module.py:7-47) on the library's
own kernels, as a differentiable module for losses computed OUTSIDE the library (NSFP's two-network Chamfer loss).

`PriorMLP(x)` is an autograd node: forward = `himo_mlp_forward` (layer 0 on CUDA cores, seven 128x128 layers on the
tcgen05 GEMM, 3-wide head), backward = `himo_mlp_backward` (weight gradients by split-K GEMMs, ReLU masks in the GEMM
epilogues, gradient w.r.t. the input when it is needed).  The parameters, their gradients and the Adam moments never
become torch tensors: they live in the C-ABI workspace and `adam_step()` is `himo_mlp_adam_step`.  CUDA only.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_longlong, c_size_t, c_void_p
from typing import Dict, Optional

import torch

from . import _lib
from .fastnsf import NUM_PARAMS, flatten_params, unflatten_params

_P = c_void_p
_lib.register("himo_mlp_init", c_int, [_P, c_size_t, c_int, c_int, _P, _P])
_lib.register("himo_mlp_forward", c_int, [_P, c_size_t, c_int, c_int, _P, _P, c_int, _P, _P])
_lib.register("himo_mlp_backward", c_int, [_P, c_size_t, c_int, c_int, _P, c_int, _P, _P, _P])
_lib.register("himo_mlp_adam_step", c_int, [_P, c_size_t, c_int, c_int, _P, c_int, c_float, _P])
_lib.register("himo_mlp_control", c_int, [_P, c_size_t, c_int, c_int, _P, c_float, c_int, _P, _P, c_longlong, _P])
_lib.register("himo_mlp_read_state", c_int, [_P, c_size_t, c_int, c_int, _P, _P, _P, _P])


class _Apply(torch.autograd.Function):
    # `anchor` is a one-element tensor that requires grad: the parameters live in the C-ABI workspace, not in torch, so
    # without it the output of a network fed with a constant input (NSFP's `net(pc0)`) would not enter the graph
    @staticmethod
    def forward(ctx, x: torch.Tensor, anchor: torch.Tensor, mlp: "PriorMLP"):
        ctx.mlp, ctx.n, ctx.need_dx = mlp, x.shape[0], bool(ctx.needs_input_grad[0])
        return mlp._forward(x.detach())

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        return ctx.mlp._backward(grad_out.contiguous(), ctx.n, ctx.need_dx), None, None


class PriorMLP:
    """One network.  `follow`: another PriorMLP whose control block (iteration count, stop flag) this one obeys."""

    def __init__(self, init_state_dict: Dict[str, torch.Tensor], n_max: int, device, precision: str = "fp32",
                 follow: Optional["PriorMLP"] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("himo_b200.mlp.PriorMLP implements CUDA only (no CPU fallback)")
        self.planes = 2 if precision == "fp32" else 1
        self.n_max = max(int(n_max), 4096)
        self.L = _lib.lib()
        with _lib.on_device(self.device):
            self.ws = torch.empty(self.L.himo_nsf_workspace_bytes(self.n_max, self.planes), dtype=torch.uint8,
                                  device=self.device)
            init = flatten_params(init_state_dict).to(self.device)
            _lib.check(self.L.himo_mlp_init(_lib.ptr(self.ws), self.ws.numel(), self.n_max, self.planes, _lib.ptr(init),
                                            _lib.stream_ptr(self.device)), "himo_mlp_init")
        if follow is not None and (follow.n_max != self.n_max or follow.planes != self.planes):
            raise ValueError("a following network must share n_max and precision with its leader")
        self._ctl = _lib.ptr(follow.ws) if follow is not None else c_void_p(0)
        self._n = 0
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)

    def _args(self):
        return _lib.ptr(self.ws), self.ws.numel(), self.n_max, self.planes

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != 3:
            raise RuntimeError("x must be a float32 [N,3] CUDA tensor")
        if x.shape[0] > self.n_max or x.shape[0] == 0:
            raise RuntimeError(f"PriorMLP was sized for 1..{self.n_max} points, got {x.shape[0]}")
        return _Apply.apply(x.contiguous(), self._anchor, self)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(x)
        self._n = x.shape[0]
        with _lib.on_device(self.device):
            _lib.check(self.L.himo_mlp_forward(*self._args(), self._ctl, _lib.ptr(x), x.shape[0], _lib.ptr(out),
                                               _lib.stream_ptr(self.device)), "himo_mlp_forward")
        return out

    def _backward(self, d_out: torch.Tensor, n: int, need_dx: bool) -> Optional[torch.Tensor]:
        dx = torch.empty((n, 3), dtype=torch.float32, device=self.device) if need_dx else None
        with _lib.on_device(self.device):
            _lib.check(self.L.himo_mlp_backward(*self._args(), self._ctl, n, _lib.ptr(d_out), _lib.ptr(dx),
                                                _lib.stream_ptr(self.device)), "himo_mlp_backward")
        return dx

    def adam_step(self, lr: float) -> None:
        with _lib.on_device(self.device):
            _lib.check(self.L.himo_mlp_adam_step(*self._args(), self._ctl, max(self._n, 1), c_float(lr),
                                                 _lib.stream_ptr(self.device)), "himo_mlp_adam_step")

    def control(self, loss: torch.Tensor, min_delta: float, patience: int, out: Optional[torch.Tensor] = None,
                best_out: Optional[torch.Tensor] = None) -> None:
        """One loss evaluation of the optimisation loop, on the device (see himo_mlp_control)."""
        loss = loss.detach().reshape(1).float()
        count = out.numel() if out is not None else 0
        with _lib.on_device(self.device):
            _lib.check(self.L.himo_mlp_control(*self._args(), _lib.ptr(loss), c_float(min_delta), int(patience),
                                               _lib.ptr(out.detach() if out is not None else None), _lib.ptr(best_out),
                                               c_longlong(count), _lib.stream_ptr(self.device)), "himo_mlp_control")

    def read_state(self, with_params: bool = False) -> Dict:
        st = (c_float * 4)()
        params = torch.empty(NUM_PARAMS, dtype=torch.float32, device=self.device) if with_params else None
        m = torch.empty(NUM_PARAMS, dtype=torch.float32, device=self.device) if with_params else None
        with _lib.on_device(self.device):
            _lib.check(self.L.himo_mlp_read_state(*self._args(), st, _lib.ptr(params), _lib.ptr(m),
                                                  _lib.stream_ptr(self.device)), "himo_mlp_read_state")
        out = {"stop": bool(st[0]), "iterations": int(st[1]), "best_loss": float(st[2]), "loss": float(st[3])}
        if with_params:
            out["params"], out["exp_avg"] = params, m
        return out

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return unflatten_params(self.read_state(with_params=True)["params"].cpu())
