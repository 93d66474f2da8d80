"""FastNSF on the B200 engine: host-side mirror of `src.models.FastNSF`
(OSF/src/models/fastnsf.py:82-222): same constructor arguments and `forward(batch) -> dict` contract
({"flow": [...], "pose_flow": [...]}), compute in libhimo_b200.so (csrc/nsf.cu).

The reference creates a fresh `Neural_Prior` from the global torch RNG for every frame pair
(fastnsf.py:108-115); here the initial parameters are an explicit input (`init_state_dict`), by default
`weights.synth_neural_prior_state_dict(seed)` with a per-frame seed, so runs are reproducible across
any GPU sharding.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int32, c_size_t, c_void_p
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, weights as W
from .deflowpp import _Timer, cal_pose0to1, rigid_flow

NUM_PARAMS = 116483


class _NsfDesc(ctypes.Structure):
    _fields_ = [
        ("pc0", c_void_p), ("n", c_int), ("n_max", c_int),
        ("D", c_void_p), ("lo", c_float * 3), ("dims", c_int32 * 3), ("grid_factor", c_float),
        ("init_params", c_void_p), ("final_params", c_void_p), ("exp_avg_out", c_void_p),
        ("planes", c_int), ("max_iters", c_int), ("lr", c_float), ("min_delta", c_float),
        ("patience", c_int), ("poll_iters", c_int),
        ("best_flow", c_void_p),
        ("iterations_out", c_void_p), ("best_loss_out", c_void_p), ("last_loss_out", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
    ]


_lib.register("himo_nsf_workspace_bytes", c_size_t, [c_int, c_int])
_lib.register("himo_nsf_volume_geometry", c_int,
              [c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p])
_lib.register("himo_nsf_dt_build", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p])
_lib.register("himo_nsf_optimize", c_int, [ctypes.POINTER(_NsfDesc), c_void_p])
_lib.register("himo_nsf_set_dt_cluster", c_int, [c_int])
_lib.register("himo_nsf_set_dt_big_tiles", c_int, [c_int])
_lib.register("himo_nsf_set_head_warp", c_int, [c_int])
_lib.register("himo_nsf_set_blocking_poll", c_int, [c_int])
_lib.register("himo_nsf_dt_pass", c_int, [c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p])


def flatten_params(sd: Dict[str, torch.Tensor], layer_size: int = 8) -> torch.Tensor:
    """Neural_Prior state_dict (nsfp_module.py:7-26) -> flat fp32 vector in state_dict order."""
    parts = []
    for i in range(layer_size):
        parts += [sd[f"nn_layers.{2 * i}.0.weight"].reshape(-1), sd[f"nn_layers.{2 * i}.0.bias"].reshape(-1)]
    parts += [sd[f"nn_layers.{2 * layer_size}.weight"].reshape(-1), sd[f"nn_layers.{2 * layer_size}.bias"].reshape(-1)]
    flat = torch.cat([p.float() for p in parts])
    assert flat.numel() == NUM_PARAMS
    return flat


def unflatten_params(flat: torch.Tensor, layer_size: int = 8) -> Dict[str, torch.Tensor]:
    sd, o = {}, 0
    dims = [3] + [128] * layer_size
    for i in range(layer_size):
        n = dims[i + 1] * dims[i]
        sd[f"nn_layers.{2 * i}.0.weight"] = flat[o:o + n].view(dims[i + 1], dims[i]); o += n
        sd[f"nn_layers.{2 * i}.0.bias"] = flat[o:o + dims[i + 1]]; o += dims[i + 1]
    sd[f"nn_layers.{2 * layer_size}.weight"] = flat[o:o + 384].view(3, 128); o += 384
    sd[f"nn_layers.{2 * layer_size}.bias"] = flat[o:o + 3]
    return sd


def volume_geometry(pc0: torch.Tensor, pc1: torch.Tensor, grid_factor: float = 10.0):
    """-> (lo [3] float32 numpy, dims (H, W, D)) of the distance volume (fastnsf.py:120-126, 34-36)."""
    dev = pc0.device
    lo = (c_float * 3)()
    dims = (c_int32 * 3)()
    with torch.cuda.device(dev):
        scratch = _lib.workspace.get(4096, dev)
        st = _lib.lib().himo_nsf_volume_geometry(_lib.ptr(pc0), pc0.shape[0], _lib.ptr(pc1), pc1.shape[0],
                                                 c_float(grid_factor), lo, dims, _lib.ptr(scratch),
                                                 _lib.stream_ptr(dev))
    _lib.check(st, "himo_nsf_volume_geometry")
    return np.array(list(lo), np.float32), tuple(int(v) for v in dims)


def dt_build(pc1: torch.Tensor, lo: np.ndarray, dims, grid_factor: float = 10.0) -> torch.Tensor:
    """Distance volume D [H,W,D] f32 (DT.__init__, fastnsf.py:30-57)."""
    dev = pc1.device
    D = torch.empty(dims, dtype=torch.float32, device=dev)
    lo_c = (c_float * 3)(*[float(v) for v in lo])
    dims_c = (c_int32 * 3)(*[int(v) for v in dims])
    with torch.cuda.device(dev):
        st = _lib.lib().himo_nsf_dt_build(_lib.ptr(pc1), pc1.shape[0], lo_c, dims_c, c_float(grid_factor),
                                          _lib.ptr(D), _lib.stream_ptr(dev))
    _lib.check(st, "himo_nsf_dt_build")
    return D


class FastNSF:
    """Drop-in for `src.models.FastNSF` (conf/model/fastnsf.yaml)."""

    def __init__(self, filter_size=128, act_fn="relu", layer_size=8, grid_factor=10.0, itr_num=5000, lr=8e-3,
                 min_delta=0.00005, early_patience=30, verbose=False,
                 point_cloud_range=(-51.2, -51.2, -3, 51.2, 51.2, 3), init_weight=True,
                 precision: str = "fp32", device="cuda", seed: int = 0):
        if filter_size != 128 or layer_size != 8 or act_fn != "relu":
            raise NotImplementedError("himo_b200.FastNSF implements the 8x128 ReLU prior of conf/model/fastnsf.yaml")
        self.grid_factor = float(grid_factor)
        self.iteration_num = int(itr_num)
        self.lr, self.min_delta, self.early_patience = float(lr), float(min_delta), int(early_patience)
        self.point_cloud_range = list(point_cloud_range)
        self.planes = 2 if precision == "fp32" else 1
        self.device = torch.device(device)
        self.timer = _Timer()
        self.seed = seed
        self._frame = 0
        self._ws = None
        self._n_max = 0
        self.last_info: Dict = {}
        _lib.lib()

    def eval(self):
        return self

    def to(self, device):
        return self

    def range_limit_(self, pc: torch.Tensor):
        r = self.point_cloud_range
        mask = ((pc[:, 0] >= r[0]) & (pc[:, 0] <= r[3]) & (pc[:, 1] >= r[1]) & (pc[:, 1] <= r[4]) &
                (pc[:, 2] >= r[2]) & (pc[:, 2] <= r[5]))
        return pc[mask], mask

    def optimize(self, pc0: torch.Tensor, pc1: torch.Tensor, init_state_dict: Optional[Dict] = None,
                 D: Optional[torch.Tensor] = None, lo=None, dims=None, return_params: bool = False) -> Dict:
        """pc0 (ego-compensated) / pc1: [N,3] f32 CUDA, already range-limited."""
        dev = pc0.device
        pc0, pc1 = pc0.contiguous(), pc1.contiguous()
        n = pc0.shape[0]
        if init_state_dict is None:
            init_state_dict = W.synth_neural_prior_state_dict(self.seed * 1000003 + self._frame)
        self._frame += 1
        init = flatten_params(init_state_dict).to(dev)
        if D is None:
            lo, dims = volume_geometry(pc0, pc1, self.grid_factor)
            D = dt_build(pc1, lo, dims, self.grid_factor)
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._ws is None or n > self._n_max:
                self._n_max = max(n, 4096)
                self._ws = torch.empty(L.himo_nsf_workspace_bytes(self._n_max, self.planes), dtype=torch.uint8,
                                       device=dev)
            best = torch.empty((n, 3), dtype=torch.float32, device=dev)
            final = torch.empty(NUM_PARAMS, dtype=torch.float32, device=dev) if return_params else None
            exp_avg = torch.empty(NUM_PARAMS, dtype=torch.float32, device=dev) if return_params else None
            iters, bl, ll = c_int32(0), c_float(0), c_float(0)
            d = _NsfDesc()
            d.pc0, d.n, d.n_max = pc0.data_ptr(), n, self._n_max
            d.D = D.data_ptr()
            for k in range(3):
                d.lo[k] = float(lo[k])
                d.dims[k] = int(dims[k])
            d.grid_factor = self.grid_factor
            d.init_params = init.data_ptr()
            d.final_params = final.data_ptr() if final is not None else None
            d.exp_avg_out = exp_avg.data_ptr() if exp_avg is not None else None
            d.planes, d.max_iters, d.lr = self.planes, self.iteration_num, self.lr
            d.min_delta, d.patience, d.poll_iters = self.min_delta, self.early_patience, 8
            d.best_flow = best.data_ptr()
            d.iterations_out = ctypes.addressof(iters)
            d.best_loss_out = ctypes.addressof(bl)
            d.last_loss_out = ctypes.addressof(ll)
            d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
            st = L.himo_nsf_optimize(ctypes.byref(d), _lib.stream_ptr(dev))
        _lib.check(st, "himo_nsf_optimize")
        out = {"flow": best, "loss": float(bl.value), "last_loss": float(ll.value), "iterations": int(iters.value),
               "D": D, "lo": lo, "dims": dims}
        if return_params:
            out["params"] = final
            out["exp_avg"] = exp_avg
        self.last_info = {k: out[k] for k in ("loss", "iterations")}
        return out

    def forward(self, batch: Dict, init_state_dicts: Optional[List[Dict]] = None) -> Dict[str, List[torch.Tensor]]:
        """fastnsf.py:180-222.  `init_state_dicts` (one Neural_Prior state_dict per batch item) replaces the
        per-frame seeded initial weights -- the reference draws them from the global torch RNG (fastnsf.py:110-113)."""
        flows, pose_flows = [], []
        for b in range(len(batch["pose0"])):
            pc0, pc1 = batch["pc0"][b], batch["pc1"][b]
            sel0, rm0 = self.range_limit_(pc0)
            sel1, _ = self.range_limit_(pc1)
            if "ego_motion" in batch:
                T = torch.as_tensor(batch["ego_motion"][b]).detach().cpu().float()
            else:
                T = cal_pose0to1(batch["pose0"][b], batch["pose1"][b])
            pf = rigid_flow(sel0.contiguous(), T)
            tr0 = sel0 + pf                                        # fastnsf.py:199 (warp), :201 (pose flow)
            res = self.optimize(tr0, sel1, init_state_dict=init_state_dicts[b] if init_state_dicts else None)
            final = torch.zeros_like(pc0)
            final[rm0] = res["flow"]
            flows.append(final)
            pose_flows.append(pf)
        return {"flow": flows, "pose_flow": pose_flows}

    __call__ = forward
