"""NSFP: work-alike of `src.models.NSFP`, OSF/src/models/nsfp.py:28-186.

Per frame pair two coordinate MLPs are fitted at run time: `net` maps the ego-compensated pc0 to a flow, `net_inv` maps
the displaced points back, and the loss is the NSFP-truncated bidirectional Chamfer distance of both (nsfp.py:48-72).
Everything heavy runs on the library's own kernels: the two 8x128 MLPs forward / backward and their Adam steps
(`himo_b200.mlp.PriorMLP` = himo_mlp_forward / _backward / _adam_step: tcgen05 GEMMs with fp32-class split-fp16 operands,
the same kernels FastNSF uses), the four exact 1-NN searches and their gradient scatter per iteration
(`himo_b200.chamfer3d`, the truncation radius sqrt(2) m pruning the search where the reference's kernel is O(N*M)).
torch only glues them (autograd bookkeeping and a handful of element-wise ops on [N,3] tensors).  CUDA only.

The control flow keeps the reference's meaning -- best flow so far under `loss <= best`, early stopping on
`loss < best - min_delta` in float32 with `patience` bad iterations, a NaN loss stops at once, at least one iteration --
but runs on the DEVICE (himo_mlp_control): the reference reads the loss back every iteration (`loss.item()`, nsfp.py:
108-112); here the host polls a stop flag every `poll_iters` iterations and a stopped network's kernels are no-ops.
Initial weights: like the reference, a fresh default-initialised `Linear` stack drawn from the global torch CPU RNG
(nsfp.py:78-81, so `torch.manual_seed(s)` before the call reproduces the reference's network), `net_inv` an exact copy;
`init_state_dict` (a reference `Neural_Prior.state_dict()`) overrides it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .deflowpp import _Timer, cal_pose0to1, rigid_flow

NSFP_TRUNCATE_SQ = 2        # chamfer3D/__init__.py:78


class _Prior(nn.Module):
    """State-dict compatible with the reference's Neural_Prior (basic/nsfp_module.py:7-47): `nn_layers.{2i}.0` hidden
    Linear layers (wrapped in a Sequential there), activations at the odd slots, `nn_layers.{2L}` the 3-wide head."""

    def __init__(self, dim_x=3, filter_size=128, act_fn="relu", layer_size=8):
        super().__init__()
        if layer_size < 1:
            raise NotImplementedError("layer_size >= 1")
        act = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid}[act_fn]
        mods: List[nn.Module] = []
        for i in range(layer_size):
            mods += [nn.Sequential(nn.Linear(dim_x if i == 0 else filter_size, filter_size)), act()]
        mods.append(nn.Linear(filter_size, dim_x))
        self.nn_layers = nn.ModuleList(mods)

    def forward(self, x):
        for m in self.nn_layers:
            x = m(x)
        return x


class _EarlyStop:
    """EarlyStopping(mode='min', percentage=False) of basic/nsfp_module.py:50-96, on float32 scalars."""

    def __init__(self, patience: int, min_delta: float):
        self.patience, self.min_delta = int(patience), np.float32(min_delta)
        self.best: Optional[np.float32] = None
        self.bad = 0

    def step(self, loss: float) -> bool:
        if self.patience == 0:
            return False
        v = np.float32(loss)
        if self.best is None:
            self.best = v
            return False
        if np.isnan(v):
            return True
        if v < self.best - self.min_delta:
            self.best, self.bad = v, 0
        else:
            self.bad += 1
        return self.bad >= self.patience


class NSFP:
    """Drop-in for `src.models.NSFP` (conf/model/nsfp.yaml)."""

    def __init__(self, filter_size=128, act_fn="relu", layer_size=8, itr_num=5000, lr=8e-3, min_delta=0.00005,
                 early_patience=30, verbose=False, point_cloud_range=(-51.2, -51.2, -3, 51.2, 51.2, 3), chamfer=None):
        self.filter_size, self.act_fn, self.layer_size = filter_size, act_fn, layer_size
        self.iteration_num, self.lr = int(itr_num), float(lr)
        self.min_delta, self.early_patience = float(min_delta), int(early_patience)
        self.verbose = verbose
        self.point_cloud_range = list(point_cloud_range)
        self._chamfer = chamfer
        self.timer = _Timer()          # the runner touches model.timer[...] (OSF/src/runner.py:141-143, 296)
        self.last_info: Dict = {}

    def eval(self):
        return self

    def to(self, device):
        return self

    def _ch(self):
        if self._chamfer is None:
            from .chamfer3d import nnChamferDis        # needs the CUDA library; raises if it is missing
            self._chamfer = nnChamferDis()
        return self._chamfer

    def range_limit_(self, pc: torch.Tensor):
        r = self.point_cloud_range
        mask = ((pc[:, 0] >= r[0]) & (pc[:, 0] <= r[3]) & (pc[:, 1] >= r[1]) & (pc[:, 1] <= r[4]) &
                (pc[:, 2] >= r[2]) & (pc[:, 2] <= r[5]))
        return pc[mask], mask

    def optimize(self, pc0: torch.Tensor, pc1: torch.Tensor, init_state_dict: Optional[Dict] = None,
                 poll_iters: int = 8) -> Dict:
        """pc0 (ego-compensated) / pc1: [N,3] f32 CUDA, already range-limited.  -> {'loss', 'flow', 'iterations'}."""
        from . import _lib
        from .mlp import PriorMLP
        _lib.require_cuda(pc0, "pc0")
        _lib.require_cuda(pc1, "pc1")
        ch = self._ch()
        dev = pc0.device
        if init_state_dict is None:      # the reference's fresh default-initialised network, from the global torch CPU RNG
            init_state_dict = _Prior(filter_size=self.filter_size, act_fn=self.act_fn, layer_size=self.layer_size).state_dict()
        if self.filter_size != 128 or self.layer_size != 8 or self.act_fn != "relu":
            raise NotImplementedError("himo_b200.NSFP implements the 8x128 ReLU prior of conf/model/nsfp.yaml")
        pc0 = pc0.detach().contiguous()
        pc1 = pc1.detach().contiguous()
        n = pc0.shape[0]
        net = PriorMLP(init_state_dict, n, dev)
        net_inv = PriorMLP(init_state_dict, n, dev, follow=net)          # copy.deepcopy(net), nsfp.py:82
        best_flow = torch.zeros_like(pc0)
        state = {"stop": False, "iterations": 0}
        with torch.inference_mode(False), torch.enable_grad():
            for it in range(self.iteration_num):
                flow = net(pc0)
                moved = pc0 + flow
                back = moved - net_inv(moved)
                loss = ch.truncated_dis(moved, pc1, NSFP_TRUNCATE_SQ) + ch.truncated_dis(back, pc0, NSFP_TRUNCATE_SQ)
                # nsfp.py:104-113 on the device: count the iteration, keep the best flow, EarlyStopping.step
                net.control(loss, self.min_delta, self.early_patience, out=flow, best_out=best_flow)
                loss.backward()
                net.adam_step(self.lr)
                net_inv.adam_step(self.lr)
                if (it + 1) % max(1, poll_iters) == 0 or it + 1 == self.iteration_num:
                    state = net.read_state()
                    if state["stop"]:
                        break
        best_loss, iters = state["best_loss"], state["iterations"]
        if not np.isfinite(best_loss):
            raise RuntimeError("NSFP: the loss was never finite")        # the reference dies on model_res['flow'] here
        self.last_info = {"loss": best_loss, "iterations": iters}
        return {"loss": best_loss, "flow": best_flow, "iterations": iters}

    def forward(self, batch: Dict, init_state_dicts: Optional[List[Dict]] = None) -> Dict[str, List[torch.Tensor]]:
        """nsfp.py:142-186."""
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
            res = self.optimize(sel0 + pf, sel1, init_state_dict=init_state_dicts[b] if init_state_dicts else None)
            final = torch.zeros_like(pc0)
            final[rm0] = res["flow"]
            flows.append(final)
            pose_flows.append(pf)
        return {"flow": flows, "pose_flow": pose_flows}

    __call__ = forward
