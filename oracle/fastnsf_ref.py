"""oracle/fastnsf_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch-CPU restatement of FastNSF (OSF/src/models/fastnsf.py:30-222) with the Neural_Prior MLP and the
EarlyStopping state machine of OSF/src/models/basic/nsfp_module.py, written functionally over a plain
state_dict so that the initial weights are an explicit input (the reference draws them from the global
torch RNG inside optimize(), fastnsf.py:110-115).

Pinned against the reference's own FastNSF / Neural_Prior / EarlyStopping classes run in the build
container (tests/golden/fastnsf_*.npz, tests/test_oracle_pinned.py).  The distance transform comes from
the third-party FastGeodis package, which is NOT vendored: its restatement (oracle/leaf_ops.c) is
PARITY UNPINNED, so every FastNSF parity statement is conditional on the distance volume D.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import leaf

POINT_CLOUD_RANGE = [-51.2, -51.2, -3.0, 51.2, 51.2, 3.0]


def range_mask(pc: torch.Tensor, rng=POINT_CLOUD_RANGE) -> torch.Tensor:
    """FastNSF.range_limit_ (fastnsf.py:171-178): inclusive box."""
    return ((pc[:, 0] >= rng[0]) & (pc[:, 0] <= rng[3]) & (pc[:, 1] >= rng[1]) & (pc[:, 1] <= rng[4]) &
            (pc[:, 2] >= rng[2]) & (pc[:, 2] <= rng[5]))


def dt_bounds(pc0: torch.Tensor, pc1: torch.Tensor, gf: float):
    """fastnsf.py:120-126: lo = floor(min*gf - 1)/gf, hi = ceil(max*gf + 1)/gf per axis (fp32)."""
    mn = torch.minimum(pc0.min(0)[0], pc1.min(0)[0])
    mx = torch.maximum(pc0.max(0)[0], pc1.max(0)[0])
    lo = torch.floor(mn * gf - 1) / gf
    hi = torch.ceil(mx * gf + 1) / gf
    return lo, hi


def dt_dims(lo: torch.Tensor, hi: torch.Tensor, gf: float) -> Tuple[int, int, int]:
    """DT.__init__ (fastnsf.py:34-36): samples per axis = ceil((hi-lo)*gf) + 2."""
    s = ((hi - lo) * gf).ceil().int() + 2
    return int(s[0]), int(s[1]), int(s[2])


def dt_build(pc1: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, gf: float) -> torch.Tensor:
    """DT.__init__ (fastnsf.py:30-57): occupancy at round((p - V[0])*gf) (half-to-even), then
    FastGeodis.generalised_geodesic3d(zeros, mask, [1/gf]*3, 1e10, 0.0, 1).  V_k[0] == lo_k."""
    H, W, D = dt_dims(lo, hi, gf)
    mask = torch.ones(H, W, D)
    ix = ((pc1[:, 0] - lo[0]) * gf).round().long()
    iy = ((pc1[:, 1] - lo[1]) * gf).round().long()
    iz = ((pc1[:, 2] - lo[2]) * gf).round().long()
    mask[ix, iy, iz] = 0.0
    d0 = (mask * np.float32(1e10)).numpy()
    out = leaf.geodesic3d_euclid(d0, [1.0 / gf] * 3, 1)
    return torch.from_numpy(out)


def dt_lookup(Dvol: torch.Tensor, lo: torch.Tensor, gf: float, Y: torch.Tensor) -> torch.Tensor:
    """DT.torch_bilinear_distance (fastnsf.py:59-80): clip, normalise to [-1,1], trilinear grid_sample
    with align_corners=True."""
    H, W, D = Dvol.shape
    sx = ((Y[:, 0:1] - lo[0]) * gf).clip(0, H - 1)
    sy = ((Y[:, 1:2] - lo[1]) * gf).clip(0, W - 1)
    sz = ((Y[:, 2:3] - lo[2]) * gf).clip(0, D - 1)
    s = torch.cat([sx, sy, sz], -1)
    s = 2 * s
    s = torch.stack([s[..., 0] / (H - 1), s[..., 1] / (W - 1), s[..., 2] / (D - 1)], -1)
    s = s - 1
    g = torch.cat([s[..., 2:3], s[..., 1:2], s[..., 0:1]], -1)
    return F.grid_sample(Dvol[None, None], g.view(1, -1, 1, 1, 3), mode="bilinear", align_corners=True).view(-1)


def mlp_forward(params: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Neural_Prior.forward (nsfp_module.py:41-47): 8 x (Linear + ReLU), final Linear."""
    n = len(params) // 2
    for l in range(n):
        x = F.linear(x, params[2 * l], params[2 * l + 1])
        if l < n - 1:
            x = F.relu(x)
    return x


def params_from_state_dict(sd: Dict[str, torch.Tensor], layer_size: int = 8) -> List[torch.Tensor]:
    out = []
    for i in range(layer_size):
        out += [sd[f"nn_layers.{2 * i}.0.weight"].clone(), sd[f"nn_layers.{2 * i}.0.bias"].clone()]
    out += [sd[f"nn_layers.{2 * layer_size}.weight"].clone(), sd[f"nn_layers.{2 * layer_size}.bias"].clone()]
    return out


def optimize(sd_init: Dict[str, torch.Tensor], pc0: torch.Tensor, pc1: torch.Tensor, itr_num: int = 5000,
             lr: float = 8e-3, min_delta: float = 5e-5, patience: int = 10, gf: float = 10.0,
             Dvol: Optional[torch.Tensor] = None, trace: bool = False) -> Dict:
    """FastNSF.optimize (fastnsf.py:105-169) + EarlyStopping.step (nsfp_module.py:65-82).
    pc0: ego-compensated, range-limited source cloud [N,3]; pc1: range-limited target cloud."""
    params = [p.requires_grad_(True) for p in params_from_state_dict(sd_init)]
    lo, hi = dt_bounds(pc0, pc1, gf)
    if Dvol is None:
        Dvol = dt_build(pc1, lo, hi, gf)
    opt = torch.optim.Adam(params, lr=lr, weight_decay=0)
    best_loss, best_flow = float("inf"), None
    es_best, es_bad = None, 0
    losses = []
    it = 0
    for it in range(itr_num):
        opt.zero_grad()
        flow = mlp_forward(params, pc0[None])[0]
        loss = dt_lookup(Dvol, lo, gf, pc0 + flow).mean()
        losses.append(float(loss))
        if float(loss) <= best_loss:                       # fastnsf.py:151-153 (snapshot before the update)
            best_loss, best_flow = float(loss), flow.detach().clone()
        # EarlyStopping.step
        if es_best is None:
            es_best, stop = loss.detach(), False
        elif torch.isnan(loss):
            stop = True
        else:
            if loss < es_best - min_delta:
                es_bad, es_best = 0, loss.detach()
            else:
                es_bad += 1
            stop = es_bad >= patience
        if stop:
            break
        loss.backward()
        opt.step()
    out = {"flow": best_flow, "loss": best_loss, "iterations": it + 1, "D": Dvol, "lo": lo}
    if trace:
        out["losses"] = losses
        out["params"] = [p.detach().clone() for p in params]
    return out


def fastnsf_forward(sd_init, pc0, pc1, pose0, pose1, **kw) -> Dict:
    """FastNSF.forward for one batch item (fastnsf.py:180-222): range crop, ego-motion warp of pc0,
    optimisation, scatter of the best flow back to all pc0 points."""
    from .deflowpp_ref import pose0to1
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    pc0, pc1 = t(pc0).float(), t(pc1).float()
    rm0, rm1 = range_mask(pc0), range_mask(pc1)
    T = pose0to1(t(pose0), t(pose1))
    sel0 = pc0[rm0]
    tr0 = sel0 @ T[:3, :3].T + T[:3, 3]
    res = optimize(sd_init, tr0.clone(), pc1[rm1].clone(), **kw)
    final = torch.zeros_like(pc0)
    final[rm0] = res["flow"]
    res.update(final_flow=final, pose_flow=tr0 - sel0, range_mask0=rm0)
    return res
