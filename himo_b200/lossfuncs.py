"""SeFlow / SeFlow++ self-supervised losses: OSF/src/lossfuncs/selfsupervise.py:25-190.

Same dictionary in (`pc0`, `pc1`, [`pch1`], `est_flow`, `pc0_labels`, `pc1_labels`, [`pch1_labels`]), same four terms
out (`chamfer_dis`, `dynamic_chamfer_dis`, `static_flow_loss`, `cluster_based_pc0pc1`).

Built differently from the reference in two places:
  * the Chamfer terms go through `himo_b200.chamfer3d.nnChamferDis`, whose truncated calls prune the search by the
    truncation radius (2 m) instead of masking a full search afterwards;
  * the per-cluster term (SeFlow Eq. 8-11) is segmented: the reference walks `torch.unique(pc0_label)` in a Python loop
    with an argsort, a nonzero and a cat per cluster (hundreds of tiny launches and host syncs per frame); here one
    `scatter_reduce(amax)` finds, for every cluster at once, the member with the largest nearest-neighbour distance
    among those whose neighbour in pc1 is dynamic, and one gather broadcasts its displacement back.  Ties on the
    distance are broken by the lowest point index (the reference leaves them to an unstable argsort).
`chamfer` is injectable so that the host logic can be tested on CPU against the reference's own functions
(tests/test_lossfuncs.py); the default is the CUDA module.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .chamfer3d import nnChamferDis

TRUNCATED_DIST = 4      # squared metres: 2 m per 0.1 s (selfsupervise.py:21-23)
_MIN_DYNAMIC = 256      # selfsupervise.py:44-45 (THREADS_PER_BLOCK of the reference kernel)

_default_chamfer: Optional[nnChamferDis] = None


def _chamfer(chamfer):
    global _default_chamfer
    if chamfer is not None:
        return chamfer
    if _default_chamfer is None:
        _default_chamfer = nnChamferDis()
    return _default_chamfer


def _truncated_mean(d: torch.Tensor) -> torch.Tensor:
    return torch.mean(d[d <= TRUNCATED_DIST])


def _static_and_cluster_terms(pc0, pc1, est_flow, pc0_label, pc1_label, raw_dist0, raw_dist1, raw_idx0, have_dynamic_cluster):
    """selfsupervise.py:61-100 / 146-181 without the per-label loop."""
    zero = torch.tensor(0.0, device=est_flow.device)
    static_mask = pc0_label == 0
    static_loss = zero + torch.linalg.vector_norm(est_flow[static_mask], dim=-1).mean() if bool(static_mask.any()) else zero

    moved_loss = zero
    if not have_dynamic_cluster:
        return static_loss, moved_loss
    member = pc0_label > 1                                              # label 1: dynamic but unclustered
    nn_idx = raw_idx0.long()
    cand = member & (pc1_label[nn_idx.clamp(min=0)] > 0) & (nn_idx >= 0)  # neighbour in pc1 is dynamic
    n_contrib = 0
    if bool(cand.any()):
        labels, inv = torch.unique(pc0_label, return_inverse=True)
        k = labels.numel()
        neg = torch.full((k,), -float("inf"), device=est_flow.device, dtype=raw_dist0.dtype)
        seg_max = neg.scatter_reduce(0, inv[cand], raw_dist0.detach()[cand], reduce="amax", include_self=True)
        point = torch.arange(pc0.shape[0], device=est_flow.device)
        is_rep = cand & (raw_dist0.detach() == seg_max[inv])
        big = torch.full((k,), pc0.shape[0], device=est_flow.device, dtype=torch.long)
        rep = big.scatter_reduce(0, inv[is_rep], point[is_rep], reduce="amin", include_self=True)
        has_rep = rep < pc0.shape[0]
        rep = rep.clamp(max=pc0.shape[0] - 1)
        max_flow = pc1[nn_idx[rep]] - pc0[rep]                          # Eq. 9, one row per cluster
        contrib = member & has_rep[inv]
        n_contrib = int(contrib.sum())
        if n_contrib:
            moved_loss = torch.linalg.vector_norm(est_flow[contrib] - max_flow[inv[contrib]], dim=-1).mean()
    if n_contrib == 0:
        moved_loss = _truncated_mean(raw_dist0) + _truncated_mean(raw_dist1)     # selfsupervise.py:99-100
    return static_loss, moved_loss


def seflowLoss(res_dict: Dict[str, torch.Tensor], timer=None, chamfer=None) -> Dict[str, torch.Tensor]:
    """selfsupervise.py:113-190."""
    ch = _chamfer(chamfer)
    pc0_label, pc1_label = res_dict["pc0_labels"], res_dict["pc1_labels"]
    pc0, pc1, est_flow = res_dict["pc0"], res_dict["pc1"], res_dict["est_flow"]
    pseudo_pc1from0 = pc0 + est_flow
    dyn0 = pc0_label > 0
    pc1_dynamic = pc1[pc1_label > 0]
    have_dynamic_cluster = int(dyn0.sum()) > _MIN_DYNAMIC and pc1_dynamic.shape[0] > _MIN_DYNAMIC

    chamfer_dis = ch(pseudo_pc1from0, pc1, truncate_dist=TRUNCATED_DIST)
    raw_dist0, raw_dist1, raw_idx0, _ = ch.disid_res(pc0, pc1)
    dynamic_chamfer_dis = torch.tensor(0.0, device=est_flow.device)
    if have_dynamic_cluster:
        dynamic_chamfer_dis = dynamic_chamfer_dis + ch(pseudo_pc1from0[dyn0], pc1_dynamic, truncate_dist=TRUNCATED_DIST)
    static_loss, moved_loss = _static_and_cluster_terms(pc0, pc1, est_flow, pc0_label, pc1_label, raw_dist0, raw_dist1,
                                                        raw_idx0, have_dynamic_cluster)
    return {"chamfer_dis": chamfer_dis, "dynamic_chamfer_dis": dynamic_chamfer_dis,
            "static_flow_loss": static_loss, "cluster_based_pc0pc1": moved_loss}


def seflowppLoss(res_dict: Dict[str, torch.Tensor], timer=None, chamfer=None) -> Dict[str, torch.Tensor]:
    """selfsupervise.py:25-111: the three-frame form; the flow is applied backwards to reach the history frame."""
    ch = _chamfer(chamfer)
    pch1_label, pc0_label, pc1_label = res_dict["pch1_labels"], res_dict["pc0_labels"], res_dict["pc1_labels"]
    pch1, pc0, pc1, est_flow = res_dict["pch1"], res_dict["pc0"], res_dict["pc1"], res_dict["est_flow"]
    pseudo_pc1from0 = pc0 + est_flow
    pseudo_pch1from0 = pc0 - est_flow
    dyn0 = pc0_label > 0
    pc1_dynamic = pc1[pc1_label > 0]
    have_dynamic_cluster = int(dyn0.sum()) > _MIN_DYNAMIC and pc1_dynamic.shape[0] > _MIN_DYNAMIC

    chamfer_dis = ch(pseudo_pc1from0, pc1, truncate_dist=TRUNCATED_DIST) + ch(pseudo_pch1from0, pch1, truncate_dist=TRUNCATED_DIST)
    dynamic_chamfer_dis = torch.tensor(0.0, device=est_flow.device)
    if have_dynamic_cluster:
        dynamic_chamfer_dis = dynamic_chamfer_dis + ch(pseudo_pc1from0[dyn0], pc1_dynamic, truncate_dist=TRUNCATED_DIST)
        pch1_dynamic = pch1[pch1_label > 0]
        if pch1_dynamic.shape[0] > _MIN_DYNAMIC:
            dynamic_chamfer_dis = dynamic_chamfer_dis + ch(pseudo_pch1from0[dyn0], pch1_dynamic, truncate_dist=TRUNCATED_DIST)
    raw_dist0, raw_dist1, raw_idx0, _ = ch.disid_res(pc0, pc1)
    static_loss, moved_loss = _static_and_cluster_terms(pc0, pc1, est_flow, pc0_label, pc1_label, raw_dist0, raw_dist1,
                                                        raw_idx0, have_dynamic_cluster)
    return {"chamfer_dis": chamfer_dis / 2.0, "dynamic_chamfer_dis": dynamic_chamfer_dis / 2.0,
            "static_flow_loss": static_loss, "cluster_based_pc0pc1": moved_loss}
