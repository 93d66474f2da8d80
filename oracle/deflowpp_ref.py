"""oracle/deflowpp_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch-CPU fp32 restatement of the SeFlow++ network forward, `DeFlowPP.forward`
(OSF/src/models/deflow.py:115-158), written functionally over a plain state_dict.  It is the
parity checker for the fused CUDA path in himo_b200/ and the CPU baseline ("port") of bench.py.
Pinned against the reference's own classes run in the build container: tests/golden/*.npz
(generator: tests/golden/make_golden.py) and tests/test_oracle_pinned.py.

Each stage cites the reference lines it follows.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import leaf

VOXEL_SIZE = [0.2, 0.2, 6.0]
POINT_CLOUD_RANGE = [-51.2, -51.2, -3.0, 51.2, 51.2, 3.0]
GRID = 512


def pose0to1(pose0: torch.Tensor, pose1: torch.Tensor) -> torch.Tensor:
    """cal_pose0to1 (OSF/src/models/basic/__init__.py:20-30): inv(pose1) @ pose0 with the rigid
    inverse formed in float64, result cast to float32."""
    inv = torch.eye(4, dtype=torch.float64)
    inv[:3, :3] = pose1[:3, :3].T
    inv[:3, 3] = (pose1[:3, :3].T * -pose1[:3, 3]).sum(axis=1)
    return (inv @ pose0.type(torch.float64)).type(torch.float32)


def rigid_apply(pc: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """wrap_batch_pcs warp (basic/__init__.py:50,57): pc @ R^T + t in fp32."""
    return pc @ T[:3, :3].T + T[:3, 3]


def voxelize_frame(points: torch.Tensor) -> Dict[str, torch.Tensor]:
    """DynamicVoxelizer.forward for one batch item (OSF/src/models/basic/encoder.py:567-600)."""
    idx = torch.arange(points.shape[0])
    keep = ~torch.isnan(points).any(dim=1)                                  # :576-578
    pts = points[keep]
    idx = idx[keep]
    coors = torch.from_numpy(leaf.dynamic_voxelize(pts.numpy(), VOXEL_SIZE, POINT_CLOUD_RANGE))
    inside = (coors != -1).all(dim=1)                                       # :581
    pts, coors, idx = pts[inside], coors[inside], idx[inside]
    rng = torch.tensor(POINT_CLOUD_RANGE, dtype=pts.dtype)
    vs = torch.tensor(VOXEL_SIZE, dtype=pts.dtype)
    centers = coors[:, [2, 1, 0]] * vs + rng[:3] + vs / 2                   # :506-523
    return {"points": pts, "voxel_coords": coors, "point_idxes": idx,
            "point_offsets": pts[:, :3] - centers}


def pillar_features(sd: Dict[str, torch.Tensor], pts: torch.Tensor, coors: torch.Tensor,
                    accum: str = "f32_seq"):
    """DynamicPillarFeatureNet.forward (encoder.py:430-475) with one PFN layer
    Linear(9,32,no bias) + BatchNorm1d(eps=1e-3, eval) + ReLU (:362-371), mean scatter."""
    p = "embedder.feature_net.pfn_layers.0"
    np_pts, np_co = pts.numpy(), coors.numpy()
    vmean, vco, p2v, cnt = leaf.dynamic_point_to_voxel(np_pts, np_co, "mean", accum)   # :442
    if pts.shape[0]:
        points_mean = torch.from_numpy(vmean)[torch.from_numpy(p2v).long()]            # :380-428
    else:
        points_mean = torch.zeros((0, 3))
    f_cluster = pts[:, :3] - points_mean[:, :3]                                        # :446
    vx, vy, vz = VOXEL_SIZE
    x_off = vx / 2 + POINT_CLOUD_RANGE[0]                                             # :257-259
    y_off = vy / 2 + POINT_CLOUD_RANGE[1]
    z_off = vz / 2 + POINT_CLOUD_RANGE[2]
    f_center = pts.new_zeros((pts.shape[0], 3))                                        # :451-458
    f_center[:, 0] = pts[:, 0] - (coors[:, 2].type_as(pts) * vx + x_off)
    f_center[:, 1] = pts[:, 1] - (coors[:, 1].type_as(pts) * vy + y_off)
    f_center[:, 2] = pts[:, 2] - (coors[:, 0].type_as(pts) * vz + z_off)
    feats = torch.cat([pts, f_cluster, f_center], dim=-1)                              # :465
    y = F.linear(feats, sd[p + ".0.weight"])
    y = F.batch_norm(y, sd[p + ".1.running_mean"], sd[p + ".1.running_var"], sd[p + ".1.weight"],
                     sd[p + ".1.bias"], training=False, eps=1e-3)
    y = F.relu(y)
    vfeat, vco2, _, _ = leaf.dynamic_point_to_voxel(y.numpy(), np_co, "mean", accum)   # :468
    return (torch.from_numpy(vfeat), torch.from_numpy(vco2), y, torch.from_numpy(vmean),
            torch.from_numpy(p2v), torch.from_numpy(cnt))


def pseudo_image(voxel_feats: torch.Tensor, voxel_coors: torch.Tensor) -> torch.Tensor:
    """PointPillarsScatter.forward_single (encoder.py:126-147) -> [1,32,512,512]."""
    canvas = torch.zeros(voxel_feats.shape[1], GRID * GRID, dtype=voxel_feats.dtype)
    if voxel_feats.shape[0]:
        flat = (voxel_coors[:, 1] * GRID + voxel_coors[:, 2]).long()
        canvas[:, flat] = voxel_feats.t()
    return canvas.view(1, voxel_feats.shape[1], GRID, GRID)


def embed_frame(sd, points: torch.Tensor, accum: str = "f32_seq"):
    """DynamicEmbedder.forward for one frame (encoder.py:618-631)."""
    info = voxelize_frame(points)
    vfeat, vco, pfeat, vmean, p2v, cnt = pillar_features(sd, info["points"], info["voxel_coords"], accum)
    info.update(voxel_feats=vfeat, voxel_coors=vco, point_feats=pfeat, voxel_mean=vmean,
                point2voxel=p2v, voxel_count=cnt)
    return pseudo_image(vfeat, vco), info


def _conv_bn_gelu(sd, name: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """ConvWithNorms.forward (basic/__init__.py:76-94): Conv2d(3x3, pad 1) + BN2d(eval) + exact GELU."""
    q = "backbone." + name
    y = F.conv2d(x, sd[q + ".conv.weight"], sd[q + ".conv.bias"], stride=stride, padding=1)
    y = F.batch_norm(y, sd[q + ".batchnorm.running_mean"], sd[q + ".batchnorm.running_var"],
                     sd[q + ".batchnorm.weight"], sd[q + ".batchnorm.bias"], training=False, eps=1e-5)
    return F.gelu(y)


def _encoder(sd, x: torch.Tensor):
    """encoder_step_1..3 (unet.py:110-127): returns the three scales F (64@256), L (128@128), R (256@64)."""
    outs = []
    for step, n in (("encoder_step_1", 4), ("encoder_step_2", 6), ("encoder_step_3", 6)):
        for i in range(n):
            x = _conv_bn_gelu(sd, f"{step}.{i}", x, 2 if i == 0 else 1)
        outs.append(x)
    return outs


def _upsample_skip(sd, name: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """UpsampleSkip.forward (unet.py:31-35)."""
    q = "backbone." + name
    u2 = F.conv2d(a, sd[q + ".u1_u2.0.weight"], sd[q + ".u1_u2.0.bias"])
    u2 = F.interpolate(u2, scale_factor=2, mode="bilinear", align_corners=False)
    u3 = F.conv2d(b, sd[q + ".u3.weight"], sd[q + ".u3.bias"])
    x = torch.cat([u2, u3], dim=1)
    x = F.conv2d(x, sd[q + ".u4_u5.0.weight"], sd[q + ".u4_u5.0.bias"], padding=1)
    return F.conv2d(x, sd[q + ".u4_u5.1.weight"], sd[q + ".u4_u5.1.bias"], padding=1)


def unet_three_frame(sd, img_h1, img_0, img_1, return_stages: bool = False):
    """UNetThreeFrame.forward (unet.py:131-166)."""
    enc = [_encoder(sd, im) for im in (img_h1, img_0, img_1)]
    Fs = torch.cat([e[0] for e in enc], dim=1)
    Ls = torch.cat([e[1] for e in enc], dim=1)
    Rs = torch.cat([e[2] for e in enc], dim=1)
    Bs = torch.cat([img_h1, img_0, img_1], dim=1)
    S = _upsample_skip(sd, "decoder_step1", Rs, Ls)
    T = _upsample_skip(sd, "decoder_step2", S, Fs)
    U = _upsample_skip(sd, "decoder_step3", T, Bs)
    V = F.conv2d(U, sd["backbone.decoder_step4.weight"], sd["backbone.decoder_step4.bias"], padding=1)
    if return_stages:
        return V, {"F": Fs, "L": Ls, "R": Rs, "S": S, "T": T, "U": U}
    return V


def gru_decoder(sd, before: torch.Tensor, after: torch.Tensor, offsets: torch.Tensor,
                coords: torch.Tensor, num_iters: int = 2) -> torch.Tensor:
    """ConvGRUDecoder.forward_single (decoder.py:210-237) with ConvGRU (:177-193).
    before/after: [96,512,512]; offsets [N,3]; coords [N,3] (z,y,x)."""
    co = coords.long()
    a = after[:, co[:, 1], co[:, 2]].T
    b = before[:, co[:, 1], co[:, 2]].T
    h = torch.cat([b, a], dim=1)                                                   # [N,192]
    x = F.linear(offsets, sd["head.offset_encoder.weight"], sd["head.offset_encoder.bias"])
    wz, bz = sd["head.gru.convz.weight"][:, :, 0], sd["head.gru.convz.bias"]
    wr, br = sd["head.gru.convr.weight"][:, :, 0], sd["head.gru.convr.bias"]
    wq, bq = sd["head.gru.convq.weight"][:, :, 0], sd["head.gru.convq.bias"]
    for _ in range(num_iters):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(F.linear(hx, wz, bz))
        r = torch.sigmoid(F.linear(hx, wr, br))
        q = torch.tanh(F.linear(torch.cat([r * h, x], dim=1), wq, bq))
        h = (1 - z) * h + z * q
    y = F.gelu(F.linear(torch.cat([h, x], dim=1), sd["head.decoder.0.weight"], sd["head.decoder.0.bias"]))
    return F.linear(y, sd["head.decoder.2.weight"], sd["head.decoder.2.bias"])


@torch.no_grad()
def deflowpp_forward(sd: Dict[str, torch.Tensor], pch1, pc0, pc1, poseh1, pose0, pose1,
                     accum: str = "f32_seq", num_iters: int = 2, keep: bool = False) -> Dict:
    """DeFlowPP.forward for ONE frame triple (batch size 1), ground points already removed
    (OSF/src/trainer.py:290-297).  Inputs: [N,3] float32 clouds and [4,4] poses (numpy or torch).
    Returns flow [Nvalid,3], pose_flow [N0,3], pc0_valid_point_idxes [Nvalid] (+ stage tensors)."""
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    pch1, pc0, pc1 = t(pch1).float(), t(pc0).float(), t(pc1).float()
    poseh1, pose0, pose1 = t(poseh1), t(pose0), t(pose1)
    T01 = pose0to1(pose0, pose1)                                   # wrap_batch_pcs, basic/__init__.py:32-74
    Th1 = pose0to1(poseh1, pose1)
    pc0_w = rigid_apply(pc0, T01)
    pose_flow = pc0_w - pc0
    pch1_w = rigid_apply(pch1, Th1)
    img_h1, info_h1 = embed_frame(sd, pch1_w, accum)               # deflow.py:129-131
    img_0, info_0 = embed_frame(sd, pc0_w, accum)
    img_1, info_1 = embed_frame(sd, pc1, accum)
    after = unet_three_frame(sd, img_h1, img_0, img_1)             # deflow.py:135-136
    before = torch.cat([img_h1, img_0, img_1], dim=1)              # deflow.py:140-141
    flow = gru_decoder(sd, before[0], after[0], info_0["point_offsets"], info_0["voxel_coords"],
                       num_iters)
    out = {"flow": flow, "pose_flow": pose_flow, "pc0_valid_point_idxes": info_0["point_idxes"],
           "pc1_valid_point_idxes": info_1["point_idxes"],
           "pch1_valid_point_idxes": info_h1["point_idxes"]}
    if keep:
        out.update(before=before[0], after=after[0], info_h1=info_h1, info_0=info_0, info_1=info_1,
                   pc0_w=pc0_w, pch1_w=pch1_w)
    return out


def final_flow(pc0_all: torch.Tensor, gm0: torch.Tensor, pose0, pose1, res: Dict) -> torch.Tensor:
    """ModelWrapper.test_step packing (OSF/src/trainer.py:318-335): total flow for ALL points of
    pc0 (ground included): pose_flow everywhere, + network flow on the valid non-ground points."""
    T01 = pose0to1(torch.as_tensor(pose0), torch.as_tensor(pose1))
    pose_flow = rigid_apply(pc0_all, T01) - pc0_all
    out = pose_flow.clone()
    pred = pose_flow[~gm0].clone()
    v = res["pc0_valid_point_idxes"]
    pred[v] = pose_flow[~gm0][v] + res["flow"]
    out[~gm0] = pred
    return out
