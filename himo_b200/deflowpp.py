"""SeFlow++ network (`DeFlowPP`) on the B200 engine: host-side mirror of the reference model class.

Same constructor arguments, `load_state_dict` keys, `forward(batch) -> dict` contract and `.timer`
attribute as `src.models.DeFlowPP` (OSF/src/models/deflow.py:90-158), so the reference drivers
(`ModelWrapper.test_step`, OSF/src/trainer.py:290-343) can hold it unchanged.  All compute is one
C-ABI call, `himo_deflowpp_forward` (csrc/deflowpp.cu); there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_size_t, c_void_p
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, conv, weights as W

HIMO_MAX_FRAMES = 4


class _Weights(ctypes.Structure):
    _fields_ = [
        ("planes", c_int),
        ("pfn_w", c_void_p), ("pfn_b", c_void_p),
        ("enc_w", c_void_p * 16), ("enc_b", c_void_p * 16),
        ("dec_w", (c_void_p * 4) * 3), ("dec_b", (c_void_p * 4) * 3),
        ("dec4_w", c_void_p), ("dec4_b", c_void_p),
        ("off_w", c_void_p), ("off_b", c_void_p),
        ("gru_zr_w", c_void_p), ("gru_zr_b", c_void_p),
        ("gru_q_w", c_void_p), ("gru_q_b", c_void_p),
        ("dec0_w", c_void_p), ("dec0_b", c_void_p),
        ("dec2_w", c_void_p), ("dec2_b", c_void_p),
        ("enc_s", c_float * 16), ("dec_s", (c_float * 4) * 3), ("dec4_s", c_float),
        ("gru_zr_s", c_float), ("gru_q_s", c_float), ("dec0_s", c_float),
        ("dec_bb", c_void_p * 3),
    ]


class _IO(ctypes.Structure):
    _fields_ = [
        ("pch1", c_void_p), ("n_h1", c_int),
        ("pc0", c_void_p), ("n0", c_int),
        ("pc1", c_void_p), ("n1", c_int),
        ("T_h1", c_float * 12), ("T_0", c_float * 12),
        ("n_max", c_int), ("num_iters", c_int),
        ("flow_all", c_void_p), ("valid_idx", c_void_p), ("flow_valid", c_void_p), ("n_valid", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("stage_events", c_void_p * 4),
    ]


class _View(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("canvas", "Fstar", "Lstar", "Rstar", "S", "T", "U", "V",
                                        "embed_ws", "h32")]


_lib.register("himo_deflowpp_workspace_bytes", c_size_t, [c_int, c_int])
_lib.register("himo_deflowpp_forward", c_int, [ctypes.POINTER(_Weights), ctypes.POINTER(_IO), c_void_p])
_lib.register("himo_deflowpp_views", c_int, [c_int, c_int, c_void_p, ctypes.POINTER(_View)])
_lib.register("himo_rigid_flow", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p])
_lib.register("himo_final_flow", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p])
_lib.register("himo_launch_count", ctypes.c_ulonglong, [])
_lib.register("himo_deflowpp_set_fused_decoder", c_int, [c_int])


def compose_u3_u4(w3: torch.Tensor, b3: torch.Tensor, w4: torch.Tensor, b4: torch.Tensor):
    """UpsampleSkip.forward (OSF/src/models/basic/unet.py:31-35): u4(cat([up(u1(a)), u3(b)])) with u3 a 1x1 and u4 a
    zero-padded 3x3 convolution and nothing in between.  Returns (W4', b4', border) in float32, composed in float64:
      W4'[:, :L] = W4[:, :L];  W4'[:, L:, ky, kx] = W4[:, L:, ky, kx] @ W3        (L = latent channels)
      b4' = b4 + sum over the 9 taps of W4[:, L:, tap] @ b3                        (interior pixels)
      border[yc][xc] = b4 + sum over the taps that stay inside the image of W4[:, L:, tap] @ b3, yc / xc = 0 first
      row / column, 1 interior, 2 last: a tap that lands in u4's zero padding sees 0, not u3's bias."""
    L = w3.shape[0]
    w3m, b3d = w3.double().reshape(L, -1), b3.double()
    w4d = w4.double()
    out = w4d.clone()
    out[:, L:] = torch.einsum("omyx,mi->oiyx", w4d[:, L:], w3m)
    tap_bias = torch.einsum("omyx,m->oyx", w4d[:, L:], b3d)                 # [Cout, 3, 3]
    border = torch.empty(3, 3, w4.shape[0], dtype=torch.float64)
    for yc in range(3):
        for xc in range(3):
            ky = [k for k in range(3) if not (yc == 0 and k == 0) and not (yc == 2 and k == 2)]
            kx = [k for k in range(3) if not (xc == 0 and k == 0) and not (xc == 2 and k == 2)]
            border[yc, xc] = b4.double() + tap_bias[:, ky][:, :, kx].sum((1, 2))
    return out.float(), border[1, 1].float().contiguous(), border.float().contiguous()


def cal_pose0to1(pose0: torch.Tensor, pose1: torch.Tensor) -> torch.Tensor:
    """inv(pose1) @ pose0 with the rigid inverse assembled in float64 and the result cast to
    float32 -- the arithmetic of cal_pose0to1 (OSF/src/models/basic/__init__.py:20-30), on the host
    (4x4 plumbing, not hot-path compute)."""
    pose0 = torch.as_tensor(pose0).detach().cpu()
    pose1 = torch.as_tensor(pose1).detach().cpu()
    inv = torch.eye(4, dtype=torch.float64)
    inv[:3, :3] = pose1[:3, :3].T
    inv[:3, 3] = (pose1[:3, :3].T * -pose1[:3, 3]).sum(axis=1)
    return (inv @ pose0.type(torch.float64)).type(torch.float32)


class _Timer:
    """dztimer.Timing work-alike: the reference drivers call model.timer[i].start()/stop()/print()."""

    def __getitem__(self, _):
        return self

    def start(self, *_a, **_k):
        pass

    def stop(self, *_a, **_k):
        pass

    def print(self, *_a, **_k):
        pass


class DeFlowPP:
    """Drop-in for `src.models.DeFlowPP` (inference only).  precision: "fp32" = split-bf16 planes
    (three tensor-core products per k-step, fp32-class results; the parity mode), "bf16" = single
    plane."""

    def __init__(self, voxel_size=(0.2, 0.2, 6), point_cloud_range=(-51.2, -51.2, -3, 51.2, 51.2, 3),
                 grid_feature_size=(512, 512), decoder_option="gru", num_iters=2, num_frames=3,
                 precision: str = "fp32", device="cuda", max_points: int = 131072, compose_skip: Optional[bool] = None):
        if list(voxel_size) != [0.2, 0.2, 6] or list(point_cloud_range) != [-51.2, -51.2, -3, 51.2, 51.2, 3] \
                or list(grid_feature_size)[:2] != [512, 512]:
            raise NotImplementedError("himo_b200.DeFlowPP is built for the SeFlow++ grid (conf/model/deflowpp.yaml)")
        if decoder_option != "gru" or num_frames != 3:
            raise NotImplementedError("DeFlowPP only supports the gru decoder with num_frames = 3")
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        # UpsampleSkip applies u3 (1x1 on the skip) and u4 (3x3 on the concatenation) back to back with no activation in
        # between (unet.py:31-35): composed on the host into one 3x3 convolution (see load_state_dict)
        self.compose_skip = (os.environ.get("HIMO_COMPOSE_SKIP", "1") != "0") if compose_skip is None else bool(compose_skip)
        self.num_iters = int(num_iters)
        self.num_frames = 3
        self.planes = 2 if precision == "fp32" else 1
        self.precision = precision
        self.device = torch.device(device)
        self.timer = _Timer()
        self._tensors: List[torch.Tensor] = []      # keeps the packed device weights alive
        self._w: Optional[_Weights] = None
        self._ws: Optional[torch.Tensor] = None
        self._n_max = 0
        self._reserve = int(max_points)
        _lib.lib()                                   # fail now if the CUDA library is missing

    # ------------------------------------------------------------------ weights
    def _dev(self, t: torch.Tensor):
        t = t.contiguous().to(self.device)
        self._tensors.append(t)
        return t.data_ptr()

    def _pack(self, w4: torch.Tensor):
        """conv weight -> (device pointer of the packed planes, accumulator scale)."""
        s = conv.weight_prescale(w4, self.planes)
        return self._dev(conv.pack_conv_weight(w4.float(), self.planes, s)), 1.0 / s

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        P = self.planes
        self._tensors = []
        w = _Weights()
        w.planes = P
        p = "embedder.feature_net.pfn_layers.0"
        pw, pb = W.fold_bn(sd[p + ".0.weight"], None, sd[p + ".1.weight"], sd[p + ".1.bias"],
                           sd[p + ".1.running_mean"], sd[p + ".1.running_var"], 1e-3)
        w.pfn_w, w.pfn_b = self._dev(pw), self._dev(pb)
        for i, (name, cin, cout, stride) in enumerate(W.ENCODER_LAYERS):
            q = "backbone." + name
            cw, cb = W.fold_bn(sd[q + ".conv.weight"], sd[q + ".conv.bias"], sd[q + ".batchnorm.weight"],
                               sd[q + ".batchnorm.bias"], sd[q + ".batchnorm.running_mean"],
                               sd[q + ".batchnorm.running_var"], 1e-5)
            w.enc_w[i], w.enc_s[i] = self._pack(cw)
            w.enc_b[i] = self._dev(cb)
        for bi, (name, _, _, _) in enumerate(W.DECODER_BLOCKS):
            q = "backbone." + name
            for j, sub in enumerate((".u1_u2.0", ".u3", ".u4_u5.0", ".u4_u5.1")):
                w.dec_w[bi][j], w.dec_s[bi][j] = self._pack(sd[q + sub + ".weight"])
                w.dec_b[bi][j] = self._dev(sd[q + sub + ".bias"].float())
            if self.compose_skip:
                w4, b4, bb = compose_u3_u4(sd[q + ".u3.weight"], sd[q + ".u3.bias"], sd[q + ".u4_u5.0.weight"],
                                           sd[q + ".u4_u5.0.bias"])
                w.dec_w[bi][2], w.dec_s[bi][2] = self._pack(w4)
                w.dec_b[bi][2] = self._dev(b4)
                w.dec_bb[bi] = self._dev(bb)
        w.dec4_w, w.dec4_s = self._pack(sd["backbone.decoder_step4.weight"])
        w.dec4_b = self._dev(sd["backbone.decoder_step4.bias"].float())
        w.off_w = self._dev(sd["head.offset_encoder.weight"].float())
        w.off_b = self._dev(sd["head.offset_encoder.bias"].float())
        zr = torch.cat([sd["head.gru.convz.weight"], sd["head.gru.convr.weight"]], 0)     # [384,288,1]
        w.gru_zr_w, w.gru_zr_s = self._pack(zr.unsqueeze(-1))
        w.gru_zr_b = self._dev(torch.cat([sd["head.gru.convz.bias"], sd["head.gru.convr.bias"]]).float())
        w.gru_q_w, w.gru_q_s = self._pack(sd["head.gru.convq.weight"].unsqueeze(-1))
        w.gru_q_b = self._dev(sd["head.gru.convq.bias"].float())
        d0 = torch.zeros(64, 288)
        d0[:48] = sd["head.decoder.0.weight"]
        b0 = torch.zeros(64)
        b0[:48] = sd["head.decoder.0.bias"]
        w.dec0_w, w.dec0_s = self._pack(d0.view(64, 288, 1, 1))
        w.dec0_b = self._dev(b0)
        w.dec2_w = self._dev(sd["head.decoder.2.weight"].float())
        w.dec2_b = self._dev(sd["head.decoder.2.bias"].float())
        self._w = w
        return self

    def replica(self) -> "DeFlowPP":
        """A second network over the SAME packed device weights with its own workspace, so that two frame triples can be
        in flight on two CUDA streams (engine.SeFlowPPEngine: the second stream's kernels fill the tail waves of the first)."""
        if self._w is None:
            raise RuntimeError("load_state_dict first")
        r = DeFlowPP(precision=self.precision, device=self.device, max_points=self._reserve, num_iters=self.num_iters,
                     compose_skip=self.compose_skip)
        r._w, r._tensors = self._w, self._tensors
        return r

    def load_from_checkpoint(self, ckpt_path: str):
        """BaseModel.load_from_checkpoint (OSF/src/models/basic/__init__.py:10-16)."""
        return self.load_state_dict(W.load_deflowpp_checkpoint(ckpt_path))

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("construct himo_b200.DeFlowPP with device=... instead of moving it")
        return self

    # ------------------------------------------------------------------ compute
    def _workspace(self, n_max: int) -> torch.Tensor:
        n_max = max(n_max, self._reserve)
        if self._ws is None or n_max > self._n_max:
            nbytes = _lib.lib().himo_deflowpp_workspace_bytes(n_max, self.planes)
            if nbytes == 0:
                raise RuntimeError("himo_deflowpp_workspace_bytes failed")
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._n_max = n_max
        return self._ws

    def forward_triple(self, pch1: torch.Tensor, pc0: torch.Tensor, pc1: torch.Tensor,
                       T_h1: torch.Tensor, T_0: torch.Tensor, compact: bool = True,
                       stage_events=None, flow_all_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One frame triple.  pch1/pc0/pc1: [N,3] f32 CUDA (ground-free, sensor frames);
        T_h1/T_0: [4,4] f32 host transforms into the pc1 frame (cal_pose0to1).
        Returns flow_all [N0,3] (0 where the point was dropped) and, if compact, the reference's
        compact outputs flow_valid / valid_idx / n_valid (device scalar).  `flow_all_out`: caller-owned [>=N0,3] f32
        CUDA buffer to write flow_all into (the streaming engine passes a preallocated slot; nothing is allocated then)."""
        if self._w is None:
            raise RuntimeError("load_state_dict first")
        for name, t in (("pch1", pch1), ("pc0", pc0), ("pc1", pc1)):
            _lib.require_cuda(t, name)
            if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 3 or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous float32 [N,3] tensor")
        dev = pc0.device
        n_max = max(pch1.shape[0], pc0.shape[0], pc1.shape[0], 1)
        with _lib.on_device(dev):
            ws = self._workspace(n_max)
            io = _IO()
            io.pch1, io.n_h1 = pch1.data_ptr(), pch1.shape[0]
            io.pc0, io.n0 = pc0.data_ptr(), pc0.shape[0]
            io.pc1, io.n1 = pc1.data_ptr(), pc1.shape[0]
            th = T_h1[:3, :4].contiguous().float().flatten().tolist()
            t0 = T_0[:3, :4].contiguous().float().flatten().tolist()
            for k in range(12):
                io.T_h1[k] = th[k]
                io.T_0[k] = t0[k]
            io.n_max = self._n_max
            io.num_iters = self.num_iters
            n0 = pc0.shape[0]
            if flow_all_out is not None:
                if flow_all_out.dtype != torch.float32 or not flow_all_out.is_contiguous() or flow_all_out.shape[0] < n0:
                    raise RuntimeError("flow_all_out must be a contiguous float32 [>=N0,3] CUDA tensor")
                flow_all = flow_all_out[:n0]
            else:
                flow_all = torch.empty((n0, 3), dtype=torch.float32, device=dev)
            io.flow_all = flow_all.data_ptr()
            out = {"flow_all": flow_all}
            if compact:
                valid_idx = torch.empty((n0,), dtype=torch.int64, device=dev)
                flow_valid = torch.empty((n0, 3), dtype=torch.float32, device=dev)
                n_valid = torch.zeros((1,), dtype=torch.int32, device=dev)
                io.valid_idx, io.flow_valid, io.n_valid = valid_idx.data_ptr(), flow_valid.data_ptr(), n_valid.data_ptr()
                out.update(valid_idx=valid_idx, flow_valid=flow_valid, n_valid=n_valid)
            io.workspace = ws.data_ptr()
            io.workspace_bytes = ws.numel()
            if stage_events is not None:            # 4 torch.cuda.Event(enable_timing=True), already created
                for k in range(4):
                    io.stage_events[k] = stage_events[k].cuda_event
            st = _lib.lib().himo_deflowpp_forward(ctypes.byref(self._w), ctypes.byref(io), _lib.stream_ptr(dev))
        _lib.check(st, "himo_deflowpp_forward")
        return out

    def forward(self, batch: Dict) -> Dict[str, List[torch.Tensor]]:
        """`model(batch)` of the reference (deflow.py:115-158): batch = {pc0,pc1,pch1: [B,N,3],
        pose0,pose1,poseh1: list of [4,4]} -> {"flow": [...], "pose_flow": [...],
        "pc0_valid_point_idxes": [...], ...}."""
        B = len(batch["pose0"])
        res = {k: [] for k in ("flow", "pose_flow", "pc0_valid_point_idxes", "pc0_points_lst")}
        for b in range(B):
            if "ego_motion" in batch:
                T0 = torch.as_tensor(batch["ego_motion"][b]).detach().cpu().float()
            else:
                T0 = cal_pose0to1(batch["pose0"][b], batch["pose1"][b])
            Th = cal_pose0to1(batch["poseh1"][b], batch["pose1"][b])
            pc0 = batch["pc0"][b].contiguous()
            out = self.forward_triple(batch["pch1"][b].contiguous(), pc0, batch["pc1"][b].contiguous(), Th, T0)
            nv = int(out["n_valid"].item())       # the reference returns data-dependent shapes too
            res["flow"].append(out["flow_valid"][:nv])
            res["pc0_valid_point_idxes"].append(out["valid_idx"][:nv])
            res["pose_flow"].append(rigid_flow(pc0, T0))
            res["pc0_points_lst"].append(None)
        return res

    __call__ = forward

    def views(self) -> Dict[str, int]:
        v = _View()
        _lib.check(_lib.lib().himo_deflowpp_views(self._n_max, self.planes, _lib.ptr(self._ws), ctypes.byref(v)),
                   "himo_deflowpp_views")
        return {k: getattr(v, k) for k, _ in _View._fields_}


def rigid_flow(points: torch.Tensor, T: torch.Tensor, add_flow: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(p @ R^T + t) - p [+ add_flow] on the device (pose flow / final-flow assembly,
    OSF/src/trainer.py:320-335)."""
    _lib.require_cuda(points, "points")
    dev = points.device
    pts = points.contiguous()
    T12 = torch.as_tensor(T)[:3, :4].contiguous().float().flatten().to(dev)
    out = torch.empty_like(pts)
    with _lib.on_device(dev):
        st = _lib.lib().himo_rigid_flow(_lib.ptr(pts), pts.shape[0], _lib.ptr(T12),
                                        _lib.ptr(add_flow.contiguous()) if add_flow is not None else None,
                                        _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(st, "himo_rigid_flow")
    return out
