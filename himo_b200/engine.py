"""Public inference API of the SeFlow++ path: host frames in, per-point total flow out.

`SeFlowPPEngine.infer(frame)` is the work of `ModelWrapper.test_step` minus the .h5 write
(OSF/src/trainer.py:290-343): ground removal, DeFlowPP.forward, pose-flow + network-flow assembly
for ALL points of pc0.  Inputs are host arrays (as the reference's DataLoader delivers them);
the H2D copies, the network and the D2H copy of the result run on one CUDA stream per engine.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .deflowpp import DeFlowPP, cal_pose0to1


class _Pinned:
    """Grow-only pinned host staging buffer."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.buf: Optional[torch.Tensor] = None

    def get(self, n: int, cols: int = 3) -> torch.Tensor:
        need = max(n, 1) * cols
        if self.buf is None or self.buf.numel() < need:
            self.buf = torch.empty(int(need * 1.25) + 16, dtype=self.dtype, pin_memory=True)
        return self.buf[: n * cols].view(n, cols) if cols > 1 else self.buf[:n]


class SeFlowPPEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", precision: str = "fp32",
                 max_points: int = 131072):
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.net = DeFlowPP(precision=precision, device=self.device, max_points=max_points)
        self.net.load_state_dict(state_dict)
        self.stream = torch.cuda.Stream(self.device)
        self._pin = {k: _Pinned(torch.float32) for k in ("pch1", "pc0", "pc1", "pc0_all", "out")}
        self._pin_idx = _Pinned(torch.int32)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _strip(pc: np.ndarray, gm: Optional[np.ndarray]):
        pc = np.asarray(pc, dtype=np.float32)[:, :3]
        if gm is None or not np.any(gm):
            return pc, None
        keep = ~np.asarray(gm, dtype=bool)
        return pc[keep], keep

    def _upload(self, key: str, arr: np.ndarray) -> torch.Tensor:
        n = arr.shape[0]
        pin = self._pin[key].get(n)
        pin.numpy()[...] = arr
        dev = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        dev.copy_(pin, non_blocking=True)
        self.h2d_bytes += n * 12
        return dev

    def infer(self, frame: Dict) -> np.ndarray:
        """frame: pc0, pc1, pch1 [N,>=3] float32; gm0, gm1, gmh1 [N] bool (optional);
        pose0, pose1, poseh1 [4,4].  Returns final_flow [N0_all, 3] float32: pose flow for every
        point + network flow on the valid non-ground points (what the reference writes to the .h5)."""
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        pc0_all = np.asarray(frame["pc0"], dtype=np.float32)[:, :3]
        pc0, keep0 = self._strip(frame["pc0"], frame.get("gm0"))
        pc1, _ = self._strip(frame["pc1"], frame.get("gm1"))
        pch1, _ = self._strip(frame["pch1"], frame.get("gmh1"))
        T0 = cal_pose0to1(torch.as_tensor(frame["pose0"]), torch.as_tensor(frame["pose1"]))
        Th = cal_pose0to1(torch.as_tensor(frame["poseh1"]), torch.as_tensor(frame["pose1"]))
        n_all = pc0_all.shape[0]
        with torch.cuda.stream(self.stream):
            d0 = self._upload("pc0", pc0)
            d1 = self._upload("pc1", pc1)
            dh = self._upload("pch1", pch1)
            if keep0 is not None:
                d0_all = self._upload("pc0_all", pc0_all)
                src = np.full(n_all, -1, np.int32)
                src[keep0] = np.arange(int(keep0.sum()), dtype=np.int32)
                pin = self._pin_idx.get(n_all, 1)
                pin.numpy()[...] = src
                src_dev = torch.empty(n_all, dtype=torch.int32, device=self.device)
                src_dev.copy_(pin, non_blocking=True)
                self.h2d_bytes += n_all * 4
            else:
                d0_all, src_dev = d0, None
            out = self.net.forward_triple(dh, d0, d1, Th, T0, compact=False)
            T12 = T0[:3, :4].contiguous().float().flatten().to(self.device, non_blocking=True)
            final = torch.empty((n_all, 3), dtype=torch.float32, device=self.device)
            st = _lib.lib().himo_final_flow(_lib.ptr(d0_all), n_all, _lib.ptr(T12), _lib.ptr(out["flow_all"]),
                                            _lib.ptr(src_dev), _lib.ptr(final), _lib.stream_ptr(self.device))
            _lib.check(st, "himo_final_flow")
            host = self._pin["out"].get(n_all)
            host.copy_(final, non_blocking=True)
            self.d2h_bytes += n_all * 12
        self.stream.synchronize()
        return host.numpy().copy()

    # ------------------------------------------------------------------ pipelined streaming API
    def infer_stream(self, frames):
        """Iterate over host frames and yield `final_flow` arrays in order, with the H2D copy of frame i+1
        and the D2H copy of frame i-1 overlapping the network of frame i (three CUDA streams, two slots).
        Same results as `infer`; this is the call the multi-frame drivers (runner.run_save, bench.py) use."""
        dev = self.device
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._slots = [dict(pin={k: _Pinned(torch.float32) for k in ("pch1", "pc0", "pc1", "pc0_all", "out")},
                                pin_idx=_Pinned(torch.int32)) for _ in range(2)]
        pending = []          # (slot, ev_out, n_all)
        ev_done = [None, None]

        def launch(i, frame):
            slot = self._slots[i % 2]
            pc0_all = np.asarray(frame["pc0"], dtype=np.float32)[:, :3]
            pc0, keep0 = self._strip(frame["pc0"], frame.get("gm0"))
            pc1, _ = self._strip(frame["pc1"], frame.get("gm1"))
            pch1, _ = self._strip(frame["pch1"], frame.get("gmh1"))
            T0 = cal_pose0to1(torch.as_tensor(frame["pose0"]), torch.as_tensor(frame["pose1"]))
            Th = cal_pose0to1(torch.as_tensor(frame["poseh1"]), torch.as_tensor(frame["pose1"]))
            n_all = pc0_all.shape[0]
            h2d = d2h = 0
            with torch.cuda.stream(self._s_in):
                if ev_done[i % 2] is not None:          # the network has finished reading this slot's inputs
                    self._s_in.wait_event(ev_done[i % 2])
                devs = {}
                for key, arr in (("pc0", pc0), ("pc1", pc1), ("pch1", pch1)):
                    pin = slot["pin"][key].get(arr.shape[0])
                    pin.numpy()[...] = arr
                    d = torch.empty((arr.shape[0], 3), dtype=torch.float32, device=dev)
                    d.copy_(pin, non_blocking=True)
                    devs[key] = d
                    h2d += arr.shape[0] * 12
                if keep0 is not None:
                    pin = slot["pin"]["pc0_all"].get(n_all)
                    pin.numpy()[...] = pc0_all
                    d0_all = torch.empty((n_all, 3), dtype=torch.float32, device=dev)
                    d0_all.copy_(pin, non_blocking=True)
                    src = np.full(n_all, -1, np.int32)
                    src[keep0] = np.arange(int(keep0.sum()), dtype=np.int32)
                    pidx = slot["pin_idx"].get(n_all, 1)
                    pidx.numpy()[...] = src
                    src_dev = torch.empty(n_all, dtype=torch.int32, device=dev)
                    src_dev.copy_(pidx, non_blocking=True)
                    h2d += n_all * 16
                else:
                    d0_all, src_dev = devs["pc0"], None
                T12 = T0[:3, :4].contiguous().float().flatten().pin_memory().to(dev, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._s_in)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev_in)
                out = self.net.forward_triple(devs["pch1"], devs["pc0"], devs["pc1"], Th, T0, compact=False)
                final = torch.empty((n_all, 3), dtype=torch.float32, device=dev)
                st = _lib.lib().himo_final_flow(_lib.ptr(d0_all), n_all, _lib.ptr(T12), _lib.ptr(out["flow_all"]),
                                                _lib.ptr(src_dev), _lib.ptr(final), _lib.stream_ptr(dev))
                _lib.check(st, "himo_final_flow")
                ev = torch.cuda.Event()
                ev.record(self.stream)
                ev_done[i % 2] = ev
                for t in list(devs.values()) + [d0_all, T12, final] + ([src_dev] if src_dev is not None else []):
                    t.record_stream(self.stream)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev)
                host = slot["pin"]["out"].get(n_all)
                host.copy_(final, non_blocking=True)
                final.record_stream(self._s_out)
                ev_out = torch.cuda.Event()
                ev_out.record(self._s_out)
            self.h2d_bytes, self.d2h_bytes = h2d, n_all * 12
            return (host, ev_out)

        def collect(item):
            host, ev_out = item
            ev_out.synchronize()
            return host.numpy().copy()

        for i, frame in enumerate(frames):
            pending.append(launch(i, frame))
            if len(pending) == 2:            # the slot about to be reused must be drained first
                yield collect(pending.pop(0))
        while pending:
            yield collect(pending.pop(0))


class FastNSFEngine:
    """`InferenceRunner._process_step` for FastNSF (OSF/src/runner.py:130-155): ground removal, per-pair
    optimisation, final_flow = pose_flow everywhere + optimised flow on the non-ground points."""

    def __init__(self, device="cuda:0", precision: str = "fp32", seed: int = 0, **model_kw):
        from .fastnsf import FastNSF
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.net = FastNSF(device=self.device, precision=precision, seed=seed, **model_kw)

    def infer(self, frame: Dict) -> np.ndarray:
        from .deflowpp import rigid_flow
        pc0_all = torch.from_numpy(np.ascontiguousarray(np.asarray(frame["pc0"], np.float32)[:, :3])).to(self.device)
        pc1_all = torch.from_numpy(np.ascontiguousarray(np.asarray(frame["pc1"], np.float32)[:, :3])).to(self.device)
        gm0 = torch.from_numpy(np.asarray(frame.get("gm0", np.zeros(pc0_all.shape[0], bool)), bool)).to(self.device)
        gm1 = torch.from_numpy(np.asarray(frame.get("gm1", np.zeros(pc1_all.shape[0], bool)), bool)).to(self.device)
        batch = {"pc0": [pc0_all[~gm0].contiguous()], "pc1": [pc1_all[~gm1].contiguous()],
                 "pose0": [torch.as_tensor(frame["pose0"])], "pose1": [torch.as_tensor(frame["pose1"])]}
        res = self.net(batch)
        T = cal_pose0to1(torch.as_tensor(frame["pose0"]), torch.as_tensor(frame["pose1"]))
        final = rigid_flow(pc0_all, T)                          # pose flow for every point
        final[~gm0] = final[~gm0] + res["flow"][0]              # runner.py:149-155
        return final.cpu().numpy()


class NSFPEngine(FastNSFEngine):
    """Same step for NSFP (conf/model/nsfp.yaml): only the per-pair optimiser differs (himo_b200/nsfp.py)."""

    def __init__(self, device="cuda:0", seed: int = 0, **model_kw):
        from . import _lib
        from .nsfp import NSFP
        _lib.lib()                                              # fail now, not at the first frame, if the library is missing
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.manual_seed(seed)                                 # the networks are drawn from the global CPU RNG
        self.net = NSFP(**model_kw)
