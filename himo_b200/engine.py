"""Public inference API of the SeFlow++ path: host frames in, per-point total flow out.

`SeFlowPPEngine.infer(frame)` is the work of `ModelWrapper.test_step` minus the .h5 write
(OSF/src/trainer.py:290-343): ground removal, DeFlowPP.forward, pose-flow + network-flow assembly
for ALL points of pc0.  Inputs are host arrays (as the reference's DataLoader delivers them).
`infer_stream(frames)` is the multi-frame form every driver uses: preallocated slots (pinned host staging +
device buffers + a network replica with its own workspace and compute stream), one copy-in and one copy-out
stream, so that the H2D copy of frame i+1 and the D2H copy of frame i-1 overlap the network of frame i, TWO
networks are in flight at any time (the kernels of one fill the tail waves and idle SMs of the other: 2.44 ->
2.28 ms per frame, profiles/r02_two_streams.txt) and no frame allocates memory.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .deflowpp import DeFlowPP, cal_pose0to1


class _Slot:
    """One in-flight frame of the streaming engine: pinned host staging + device buffers, allocated once
    (grow-only) so that no frame of a run allocates device or pinned memory."""

    def __init__(self, device, reserve: int):
        self.device = device
        self.cap = 0
        self.reserve = int(reserve)
        self.ev_net = None                   # the network has finished reading this slot's inputs
        self.ev_out = None                   # the D2H copy of this slot's result has landed

    def ensure(self, n: int):
        if n <= self.cap:
            return
        cap = max(int(n * 1.25) + 16, self.reserve)
        dev = self.device
        self.pin = {k: torch.empty((cap, 3), dtype=torch.float32, pin_memory=True) for k in ("pch1", "pc0", "pc1", "pc0_all", "out")}
        self.pin_idx = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        self.pin_T = torch.empty(12, dtype=torch.float32, pin_memory=True)
        self.dev = {k: torch.empty((cap, 3), dtype=torch.float32, device=dev) for k in ("pch1", "pc0", "pc1", "pc0_all", "flow_all", "final")}
        self.dev_idx = torch.empty(cap, dtype=torch.int32, device=dev)
        self.dev_T = torch.empty(12, dtype=torch.float32, device=dev)
        self.cap = cap


class SeFlowPPEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", precision: str = "fp32",
                 max_points: int = 131072, n_slots: Optional[int] = None):
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.net = DeFlowPP(precision=precision, device=self.device, max_points=max_points)
        self.net.load_state_dict(state_dict)
        self._s_in, self._s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        n_slots = int(os.environ.get("HIMO_SLOTS", "3")) if n_slots is None else int(n_slots)
        if n_slots < 1:
            raise ValueError("n_slots must be >= 1")
        self._slots = [_Slot(self.device, max_points) for _ in range(n_slots)]
        for k, slot in enumerate(self._slots):      # every slot: its own workspace and compute stream, shared weights
            slot.net = self.net if k == 0 else self.net.replica()
            slot.stream = torch.cuda.Stream(self.device)
        self.stream = self._slots[0].stream
        if n_slots > 1 and os.environ.get("HIMO_PDL") is None:
            # With several networks in flight the tail of every kernel is filled by another stream's CTAs; dependents
            # launched early (programmatic dependent launch) would only hold SMs while they wait for their primary:
            # 453-461 frames/s without it against 444-447 with it (profiles/r02_two_streams.txt).  Process-wide knob.
            _lib.lib().himo_conv_set_pdl(0)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _strip(pc: np.ndarray, gm: Optional[np.ndarray]):
        pc = np.asarray(pc, dtype=np.float32)[:, :3]
        if gm is None or not np.any(gm):
            return pc, None
        keep = ~np.asarray(gm, dtype=bool)
        return pc[keep], keep

    # ------------------------------------------------------------------ one frame through a slot
    def _launch(self, slot: _Slot, frame: Dict):
        """Stage `frame` into `slot` and enqueue H2D (copy-in stream) -> network + final-flow assembly (compute
        stream) -> D2H (copy-out stream).  Returns the number of pc0 points; `slot.ev_out` fires when the result is
        in `slot.pin["out"]`."""
        pc0_all = np.asarray(frame["pc0"], dtype=np.float32)[:, :3]
        pc0, keep0 = self._strip(frame["pc0"], frame.get("gm0"))
        pc1, _ = self._strip(frame["pc1"], frame.get("gm1"))
        pch1, _ = self._strip(frame["pch1"], frame.get("gmh1"))
        # pose-flow assembly uses inv(pose1) @ pose0 (OSF/src/trainer.py:320-323); the network warp uses the stored
        # ego_motion when the frame carries one (wrap_batch_pcs, OSF/src/models/basic/__init__.py:44-50)
        T_pose = cal_pose0to1(torch.as_tensor(frame["pose0"]), torch.as_tensor(frame["pose1"]))
        T0 = torch.as_tensor(frame["ego_motion"]).detach().cpu().float() if frame.get("ego_motion") is not None else T_pose
        Th = cal_pose0to1(torch.as_tensor(frame["poseh1"]), torch.as_tensor(frame["pose1"]))
        n_all = pc0_all.shape[0]
        if slot.ev_out is not None:
            slot.ev_out.synchronize()           # the previous result of this slot has been consumed by the caller
        slot.ensure(max(n_all, pc1.shape[0], pch1.shape[0], 1))
        h2d = 0
        with torch.cuda.stream(self._s_in):
            if slot.ev_net is not None:
                self._s_in.wait_event(slot.ev_net)
            for key, arr in (("pc0", pc0), ("pc1", pc1), ("pch1", pch1)):
                n = arr.shape[0]
                slot.pin[key].numpy()[:n] = arr
                slot.dev[key][:n].copy_(slot.pin[key][:n], non_blocking=True)
                h2d += n * 12
            if keep0 is not None:
                slot.pin["pc0_all"].numpy()[:n_all] = pc0_all
                slot.dev["pc0_all"][:n_all].copy_(slot.pin["pc0_all"][:n_all], non_blocking=True)
                src = slot.pin_idx.numpy()[:n_all]
                src[...] = -1
                src[keep0] = np.arange(int(keep0.sum()), dtype=np.int32)
                slot.dev_idx[:n_all].copy_(slot.pin_idx[:n_all], non_blocking=True)
                h2d += n_all * 16
            slot.pin_T.numpy()[...] = T_pose[:3, :4].contiguous().float().flatten().numpy()
            slot.dev_T.copy_(slot.pin_T, non_blocking=True)
            h2d += 48
            ev_in = torch.cuda.Event()
            ev_in.record(self._s_in)
        with torch.cuda.stream(slot.stream):
            slot.stream.wait_event(ev_in)
            d0 = slot.dev["pc0"][:pc0.shape[0]]
            slot.net.forward_triple(slot.dev["pch1"][:pch1.shape[0]], d0, slot.dev["pc1"][:pc1.shape[0]], Th, T0,
                                    compact=False, flow_all_out=slot.dev["flow_all"])
            d0_all = slot.dev["pc0_all"] if keep0 is not None else slot.dev["pc0"]
            st = _lib.lib().himo_final_flow(_lib.ptr(d0_all), n_all, _lib.ptr(slot.dev_T), _lib.ptr(slot.dev["flow_all"]),
                                            _lib.ptr(slot.dev_idx) if keep0 is not None else None,
                                            _lib.ptr(slot.dev["final"]), _lib.stream_ptr(self.device))
            _lib.check(st, "himo_final_flow")
            slot.ev_net = torch.cuda.Event()
            slot.ev_net.record(slot.stream)
        with torch.cuda.stream(self._s_out):
            self._s_out.wait_event(slot.ev_net)
            slot.pin["out"][:n_all].copy_(slot.dev["final"][:n_all], non_blocking=True)
            slot.ev_out = torch.cuda.Event()
            slot.ev_out.record(self._s_out)
        self.h2d_bytes, self.d2h_bytes = h2d, n_all * 12
        return n_all

    @staticmethod
    def _collect(slot: _Slot, n_all: int) -> np.ndarray:
        slot.ev_out.synchronize()
        out = slot.pin["out"].numpy()[:n_all].copy()
        slot.ev_out = None
        return out

    def infer(self, frame: Dict) -> np.ndarray:
        """frame: pc0, pc1, pch1 [N,>=3] float32; gm0, gm1, gmh1 [N] bool (optional); pose0, pose1, poseh1 [4,4];
        ego_motion [4,4] (optional).  Returns final_flow [N0_all, 3] float32: pose flow for every point + network
        flow on the valid non-ground points (what the reference writes to the .h5).  One frame, synchronous."""
        slot = self._slots[0]
        return self._collect(slot, self._launch(slot, frame))

    # ------------------------------------------------------------------ pipelined streaming API
    def infer_stream(self, frames):
        """Iterate over host frames and yield `final_flow` arrays in order, with the H2D copy of frame i+1
        and the D2H copy of frame i-1 overlapping the networks of the frames in between (one compute stream and
        one workspace per slot: with three slots two networks are always in flight).  Same results as `infer`.
        This is the call bench.py's `e2e` arm times and the one the multi-frame drivers (runner.run_save,
        runner.run_validate) go through."""
        pending = []          # (slot, n_all)
        depth = len(self._slots)
        for i, frame in enumerate(frames):
            slot = self._slots[i % depth]
            pending.append((slot, self._launch(slot, frame)))
            if len(pending) == depth:        # the slot about to be reused must be drained first
                yield self._collect(*pending.pop(0))
        while pending:
            yield self._collect(*pending.pop(0))


class FastNSFEngine:
    """`InferenceRunner._process_step` for FastNSF (OSF/src/runner.py:130-155): ground removal, per-pair
    optimisation, final_flow = pose_flow everywhere + optimised flow on the non-ground points."""

    def __init__(self, device="cuda:0", precision: str = "fp32", seed: int = 0, n_workers: Optional[int] = None,
                 **model_kw):
        from .fastnsf import FastNSF
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.seed = int(seed)
        self._make_net = lambda: FastNSF(device=self.device, precision=precision, seed=seed, **model_kw)
        self.net = self._make_net()
        self.n_workers = int(os.environ.get("HIMO_NSF_WORKERS", "3")) if n_workers is None else int(n_workers)
        self._lanes = None                      # [(optimiser, stream)] of infer_stream, built on first use
        self._frame_no = 0
        self.last_iterations = []

    def _init_for(self, frame_no: int):
        """The reference draws a fresh Neural_Prior per pair from the global RNG (fastnsf.py:108-115); here pair k of an
        engine gets the seeded prior k, whichever worker optimises it."""
        from . import weights as W
        return W.synth_neural_prior_state_dict(self.seed * 1000003 + frame_no)

    def _prepare(self, frame: Dict) -> Dict:
        """Everything of one pair that does not depend on the optimiser: upload, ground removal, range limit,
        ego-motion warp and (FastNSF only) the distance volume of pc1 (fastnsf.py:180-201, 117-126).  Runs on
        the CURRENT torch stream and records an event."""
        from .deflowpp import rigid_flow
        pc0_all = torch.from_numpy(np.ascontiguousarray(np.asarray(frame["pc0"], np.float32)[:, :3])).to(self.device)
        pc1_all = torch.from_numpy(np.ascontiguousarray(np.asarray(frame["pc1"], np.float32)[:, :3])).to(self.device)
        gm0 = torch.from_numpy(np.asarray(frame.get("gm0", np.zeros(pc0_all.shape[0], bool)), bool)).to(self.device)
        gm1 = torch.from_numpy(np.asarray(frame.get("gm1", np.zeros(pc1_all.shape[0], bool)), bool)).to(self.device)
        if "ego_motion" in frame:
            T = torch.as_tensor(frame["ego_motion"]).detach().cpu().float()
        else:
            T = cal_pose0to1(torch.as_tensor(frame["pose0"]), torch.as_tensor(frame["pose1"]))
        pc0, pc1 = pc0_all[~gm0].contiguous(), pc1_all[~gm1].contiguous()
        prep = {"pc0_all": pc0_all, "gm0": gm0, "T": T, "pc0": pc0, "pc1": pc1}
        if self._prebuild_volume:
            from .fastnsf import dt_build, volume_geometry
            sel0, rm0 = self.net.range_limit_(pc0)           # (stateless helpers of the optimiser object)
            sel1, _ = self.net.range_limit_(pc1)
            tr0 = (sel0 + rigid_flow(sel0.contiguous(), T)).contiguous()
            lo, dims = volume_geometry(tr0, sel1.contiguous(), self.net.grid_factor)
            prep.update(tr0=tr0, sel1=sel1.contiguous(), rm0=rm0, lo=lo, dims=dims,
                        D=dt_build(sel1.contiguous(), lo, dims, self.net.grid_factor))
        ev = torch.cuda.Event()
        ev.record()
        prep["ready"] = ev
        return prep

    def _finish(self, prep: Dict, net=None, frame_no: Optional[int] = None) -> np.ndarray:
        from .deflowpp import rigid_flow
        net = self.net if net is None else net
        torch.cuda.current_stream(self.device).wait_event(prep["ready"])
        pc0_all, gm0 = prep["pc0_all"], prep["gm0"]
        if "D" in prep:
            res = net.optimize(prep["tr0"], prep["sel1"], D=prep["D"], lo=prep["lo"], dims=prep["dims"],
                               init_state_dict=None if frame_no is None else self._init_for(frame_no))
            flow = torch.zeros_like(prep["pc0"])
            flow[prep["rm0"]] = res["flow"]
        else:
            batch = {"pc0": [prep["pc0"]], "pc1": [prep["pc1"]], "ego_motion": [prep["T"]], "pose0": [None]}
            flow = net(batch)["flow"][0]
        final = rigid_flow(pc0_all, prep["T"])                  # pose flow for every point
        final[~gm0] = final[~gm0] + flow                        # runner.py:149-155
        return final.cpu().numpy()

    _prebuild_volume = True

    def infer(self, frame: Dict) -> np.ndarray:
        no = self._frame_no
        self._frame_no += 1
        out = self._finish(self._prepare(frame), frame_no=no if self._prebuild_volume else None)
        self.last_iterations = [int(self.net.last_info.get("iterations", 0))]
        return out

    def infer_stream(self, frames):
        """Several pairs in flight (BASELINE configs[2]: pairs resident per GPU): `n_workers` host threads, each with its own
        optimiser object (workspace) and CUDA stream, take the pairs in order; results come back in order.  The
        optimiser call blocks its thread on the device stop flag (ctypes releases the GIL), so the distance-volume
        build of one pair, the iteration kernels of another and the host-side polling overlap.  Same results as
        `infer`: the initial prior of pair k depends on k only."""
        if self.n_workers <= 1 or not self._prebuild_volume:
            for frame in frames:
                yield self.infer(frame)
            return
        from concurrent.futures import ThreadPoolExecutor
        _lib.lib().himo_nsf_set_blocking_poll(1)       # the workers sleep while their chunk of iterations runs (process-wide knob)
        if self._lanes is None:
            self._lanes = [(self.net if k == 0 else self._make_net(), torch.cuda.Stream(self.device))
                           for k in range(self.n_workers)]
        free = list(range(self.n_workers))

        def work(lane, frame, no):
            net, stream = self._lanes[lane]
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(stream):
                out = self._finish(self._prepare(frame), net=net, frame_no=no)
            return lane, out, int(net.last_info.get("iterations", 0))

        pending = []
        self.last_iterations = []
        with ThreadPoolExecutor(self.n_workers) as pool:
            for frame in frames:
                if not free:                                  # oldest first: results are yielded in order
                    lane, out, its = pending.pop(0).result()
                    free.append(lane)
                    self.last_iterations.append(its)
                    yield out
                no = self._frame_no
                self._frame_no += 1
                pending.append(pool.submit(work, free.pop(0), frame, no))
            while pending:
                lane, out, its = pending.pop(0).result()
                self.last_iterations.append(its)
                yield out
        _lib.lib().himo_nsf_set_blocking_poll(0)


class NSFPEngine(FastNSFEngine):
    """Same step for NSFP (conf/model/nsfp.yaml): only the per-pair optimiser differs (himo_b200/nsfp.py)."""

    _prebuild_volume = False                                    # NSFP has no distance volume: Chamfer inside the loop

    def __init__(self, device="cuda:0", seed: int = 0, **model_kw):
        from . import _lib
        from .nsfp import NSFP
        _lib.lib()                                              # fail now, not at the first frame, if the library is missing
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.manual_seed(seed)                                 # the networks are drawn from the global CPU RNG
        self.net = NSFP(**model_kw)
        self.seed, self.n_workers, self._frame_no, self.last_iterations = int(seed), 1, 0, []
