"""Experiment: two SeFlow++ steps in flight on two CUDA streams (two networks' workspaces, shared nothing else) versus
one stream.  Does the second stream fill the tail waves / idle SMs of the first?  Timing with CUDA events over 60 steps."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from himo_b200 import _lib, weights  # noqa: E402
from himo_b200.deflowpp import DeFlowPP, cal_pose0to1  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sd = weights.synth_deflowpp_state_dict(0)
nets = [DeFlowPP(precision="fp32", device=dev, max_points=bench.N_POINTS)]
nets[0].load_state_dict(sd)
for _ in range(3):
    nets.append(nets[0].replica())
host_frames = bench.make_frames(0, 2)
frames = []
for fr in host_frames:
    d = {k: torch.from_numpy(fr[k]).to(dev) for k in ("pc0", "pc1", "pch1")}
    d["T0"] = cal_pose0to1(torch.from_numpy(fr["pose0"]), torch.from_numpy(fr["pose1"]))
    d["Th"] = cal_pose0to1(torch.from_numpy(fr["poseh1"]), torch.from_numpy(fr["pose1"]))
    frames.append(d)
streams = [torch.cuda.Stream(dev) for _ in range(4)]


def run(n_streams, steps):
    for i in range(6):
        with torch.cuda.stream(streams[i % n_streams]):
            d = frames[i % 2]
            nets[i % n_streams].forward_triple(d["pch1"], d["pc0"], d["pc1"], d["Th"], d["T0"], compact=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:n_streams]:
        s.wait_event(e0)
    for i in range(steps):
        with torch.cuda.stream(streams[i % n_streams]):
            d = frames[i % 2]
            nets[i % n_streams].forward_triple(d["pch1"], d["pc0"], d["pc1"], d["Th"], d["T0"], compact=False)
    for s in streams[:n_streams]:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
for pdl in (1, 0):
    _lib.lib().himo_conv_set_pdl(pdl)
    for ns in (1, 2, 3, 4, 1, 2, 3, 4):
        ms = run(ns, 60)
        out.setdefault(f"pdl{pdl}_streams{ns}", []).append(round(ms, 4))
print(json.dumps(out))
