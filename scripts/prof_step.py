"""One SeFlow++ step (100k-point triple) between cudaProfilerStart/Stop for ncu (--profile-from-start off):
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py
    ncu --profile-from-start off --set full --clock-control none -k regex:k_conv -o gpurun_out/conv_full python scripts/prof_step.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from himo_b200 import frames, weights
from himo_b200.deflowpp import cal_pose0to1
from himo_b200.engine import SeFlowPPEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
sd = weights.synth_deflowpp_state_dict(0)
eng = SeFlowPPEngine(sd, device="cuda:0", max_points=n)
tr = frames.lidar_triple(n, seed=2000, t=1.0)
d = {k: torch.from_numpy(tr[k]).cuda() for k in ("pc0", "pc1", "pch1")}
T0 = cal_pose0to1(torch.from_numpy(tr["pose0"]), torch.from_numpy(tr["pose1"]))
Th = cal_pose0to1(torch.from_numpy(tr["poseh1"]), torch.from_numpy(tr["pose1"]))
for _ in range(3):
    eng.net.forward_triple(d["pch1"], d["pc0"], d["pc1"], Th, T0, compact=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.net.forward_triple(d["pch1"], d["pc0"], d["pc1"], Th, T0, compact=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
