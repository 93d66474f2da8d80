"""Two FastNSF iterations at 100k points between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import fastnsf as F, frames, weights
from himo_b200.deflowpp import cal_pose0to1, rigid_flow
tr = frames.lidar_triple(100000, seed=2000, t=1.0)
pc0 = torch.from_numpy(np.ascontiguousarray(tr["pc0"][:, :3])).cuda(); pc1 = torch.from_numpy(np.ascontiguousarray(tr["pc1"][:, :3])).cuda()
PREC = sys.argv[1] if len(sys.argv) > 1 else "fp32"
net = F.FastNSF(itr_num=6, early_patience=0, precision=PREC)
sel0, _ = net.range_limit_(pc0); sel1, _ = net.range_limit_(pc1)
T = cal_pose0to1(torch.from_numpy(tr["pose0"]), torch.from_numpy(tr["pose1"]))
tr0 = (sel0 + rigid_flow(sel0.contiguous(), T)).contiguous(); sel1 = sel1.contiguous()
sd = weights.synth_neural_prior_state_dict(1)
lo, dims = F.volume_geometry(tr0, sel1, 10.0)
D = F.dt_build(sel1, lo, dims, 10.0)
net.optimize(tr0, sel1, init_state_dict=sd, D=D, lo=lo, dims=dims)
torch.cuda.synchronize()
net2 = F.FastNSF(itr_num=2, early_patience=0, precision=PREC)
torch.cuda.cudart().cudaProfilerStart()
net2.optimize(tr0, sel1, init_state_dict=sd, D=D, lo=lo, dims=dims)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
