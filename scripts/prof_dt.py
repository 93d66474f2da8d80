"""himo_nsf_dt_build at the bench pair between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import fastnsf as F, frames
from himo_b200.deflowpp import cal_pose0to1, rigid_flow
tr = frames.lidar_triple(100000, seed=2000, t=1.0)
pc0 = torch.from_numpy(np.ascontiguousarray(tr["pc0"][:, :3])).cuda(); pc1 = torch.from_numpy(np.ascontiguousarray(tr["pc1"][:, :3])).cuda()
net = F.FastNSF(itr_num=2, early_patience=0)
sel0, _ = net.range_limit_(pc0); sel1, _ = net.range_limit_(pc1)
T = cal_pose0to1(torch.from_numpy(tr["pose0"]), torch.from_numpy(tr["pose1"]))
tr0 = (sel0 + rigid_flow(sel0.contiguous(), T)).contiguous(); sel1 = sel1.contiguous()
lo, dims = F.volume_geometry(tr0, sel1, 10.0)
F.dt_build(sel1, lo, dims, 10.0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
F.dt_build(sel1, lo, dims, 10.0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
