"""Leaf kernels for an ncu counter capture (BASELINE configs[4]): voxelize at 16 M points, 1-NN on a 1 M-point lidar pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import chamfer3d_ext, frames, mmcv_ext
tr = frames.lidar_triple(1_000_000, 5005)
a, b = tr["pc0"], tr["pc1"]
big = torch.from_numpy(a).cuda().repeat(16, 1).contiguous()
coors = torch.zeros((big.shape[0], 3), dtype=torch.int32, device="cuda")
vs, cr = torch.tensor(frames.VOXEL_SIZE), torch.tensor(frames.POINT_CLOUD_RANGE)
for _ in range(3):
    mmcv_ext.dynamic_voxelize_forward(big, vs, cr, coors)
A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
d0 = torch.zeros(len(a), device="cuda"); d1 = torch.zeros(len(b), device="cuda")
i0 = torch.zeros(len(a), dtype=torch.int32, device="cuda"); i1 = torch.zeros(len(b), dtype=torch.int32, device="cuda")
for _ in range(2):
    chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
torch.cuda.synchronize()
print("ok", float(d0.mean() + d1.mean()))
