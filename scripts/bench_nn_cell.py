"""1-NN time vs finest cell size (HIMO_NN_CELL) at a given cloud size; prints ms per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import chamfer3d_ext, frames
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
clouds = {}
tr = frames.lidar_triple(n, 5005)
clouds["lidar"] = (tr["pc0"], tr["pc1"])
clouds["uniform"] = (frames.uniform_frame(n, 5003), frames.uniform_frame(n, 5004))
for cell in (0.125, 0.25, 0.5, 1.0):
    chamfer3d_ext.CELL_SIZE = cell
    for kind, (a, b) in clouds.items():
        A, B = torch.from_numpy(a).cuda().contiguous(), torch.from_numpy(b).cuda().contiguous()
        d0 = torch.zeros(len(a), device="cuda"); d1 = torch.zeros(len(b), device="cuda")
        i0 = torch.zeros(len(a), dtype=torch.int32, device="cuda"); i1 = torch.zeros(len(b), dtype=torch.int32, device="cuda")
        for _ in range(2): chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
        torch.cuda.synchronize()
        print(f"cell {cell:5.3f} {kind:8s} n={n} {(time.perf_counter() - t) / 5 * 1e3:8.3f} ms  chamfer {float(d0.mean() + d1.mean()):.6f}", flush=True)
