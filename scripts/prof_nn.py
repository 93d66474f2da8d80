"""One chamfer3D.forward call per cloud kind (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import chamfer3d_ext, frames

kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["lidar", "uniform", "fixture"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
for kind in kinds:
    if kind == "fixture":
        z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "av2_fixture_clouds.npz"))
        a, b = z["pc0"].astype(np.float32)[:, :3], z["pc1"].astype(np.float32)[:, :3]
    elif kind == "uniform":
        a, b = frames.uniform_frame(n, 5003), frames.uniform_frame(n, 5004)
    else:
        tr = frames.lidar_triple(n, 5005)
        a, b = tr["pc0"], tr["pc1"]
    A, B = torch.from_numpy(a).cuda().contiguous(), torch.from_numpy(b).cuda().contiguous()
    d0 = torch.zeros(len(a), device="cuda"); d1 = torch.zeros(len(b), device="cuda")
    i0 = torch.zeros(len(a), dtype=torch.int32, device="cuda"); i1 = torch.zeros(len(b), dtype=torch.int32, device="cuda")
    for _ in range(2):
        chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
    torch.cuda.synchronize()
    print(kind, float(d0.mean() + d1.mean()))
