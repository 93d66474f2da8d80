"""Exact 1-NN (chamfer3D.forward): warp-cooperative search (one warp per query) against the round-1 kernel (one thread per
query) on the same clouds; results must be identical.  python scripts/bench_nn_ab.py [--big]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import chamfer3d_ext, frames, _lib
L = _lib.lib()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
big = "--big" in sys.argv
z = np.load(os.path.join(ROOT, "tests", "golden", "av2_fixture_clouds.npz"))
cases = [("fixture_88k", z["pc0"].astype(np.float32)[:, :3].copy(), z["pc1"].astype(np.float32)[:, :3].copy())]
tr = frames.lidar_triple(100_000, 5005); cases.append(("lidar_100k", tr["pc0"], tr["pc1"]))
cases.append(("uniform_100k", frames.uniform_frame(100_000, 5003)[:, :3].copy(), frames.uniform_frame(100_000, 5004)[:, :3].copy()))
if big:
    tr = frames.lidar_triple(1_000_000, 5005); cases.append(("lidar_1m", tr["pc0"], tr["pc1"]))
    cases.append(("uniform_1m", frames.uniform_frame(1_000_000, 5003)[:, :3].copy(), frames.uniform_frame(1_000_000, 5004)[:, :3].copy()))
out = []
for name, a, b in cases:
    A, B = torch.from_numpy(np.ascontiguousarray(a)).cuda(), torch.from_numpy(np.ascontiguousarray(b)).cuda()
    res = {}
    rec = {"case": name, "n0": int(A.shape[0]), "n1": int(B.shape[0])}
    for mode in (0, 1):
        L.himo_chamfer_set_warp_search(mode)
        d0 = torch.zeros(A.shape[0], device="cuda"); d1 = torch.zeros(B.shape[0], device="cuda")
        i0 = torch.zeros(A.shape[0], dtype=torch.int32, device="cuda"); i1 = torch.zeros(B.shape[0], dtype=torch.int32, device="cuda")
        for _ in range(3):
            chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        reps = 10
        for _ in range(reps):
            chamfer3d_ext.forward(A, B, d0, d1, i0, i1)
        e1.record(); torch.cuda.synchronize()
        rec["ms_warp" if mode else "ms_thread"] = e0.elapsed_time(e1) / reps
        res[mode] = (d0.clone(), d1.clone(), i0.clone(), i1.clone())
    rec["identical"] = all(torch.equal(x, y) for x, y in zip(res[0], res[1]))
    rec["mismatches"] = int(sum((x != y).sum() for x, y in zip(res[0], res[1])))
    for r in (1.0, 4.4):
        L.himo_chamfer_set_warp_search(1)
        d0 = torch.zeros(A.shape[0], device="cuda"); d1 = torch.zeros(B.shape[0], device="cuda")
        i0 = torch.zeros(A.shape[0], dtype=torch.int32, device="cuda"); i1 = torch.zeros(B.shape[0], dtype=torch.int32, device="cuda")
        chamfer3d_ext.forward_radius(A, B, d0, d1, i0, i1, r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            chamfer3d_ext.forward_radius(A, B, d0, d1, i0, i1, r)
        e1.record(); torch.cuda.synchronize()
        rec[f"ms_warp_radius_{r}"] = e0.elapsed_time(e1) / 5
    out.append(rec)
    print(json.dumps(rec), flush=True)
