"""The reference's own CUDA kernels (oracle/_ref, recompiled for sm_100a) timed beside ours on the same B200 and the
same inputs, plus the raw agreement counts.  python scripts/bench_ref_kernels.py > out.json"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from himo_b200 import chamfer3d_ext, frames, mmcv_ext  # noqa: E402
from oracle import build_ref  # noqa: E402

ref_ch, ref_mm = build_ref.load("chamfer3D"), build_ref.load("mmcv")
VS, RNG = torch.tensor(frames.VOXEL_SIZE), torch.tensor(frames.POINT_CLOUD_RANGE)


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


out = {"bench": "reference_kernels_vs_ours", "device": torch.cuda.get_device_name(0), "cases": []}
z = np.load("tests/golden/av2_fixture_clouds.npz")
cases = [("av2_fixture_88k", z["pc0"].astype(np.float32), z["pc1"].astype(np.float32))]
tr = frames.lidar_triple(100000, 71)
cases.append(("lidar_100k", tr["pc0"][:, :3].copy(), tr["pc1"][:, :3].copy()))
for name, a_np, b_np in cases:
    a, b = torch.from_numpy(a_np).cuda().contiguous(), torch.from_numpy(b_np).cuda().contiguous()
    bufs = lambda: (torch.zeros(a.shape[0], device="cuda"), torch.zeros(b.shape[0], device="cuda"),
                    torch.zeros(a.shape[0], dtype=torch.int32, device="cuda"), torch.zeros(b.shape[0], dtype=torch.int32, device="cuda"))
    o, r = bufs(), bufs()
    chamfer3d_ext.forward(a, b, *o); ref_ch.forward(a, b, *r)
    rec = {"case": name, "n0": a.shape[0], "n1": b.shape[0],
           "idx_mismatches": int((o[2] != r[2]).sum() + (o[3] != r[3]).sum()),
           "dist_mismatches": int((o[0] != r[0]).sum() + (o[1] != r[1]).sum()),
           "chamfer_fwd_ms_ours": timed(lambda: chamfer3d_ext.forward(a, b, *o)),
           "chamfer_fwd_ms_reference": timed(lambda: ref_ch.forward(a, b, *r), reps=5)}
    co_o = a.new_zeros((a.shape[0], 3), dtype=torch.int32); co_r = a.new_zeros((a.shape[0], 3), dtype=torch.int32)
    mmcv_ext.dynamic_voxelize_forward(a, VS, RNG, co_o, 3); ref_mm.dynamic_voxelize_forward(a, VS, RNG, co_r, 3)
    rec["voxelize_row_mismatches"] = int((co_o != co_r).any(1).sum())
    rec["voxelize_ms_ours"] = timed(lambda: mmcv_ext.dynamic_voxelize_forward(a, VS, RNG, co_o, 3))
    rec["voxelize_ms_reference"] = timed(lambda: ref_mm.dynamic_voxelize_forward(a, VS, RNG, co_r, 3))
    feats = torch.randn(a.shape[0], 32, device="cuda")
    so = mmcv_ext.dynamic_point_to_voxel_forward(feats, co_r, "mean"); sr = ref_mm.dynamic_point_to_voxel_forward(feats, co_r, "mean")
    rec["scatter_voxels"] = int(sr[1].shape[0])
    rec["scatter_map_mismatches"] = int((so[2] != sr[2]).sum()) + int((so[1] != sr[1]).any(1).sum())
    rec["scatter_mean_max_abs_diff"] = float((so[0] - sr[0]).abs().max())
    rec["scatter_c32_ms_ours"] = timed(lambda: mmcv_ext.dynamic_point_to_voxel_forward(feats, co_r, "mean"))
    rec["scatter_c32_ms_reference"] = timed(lambda: ref_mm.dynamic_point_to_voxel_forward(feats, co_r, "mean"))
    out["cases"].append(rec)
print(json.dumps(out))
