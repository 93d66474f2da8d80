"""Leaf-operator stress bench (BASELINE.json configs[4]: 1M-point frames, KNN radius sweep, HBM roofline).

    python scripts/bench_leaf.py [--out gpurun_out/leaf.json] [--quick]

Times H1 (dynamic voxelize, dynamic point-to-voxel scatter) and H2 (bidirectional exact 1-NN) through the
C ABI with CUDA events on the launching stream, inputs rotated through more buffers than fit in the
126 MB L2, and reports achieved ALGORITHMIC GB/s (voxelize 24 B/pt, scatter
(28|144)(N+M) B, 1-NN 20 B/pt) as a fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from himo_b200 import chamfer3d_ext, frames, mmcv_ext  # noqa: E402


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except (OSError, KeyError):
        return 6650.0, "fallback"


def time_ms(fn, n_variants, reps=5, warm=2):
    """median-free mean over reps*n_variants launches; variant k uses its own buffers (L2 rotation)."""
    for k in range(warm):
        fn(k % n_variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for k in range(n_variants):
            fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_variants)


def variants_for(nbytes_per_variant):
    return max(2, min(64, int(400e6 / max(nbytes_per_variant, 1)) + 1))


def cloud(kind, n, seed):
    if kind == "uniform":
        return frames.uniform_frame(n, seed)
    tr = frames.lidar_triple(n, seed)
    return tr["pc0"]


def bench_voxelize(n, kind, peak):
    base = torch.from_numpy(cloud(kind, min(n, 1_000_000), 5001)).cuda()
    if n > base.shape[0]:
        base = base.repeat((n + base.shape[0] - 1) // base.shape[0], 1)[:n].contiguous()
    nv = variants_for(24 * n)
    pts = [base.clone() for _ in range(nv)]
    coors = [torch.zeros((n, 3), dtype=torch.int32, device="cuda") for _ in range(nv)]
    vs, cr = torch.tensor(frames.VOXEL_SIZE), torch.tensor(frames.POINT_CLOUD_RANGE)
    ms = time_ms(lambda k: mmcv_ext.dynamic_voxelize_forward(pts[k], vs, cr, coors[k]), nv)
    gbs = 24.0 * n / ms / 1e6
    return {"op": "dynamic_voxelize", "cloud": kind, "n": n, "ms": ms, "alg_bytes": 24 * n, "alg_gbs": gbs,
            "frac_hbm": gbs / peak, "l2_rotation_buffers": nv}


def bench_scatter(n, c, kind, peak):
    p = torch.from_numpy(cloud(kind, n, 5002)).cuda()
    coors = torch.zeros((n, 3), dtype=torch.int32, device="cuda")
    mmcv_ext.dynamic_voxelize_forward(p, torch.tensor(frames.VOXEL_SIZE), torch.tensor(frames.POINT_CLOUD_RANGE), coors)
    nv = variants_for((4 * c + 12) * n)
    feats = [torch.randn((n, c), device="cuda") for _ in range(nv)]
    cs = [coors.clone() for _ in range(nv)]
    m = int(mmcv_ext.dynamic_point_to_voxel_forward(feats[0], cs[0], "mean")[0].shape[0])
    ms = time_ms(lambda k: mmcv_ext.dynamic_point_to_voxel_forward(feats[k], cs[k], "mean"), nv)
    alg = (4 * c + 16) * n + (4 * c + 16) * m
    gbs = alg / ms / 1e6
    return {"op": f"dynamic_point_to_voxel(C={c},mean)", "cloud": kind, "n": n, "m": m, "ms": ms, "alg_bytes": alg,
            "alg_gbs": gbs, "frac_hbm": gbs / peak, "note": "python mirror incl. one .item() sync for M, as the reference"}


def bench_nn(n, kind, peak, radius=None):
    if kind == "fixture":
        z = np.load(os.path.join(ROOT, "tests", "golden", "av2_fixture_clouds.npz"))
        a, b = z["pc0"].astype(np.float32)[:, :3], z["pc1"].astype(np.float32)[:, :3]
    elif kind == "uniform":
        a, b = frames.uniform_frame(n, 5003), frames.uniform_frame(n, 5004)
    else:
        tr = frames.lidar_triple(n, 5005)
        a, b = tr["pc0"], tr["pc1"]
    n0, n1 = a.shape[0], b.shape[0]
    nv = variants_for(20 * (n0 + n1) * 3)
    A = [torch.from_numpy(a).cuda().contiguous() for _ in range(nv)]
    B = [torch.from_numpy(b).cuda().contiguous() for _ in range(nv)]
    d0 = torch.zeros(n0, device="cuda"); d1 = torch.zeros(n1, device="cuda")
    i0 = torch.zeros(n0, dtype=torch.int32, device="cuda"); i1 = torch.zeros(n1, dtype=torch.int32, device="cuda")
    if radius is None:
        fn = lambda k: chamfer3d_ext.forward(A[k], B[k], d0, d1, i0, i1)
    else:
        fn = lambda k: chamfer3d_ext.forward_radius(A[k], B[k], d0, d1, i0, i1, radius)
    ms = time_ms(fn, nv, reps=3)
    alg = 20 * (n0 + n1)
    gbs = alg / ms / 1e6
    out = {"op": "chamfer3D.forward (exact 1-NN both ways)" if radius is None else f"1-NN within r={radius} m",
           "cloud": kind, "n0": n0, "n1": n1, "ms": ms, "alg_bytes": alg, "alg_gbs": gbs, "frac_hbm": gbs / peak,
           "mqueries_per_s": (n0 + n1) / ms / 1e3}
    if radius is not None:
        out["found_frac"] = float((i0 >= 0).float().mean())
    else:
        out["chamfer"] = float(d0.mean() + d1.mean())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peak, src = peak_hbm()
    rows = []
    sizes = [100_000, 1_000_000] if args.quick else [100_000, 1_000_000, 16_000_000, 64_000_000]
    for n in sizes:
        rows.append(bench_voxelize(n, "lidar", peak)); print(json.dumps(rows[-1]), flush=True)
    for n in ([100_000] if args.quick else [100_000, 1_000_000]):
        for c in (3, 32):
            rows.append(bench_scatter(n, c, "lidar", peak)); print(json.dumps(rows[-1]), flush=True)
    rows.append(bench_nn(0, "fixture", peak)); print(json.dumps(rows[-1]), flush=True)
    for kind in ("lidar", "uniform"):
        for n in ([100_000] if args.quick else [100_000, 1_000_000]):
            rows.append(bench_nn(n, kind, peak)); print(json.dumps(rows[-1]), flush=True)
    if hasattr(chamfer3d_ext, "forward_radius"):
        for r in (0.5, 1.0, 1.414, 2.0, 4.4):
            rows.append(bench_nn(100_000 if args.quick else 1_000_000, "lidar", peak, radius=r))
            print(json.dumps(rows[-1]), flush=True)
    res = {"hbm_peak_gbs": peak, "peak_source": src, "rows": rows}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
