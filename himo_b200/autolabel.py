"""`nnd` auto-label pass: OSF/process.py:106-172 `run_nnd`.

For every frame the ego-motion-compensated cloud is matched against the neighbouring sweep and a point is labelled
moving when its nearest neighbour is at least `min_nnd` (1.4 m/s x 0.1 s = 0.14 m; 0.32 for Scania) and less than
4.4 m (160 km/h x 0.1 s) away (process.py:119-125).  The reference runs a full bidirectional chamfer3D pass and
thresholds afterwards; only distances below 4.4 m matter, so the radius-limited exact search
(`himo_chamfer_forward_radius`) gives the identical labels: beyond the radius it reports 1e20, which fails the
`< truncated^2` test exactly like the true distance would.
The scene loop mirrors the reference: frames of a scene in index order, the LAST frame of a scene is matched against
its predecessor (process.py:164), poses are normalised to the first frame of the scene (process.py:152-157), results
go to the frame store under the key `nnd` as uint8.
One deliberate difference: with `overwrite=True` on a scene that already holds `nnd` the reference walks the scene
without recomputing anything (process.py:147-166, `exist_dict["nnd"]` stays True); here the labels are recomputed.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import chamfer3d_ext
from .dataset import HDF5Dataset

TRUNCATED_M = 4.4


def cuda_nnd(pc0: torch.Tensor, pc1: torch.Tensor, moving_threshold: float = 0.14, truncated: float = TRUNCATED_M) -> np.ndarray:
    """process.py:119-125.  pc0 [N,3] (already ego-motion transformed), pc1 [M,3]: CUDA float32 -> uint8 labels [N]."""
    pc0, pc1 = pc0.contiguous().float(), pc1.contiguous().float()
    n0, n1 = pc0.shape[0], pc1.shape[0]
    if n0 == 0:
        return np.zeros(0, np.uint8)
    d0 = torch.empty(n0, device=pc0.device); d1 = torch.empty(n1, device=pc0.device)
    i0 = torch.empty(n0, dtype=torch.int32, device=pc0.device); i1 = torch.empty(n1, dtype=torch.int32, device=pc0.device)
    chamfer3d_ext.forward_radius(pc0, pc1, d0, d1, i0, i1, float(truncated))
    # the reference compares a float32 numpy array with python scalars, i.e. in float32 (the scalar is cast down)
    label = (d0 >= pow(moving_threshold, 2)) & (d0 < pow(truncated, 2))
    return label.to(torch.uint8).cpu().numpy()


def npcal_pose0to1(pose0: np.ndarray, pose1: np.ndarray) -> np.ndarray:
    """OSF/src/utils/mics.py npcal_pose0to1: inv(pose1) @ pose0."""
    return np.linalg.inv(pose1) @ pose0


def run_nnd(data_dir: str, scene_range: Sequence[int] = (-1, -1), overwrite: bool = True, min_nnd: float = 0.14,
            device: str = "cuda", store=None) -> int:
    """Label every frame of the selected scenes; returns the number of frames written."""
    if not torch.cuda.is_available():
        raise EnvironmentError("No cuda available, please check your cuda environment.")      # process.py:114-115
    dataset = HDF5Dataset(data_dir, store=store)
    st = dataset.store
    written = 0
    for si, scene_id in enumerate(dataset.scene_id_bounds.keys()):
        if scene_range[0] != -1 and scene_range[-1] != -1 and (si < scene_range[0] or si >= scene_range[1]):
            continue
        b = dataset.scene_id_bounds[scene_id]
        idxs = range(b["min_index"], b["max_index"] + 1)
        if not overwrite and all(st.has(scene_id, dataset.data_index[i][1], "nnd") for i in idxs):
            continue
        norm = st.read(scene_id, dataset.data_index[b["min_index"]][1], "pose")
        for i in idxs:
            ts = dataset.data_index[i][1]
            j = i - 1 if i == b["max_index"] else i + 1
            ts1 = dataset.data_index[j][1]
            pc0 = st.read(scene_id, ts, "lidar")[:, :3]
            pose0 = npcal_pose0to1(st.read(scene_id, ts, "pose"), norm)
            pose1 = npcal_pose0to1(st.read(scene_id, ts1, "pose"), norm)
            ego = npcal_pose0to1(pose0, pose1)
            tr0 = pc0 @ ego[:3, :3].T + ego[:3, 3]
            pc1 = st.read(scene_id, ts1, "lidar")[:, :3]
            lab = cuda_nnd(torch.tensor(tr0, dtype=torch.float32, device=device),
                           torch.tensor(np.ascontiguousarray(pc1), dtype=torch.float32, device=device), moving_threshold=min_nnd)
            st.write(scene_id, ts, "nnd", lab.astype(np.uint8))
            written += 1
    return written


def shift_cluster_id(cluster: np.ndarray) -> np.ndarray:
    """OSF/src/autolabel.py:18-29: 0 background, 1 reserved for "dynamic but unclustered", ids >= 1 move up by one."""
    cluster = np.asarray(cluster)
    return np.where(cluster > 0, cluster + 1, 0).astype(cluster.dtype)


def seflow_auto(input_data) -> np.ndarray:
    """OSF/src/autolabel.py:32-36 (HiMo Fig. 6 top): DUFOMap dynamic points keep their (shifted) cluster id."""
    dufo = np.asarray(input_data["dufo"][:]).astype(np.uint8)
    cluster = shift_cluster_id(np.asarray(input_data["dufocluster"][:]).astype(np.int16))
    cluster[dufo == 0] = 0
    return cluster


def seflowpp_auto(input_data, tau1: float = 0.05, tau2: float = 0.30) -> np.ndarray:
    """OSF/src/autolabel.py:39-62 (HiMo Eq. 5, Fig. 6 bottom): a cluster is dynamic when the fractions of its points
    flagged by DUFOMap and by `nnd` satisfy min > tau1 and max > tau2; its points then carry the cluster id, all others
    0.  Same result as the reference's loop over cluster ids, computed with three bincounts.  The output keeps the
    reference's dtype (that of `dufo`, uint8: ids wrap modulo 256 there too)."""
    dufo = np.asarray(input_data["dufo"][:]).astype(np.uint8)
    cluster = shift_cluster_id(np.asarray(input_data["cluster"][:]).astype(np.int16))
    nnd = np.asarray(input_data["nnd"][:]).astype(np.uint8)
    dynamic = np.zeros_like(dufo)
    if cluster.size == 0:
        return dynamic
    ids = np.where(cluster > 1, cluster, 0).astype(np.int64)            # 0 and 1 never qualify
    k = int(ids.max()) + 1
    total = np.bincount(ids, minlength=k)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_dufo = np.bincount(ids, weights=dufo > 0, minlength=k) / total
        r_nnd = np.bincount(ids, weights=nnd > 0, minlength=k) / total
    keep = (np.minimum(r_dufo, r_nnd) > tau1) & (np.maximum(r_dufo, r_nnd) > tau2)
    keep[0] = False
    sel = keep[ids]
    dynamic[sel] = cluster[sel]                                          # numpy casts int16 -> uint8 like the reference
    return dynamic


def main(argv=None):
    """`python -m himo_b200.autolabel --data_dir DIR [--scene_range 0,10] [--overwrite false] [--min_nnd 0.14]`:
    the fire entry the reference keeps commented at process.py:321.  Under torchrun the scenes of the range are dealt
    round-robin to the ranks (independent scenes, no collective)."""
    import os
    import sys
    from .runner import parse_overrides
    a = parse_overrides(sys.argv[1:] if argv is None else argv)
    if "data_dir" not in a:
        raise SystemExit("--data_dir is required")
    rng = [int(v) for v in a.get("scene_range", "-1,-1").strip("[]()").split(",")]
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    overwrite = a.get("overwrite", "true").lower() in ("1", "true", "yes")
    min_nnd = float(a.get("min_nnd", 0.14))
    if world == 1:
        n = run_nnd(a["data_dir"], rng, overwrite, min_nnd)
    else:
        torch.cuda.set_device(local)
        ds = HDF5Dataset(a["data_dir"])
        lo, hi = (0, len(ds.scene_id_bounds)) if -1 in (rng[0], rng[-1]) else (rng[0], min(rng[1], len(ds.scene_id_bounds)))
        n = sum(run_nnd(a["data_dir"], [s, s + 1], overwrite, min_nnd, device=f"cuda:{local}", store=ds.store)
                for s in range(lo + rank, hi, world))
    print(f"[rank {rank}] nnd labels written for {n} frames")


if __name__ == "__main__":
    main()
