"""Offline scorer of the public HiMo benchmark: the pair of tools either side of `save_zip.py`.

  * `save_zip_gt` (tools/test/save_zip_gt.py:64-170): ground-truth compensation distances per evaluated sweep, with the
    evaluation mask, category / instance labels, |gt flow| and pc0, as `<scene>/<timestamp>.feather` inside a zip;
  * `score` (tools/test/score.py:96-178, 545-667): ground-truth zip (or extracted directory) + prediction zip -> the
    CodaBench score dictionary (`mpe`, `chamfer`, `car_*`, `others_*`, `per_category`).

The metric arithmetic is `himo.InstanceMetrics` (one implementation for eval.py and for the scorer; the reference keeps
two copies, eval.py:50-149 and score.py:223-360, "matching exactly").  Host-side numpy / pandas like the reference.
"""
from __future__ import annotations

import json
import os
from io import BytesIO
from pathlib import Path
from typing import Dict, List, Optional, Tuple
from zipfile import ZipFile

import numpy as np

from . import himo


def write_gt_output_file(comp_dis, sweep_uuid: Tuple[str, str], output_dir, eval_mask, flow_category_indices=None,
                         flow_instance_id=None, gt_flow_norm=None, pc0=None) -> None:
    """save_zip_gt.py:64-108: column names and dtypes of the ground-truth feather."""
    import pandas as pd
    log_dir = Path(output_dir) / str(sweep_uuid[0])
    log_dir.mkdir(exist_ok=True, parents=True)
    cols = {f"comp_dis_{a}_m": np.asarray(comp_dis)[:, i].astype(np.float32) for i, a in enumerate("xyz")}
    cols["eval_mask"] = np.asarray(eval_mask).astype(np.uint8)
    if flow_category_indices is not None:
        cols["flow_category_indices"] = np.asarray(flow_category_indices).astype(np.uint8)
    if flow_instance_id is not None:
        cols["flow_instance_id"] = np.asarray(flow_instance_id).astype(np.uint32)
    if gt_flow_norm is not None:
        cols["gt_flow_norm"] = np.asarray(gt_flow_norm).astype(np.float32)
    if pc0 is not None:
        for i, a in enumerate("xyz"):
            cols[f"pc0_{a}"] = np.asarray(pc0)[:, i].astype(np.float32)
    pd.DataFrame(cols).to_feather(log_dir / f"{sweep_uuid[1]}.feather")


def save_zip_gt(data_dir: str, output_dir: str, res_name: str = "flow", store=None) -> str:
    """save_zip_gt.py:129-170 over the frame store."""
    from .dataset import HDF5Dataset
    os.makedirs(output_dir, exist_ok=True)
    data_name, _ = himo.check_valid(str(data_dir), res_name, None)
    ds = HDF5Dataset(data_dir, vis_name=res_name, eval=True, store=store)
    for i in range(len(ds)):
        data = ds[i]
        pc0 = data["pc0"]
        pose_flow = himo.pose_flow_np(pc0, data["pose0"], data["pose1"])
        dt0 = max(data["lidar_dt"]) - data["lidar_dt"]
        gt_flow = data["flow"] - pose_flow
        write_gt_output_file(himo.flow2compDis(gt_flow, dt0, sensor_dt=0.1), (data["scene_id"], str(data["timestamp"])),
                             output_dir, himo.eval_masks(data, data_name),
                             flow_category_indices=data.get("flow_category_indices"),
                             flow_instance_id=data.get("flow_instance_id"),
                             gt_flow_norm=np.linalg.norm(gt_flow, axis=1).astype(np.float32), pc0=pc0[:, :3])
    return himo.zip_res(output_dir, output_file=os.path.join(output_dir, f"{res_name}-submit.zip"))


def list_sweep_uuids(data_path: str) -> List[Tuple[str, str]]:
    """score.py:147-178: (scene, timestamp) of every `<scene>/<timestamp>.feather` in a zip or an extracted directory."""
    p = Path(data_path)
    if p.is_dir():
        names = [f.relative_to(p).as_posix() for f in p.rglob("*.feather")]
    else:
        with ZipFile(p, "r") as z:
            names = [n for n in z.namelist() if n.endswith(".feather")]
    out = []
    for n in names:
        parts = n.split("/")
        if len(parts) == 2:
            out.append((parts[0], parts[1][:-len(".feather")]))
    return out


def read_data_file(data_path: str, sweep_uuid: Tuple[str, str]):
    """score.py:96-144 -> (comp_dis, eval_mask, category, instance, gt_flow_norm, pc0); absent columns come back None
    (the mask defaults to all true)."""
    import pandas as pd
    rel = f"{sweep_uuid[0]}/{sweep_uuid[1]}.feather"
    p = Path(data_path)
    if p.is_dir():
        df = pd.read_feather(p / rel)
    else:
        with ZipFile(p, "r") as z:
            df = pd.read_feather(BytesIO(z.read(rel)))
    col = lambda name, dt: df[name].values.astype(dt) if name in df.columns else None
    comp = np.stack([df[f"comp_dis_{a}_m"].values.astype(np.float32) for a in "xyz"], axis=1)
    mask = col("eval_mask", bool)
    pc0 = np.stack([col(f"pc0_{a}", np.float32) for a in "xyz"], axis=1) if all(f"pc0_{a}" in df.columns for a in "xyz") else None
    return (comp, mask if mask is not None else np.ones(len(comp), bool), col("flow_category_indices", np.uint8),
            col("flow_instance_id", np.uint32), col("gt_flow_norm", np.float32), pc0)


def scores_from_metrics(m: himo.InstanceMetrics) -> Dict:
    """score.py:362-456: the flat CodaBench dictionary from the accumulated per-frame means."""
    s = m.summary()
    blank = {"overall": {"mpe": 0.0, "cd": 0.0, "std_mpe": 0.0, "std_cd": 0.0, "num_pts": 0, "num_obj": 0},
             "velocity": {r: {"mpe": 0.0, "cd": 0.0, "num_pts": 0, "num_obj": 0} for r in himo.RANGES}}
    per = {}
    for c in ("CAR", "OTHER_VEHICLES"):
        e = s.get(c, blank)
        per[c] = {"mpe_mean": e["overall"]["mpe"], "mpe_std": e["overall"]["std_mpe"], "cham_mean": e["overall"]["cd"],
                  "cham_std": e["overall"]["std_cd"], "num_pts": e["overall"]["num_pts"], "num_objs": e["overall"]["num_obj"],
                  "velocity": {r: dict(e["velocity"][r]) for r in himo.RANGES}}
    tot = s.get("Total", {"mpe": 0.0, "cd": 0.0, "num_pts": 0, "num_obj": 0})
    return {"mpe": tot["mpe"], "chamfer": tot["cd"], "num_frames": m.frame_cnt, "num_instances": tot["num_obj"],
            "total_points": tot["num_pts"],
            "car_cde": per["CAR"]["cham_mean"], "car_mpe": per["CAR"]["mpe_mean"],
            "car_num_objs": per["CAR"]["num_objs"], "car_num_pts": per["CAR"]["num_pts"],
            "others_cde": per["OTHER_VEHICLES"]["cham_mean"], "others_mpe": per["OTHER_VEHICLES"]["mpe_mean"],
            "others_num_objs": per["OTHER_VEHICLES"]["num_objs"], "others_num_pts": per["OTHER_VEHICLES"]["num_pts"],
            "per_category": per}


def score(gt_zip_path: str, pred_zip_path: str, output_dir: Optional[str] = None) -> Dict:
    """score.py:545-667.  Sweeps missing from the prediction or with a different point count are skipped and listed."""
    low = (str(gt_zip_path) + str(pred_zip_path)).lower()
    data_name = "av2" if ("av2" in low and "scania" not in low) else "scania"        # score.py:558-563 (default scania)
    pred = set(list_sweep_uuids(pred_zip_path))
    metrics = himo.InstanceMetrics(data_name)
    missing, mismatch = [], []
    for uuid in list_sweep_uuids(gt_zip_path):
        if uuid not in pred:
            missing.append(list(uuid))
            continue
        gt_dis, mask, cat, inst, speed, pc0 = read_data_file(gt_zip_path, uuid)
        est_dis = read_data_file(pred_zip_path, uuid)[0]
        if len(gt_dis) != len(est_dis):
            mismatch.append([list(uuid), len(gt_dis), len(est_dis)])
            continue
        if cat is None or inst is None:
            metrics.frame_cnt += 1                    # counted, nothing to evaluate (score.py:240-252)
            continue
        pc = pc0 if pc0 is not None else np.zeros_like(gt_dis)     # without pc0 the Chamfer term is taken on the distances
        if speed is None:                             # no velocity filter possible: every instance passes (score.py:290-292)
            speed = np.full(len(gt_dis), (metrics.min_vel + 1) * metrics.sensor_dt, np.float32)
        metrics.step_dis(pc[mask], gt_dis[mask], est_dis[mask], speed[mask], cat[mask], inst[mask])
    out = scores_from_metrics(metrics)
    out["missing_predictions"], out["point_count_mismatches"] = missing, mismatch
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
        with open(os.path.join(output_dir, "scores.json"), "w") as f:
            json.dump({k: v for k, v in out.items() if k != "per_category"}, f, indent=2)
    return out


def main(argv=None):
    """`python -m himo_b200.scoring --gt_zip G --pred_zip P [--output_dir D]`   (tools/test/score.py:669-760)
       `python -m himo_b200.scoring --save_gt --data_dir DIR --output_dir OUT [--res_name flow]`   (save_zip_gt.py)"""
    import sys
    from .runner import parse_overrides
    a = parse_overrides(sys.argv[1:] if argv is None else argv)
    if "save_gt" in a:
        print(save_zip_gt(a["data_dir"], a["output_dir"], a.get("res_name", "flow")))
        return
    if "gt_zip" not in a or "pred_zip" not in a:
        raise SystemExit("--gt_zip and --pred_zip are required")
    s = score(a["gt_zip"], a["pred_zip"], a.get("output_dir"))
    print(json.dumps({k: v for k, v in s.items() if k != "per_category"}, indent=2))


if __name__ == "__main__":
    main()
