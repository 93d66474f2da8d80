"""HiMo motion compensation and its metrics: flow -> compensation distance -> undistorted points,
instance-level MPE / Chamfer (CDE), prediction zip I/O.

Restates (does not import) the reference: utils/__init__.py:4-47 (check_valid, ego_pts_mask,
flow2compDis, refine_pts), eval.py:26-268 (InstanceMetrics) and save_zip.py:30-100 (feather/zip format:
members `<scene_id>/<timestamp>.feather`, float32 columns comp_dis_x_m / comp_dis_y_m / comp_dis_z_m).
"""
from __future__ import annotations

import json
import os
import shutil
from io import BytesIO
from pathlib import Path
from typing import Dict, Optional, Tuple
from zipfile import ZipFile

import numpy as np

# AV2 annotation table (index = position + 1, NONE = 0): OSF/src/utils/av2_eval.py:28-75, also
# tools/test/score.py:29-94
ANNOTATION_CATEGORIES = [
    "ANIMAL", "ARTICULATED_BUS", "BICYCLE", "BICYCLIST", "BOLLARD", "BOX_TRUCK", "BUS", "CONSTRUCTION_BARREL",
    "CONSTRUCTION_CONE", "DOG", "LARGE_VEHICLE", "MESSAGE_BOARD_TRAILER", "MOBILE_PEDESTRIAN_CROSSING_SIGN",
    "MOTORCYCLE", "MOTORCYCLIST", "OFFICIAL_SIGNALER", "PEDESTRIAN", "RAILED_VEHICLE", "REGULAR_VEHICLE",
    "SCHOOL_BUS", "SIGN", "STOP_SIGN", "STROLLER", "TRAFFIC_LIGHT_TRAILER", "TRUCK", "TRUCK_CAB",
    "VEHICULAR_TRAILER", "WHEELCHAIR", "WHEELED_DEVICE", "WHEELED_RIDER"]
CATEGORY_TO_INDEX = {"NONE": 0, **{c: i + 1 for i, c in enumerate(ANNOTATION_CATEGORIES)}}
BUCKETED_METACATAGORIES = {
    "CAR": ["REGULAR_VEHICLE"],
    "OTHER_VEHICLES": ["BOX_TRUCK", "LARGE_VEHICLE", "RAILED_VEHICLE", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER",
                       "ARTICULATED_BUS", "BUS", "SCHOOL_BUS"],
}
CLOSE_DISTANCE_THRESHOLD = 35.0
RANGES = ["0-10", "10-20", "20-30", "30+"]


def check_valid(data_dir: str, flow_mode: str, comp_dis_zip: Optional[str] = None) -> Tuple[str, int]:
    """utils/__init__.py:4-24: dataset name from the path, 1 = evaluate a zip, 2 = evaluate stored flow."""
    d = str(data_dir)
    if d.find("Scania") > 0 or d.find("scania") > 0:
        name = "scania"
    elif d.find("av2") > 0 or d.find("AV2") > 0:
        name = "av2"
    else:
        raise ValueError("Unknown dataset name in data_dir.")
    if comp_dis_zip and os.path.exists(comp_dis_zip):
        return name, 1
    return name, 2


def ego_pts_mask(pts, min_bound=(-9.5, -3 / 2, 0), max_bound=(5, 2.760004 / 2, 5)) -> np.ndarray:
    """True for points OUTSIDE the ego-vehicle box (utils/__init__.py:26-34)."""
    inside = ((pts[:, 0] > min_bound[0]) & (pts[:, 0] < max_bound[0]) & (pts[:, 1] > min_bound[1]) &
              (pts[:, 1] < max_bound[1]) & (pts[:, 2] > min_bound[2]) & (pts[:, 2] < max_bound[2]))
    return ~inside


def flow2compDis(flow, dt0, sensor_dt=10):
    """comp_dis = flow / sensor_dt * dt0 (utils/__init__.py:36-43)."""
    return flow / sensor_dt * dt0[:, None]


def refine_pts(pc, ds):
    return pc[:, :3] + ds


def pose_flow_np(pc0, pose0, pose1):
    """save_zip.py:114-116 / eval.py:283-285: numpy, in the dtype of the stored poses."""
    ego = np.linalg.inv(pose1) @ pose0
    return pc0[:, :3] @ ego[:3, :3].T + ego[:3, 3] - pc0[:, :3]


def comp_dis_from_total_flow(data: Dict, res_name: str) -> np.ndarray:
    """save_zip.py:112-121: stored total flow -> per-point compensation distance to the latest point."""
    pf = pose_flow_np(data["pc0"], data["pose0"], data["pose1"])
    est = np.zeros_like(pf) if res_name == "raw" else (data[res_name] - pf)
    dt0 = max(data["lidar_dt"]) - data["lidar_dt"]
    return flow2compDis(est, dt0, sensor_dt=0.1)


# ------------------------------------------------------------------------------------ zip I/O
def write_output_file(comp_dis: np.ndarray, sweep_uuid: Tuple[str, str], output_dir) -> None:
    import pandas as pd
    out = Path(output_dir) / sweep_uuid[0]
    out.mkdir(exist_ok=True, parents=True)
    df = pd.DataFrame({"comp_dis_x_m": comp_dis[:, 0].astype(np.float32),
                       "comp_dis_y_m": comp_dis[:, 1].astype(np.float32),
                       "comp_dis_z_m": comp_dis[:, 2].astype(np.float32)})
    df.to_feather(out / f"{sweep_uuid[1]}.feather")


def zip_res(res_folder, output_file="submit.zip") -> str:
    res_folder = str(res_folder)
    scenes = [f for f in os.listdir(res_folder) if os.path.isdir(os.path.join(res_folder, f))]
    with ZipFile(output_file, "w") as z:
        for scene in scenes:
            for log in sorted(os.listdir(os.path.join(res_folder, scene))):
                if log.endswith(".feather"):
                    z.write(os.path.join(res_folder, scene, log), arcname=os.path.join(scene, log))
    for scene in scenes:
        shutil.rmtree(os.path.join(res_folder, scene))
    return output_file


def read_output_zip(zip_path: str, sweep_uuid: Tuple[str, str]) -> np.ndarray:
    import pandas as pd
    with ZipFile(zip_path, "r") as z:
        with z.open(f"{sweep_uuid[0]}/{sweep_uuid[1]}.feather") as f:
            df = pd.read_feather(BytesIO(f.read()))
    return np.stack([df[c].values.astype(np.float32) for c in ("comp_dis_x_m", "comp_dis_y_m", "comp_dis_z_m")], 1)


# ------------------------------------------------------------------------------------ metrics
def chamfer_mean_nn(a: np.ndarray, b: np.ndarray) -> float:
    """CDE of eval.py:50-62: (mean NN distance a->b + mean NN distance b->a) / 2, unsquared."""
    if len(a) == 0 or len(b) == 0:
        return float("nan")
    from scipy.spatial import cKDTree
    d12, _ = cKDTree(b).query(a, k=1)
    d21, _ = cKDTree(a).query(b, k=1)
    return float((np.nanmean(d12) + np.nanmean(d21)) / 2.0)


def chamfer_mean_nn_batched(pairs, device) -> list:
    """The same CDE for MANY instances in one device launch (himo_segmented_nn, csrc/segnn.cu): `pairs` is a list of
    (a [na,3], b [nb,3]) arrays, one per instance; returns one float per pair.  float64 brute force on the device --
    equal to `chamfer_mean_nn` (cKDTree, float64) to rounding."""
    import ctypes
    import torch
    from . import _lib
    if not pairs:
        return []
    L = _lib.lib()
    if not getattr(L, "_segnn_registered", False):
        L.himo_segmented_nn.restype = ctypes.c_int
        L.himo_segmented_nn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
        L._segnn_registered = True
    dev = torch.device(device)
    a_off = np.zeros(len(pairs) + 1, np.int32)
    b_off = np.zeros(len(pairs) + 1, np.int32)
    a_off[1:] = np.cumsum([len(a) for a, _ in pairs])
    b_off[1:] = np.cumsum([len(b) for _, b in pairs])
    a = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(x, np.float64).reshape(-1, 3) for x, _ in pairs]))).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(y, np.float64).reshape(-1, 3) for _, y in pairs]))).to(dev)
    ao, bo = torch.from_numpy(a_off).to(dev), torch.from_numpy(b_off).to(dev)
    da = torch.empty(a.shape[0], dtype=torch.float64, device=dev)
    db = torch.empty(b.shape[0], dtype=torch.float64, device=dev)
    biggest = int(max(np.diff(a_off).max(), np.diff(b_off).max()))
    with _lib.on_device(dev):
        _lib.check(L.himo_segmented_nn(_lib.ptr(a), _lib.ptr(ao), _lib.ptr(b), _lib.ptr(bo), len(pairs), biggest,
                                       _lib.ptr(da), _lib.ptr(db), _lib.stream_ptr(dev)), "himo_segmented_nn")
    da, db = da.cpu().numpy(), db.cpu().numpy()
    out = []
    for k in range(len(pairs)):
        xa, xb = da[a_off[k]:a_off[k + 1]], db[b_off[k]:b_off[k + 1]]
        out.append(float("nan") if len(xa) == 0 or len(xb) == 0 else float((np.nanmean(xa) + np.nanmean(xb)) / 2.0))
    return out


def _bucket(v: float) -> Optional[str]:
    if 0 < v < 10:
        return "0-10"
    if 10 <= v < 20:
        return "10-20"
    if 20 <= v < 30:
        return "20-30"
    if v >= 30:
        return "30+"
    return None


class InstanceMetrics:
    """Per class {CAR, OTHER_VEHICLES} x per instance (>= 10 points, mean GT speed >= 3 m/s, 1.5 for
    Scania) MPE and CDE of the compensated points, bucketed by velocity and distance
    (eval.py:26-149); `summary()` / `print()` follow eval.py:151-268."""

    def __init__(self, data_name: str, sensor_hz: float = 10.0, device=None):
        # device = "cuda[:i]": the per-instance Chamfer of a frame runs as ONE batched device launch (himo_segmented_nn)
        # instead of one cKDTree build + query pair per instance; None = the reference's host path (scipy)
        self.device = device
        self.frame_cnt = 0
        self.sensor_dt = 1.0 / sensor_hz
        self.data_name = data_name
        self.min_vel = 1.5 if data_name == "scania" else 3.0
        self.evaluate_data = self._blank()

    @staticmethod
    def _blank():
        mk = lambda: {"num_pts": [], "mpe": [], "cham": [], "std_mpe": [], "std_cham": []}
        return {c: {"vel": {r: mk() for r in RANGES}, "dis": {r: mk() for r in RANGES}, "mean": mk()}
                for c in ("CAR", "OTHER_VEHICLES")}

    def step_eval(self, pc, gt_flow, pc_dt0, gt_category, gt_instance, est_flow=None, est_dis=None):
        if est_flow is not None:
            est_dis = flow2compDis(est_flow, pc_dt0, sensor_dt=self.sensor_dt)
        self._step(pc, flow2compDis(gt_flow, pc_dt0, sensor_dt=self.sensor_dt), est_dis,
                   np.linalg.norm(gt_flow, axis=1), gt_category, gt_instance)

    def step_dis(self, pc, gt_dis, est_dis, gt_flow_norm, gt_category, gt_instance):
        """The same frame step from compensation distances (the offline scorer's inputs: tools/test/score.py:223-360
        reads `comp_dis`, `gt_flow_norm`, labels and pc0 back from the GT / prediction zips)."""
        self._step(pc, gt_dis, est_dis, gt_flow_norm, gt_category, gt_instance)

    def _step(self, pc, gt_dis, est_dis, gt_speed, gt_category, gt_instance):
        frame = self._blank()
        refine = refine_pts(pc, est_dis)
        gt_refine = refine_pts(pc, gt_dis)
        todo = []          # eligible instances of this frame, in the reference's visiting order
        for cname in ("CAR", "OTHER_VEHICLES"):
            ids = np.array([CATEGORY_TO_INDEX[c] for c in BUCKETED_METACATAGORIES[cname]])
            mc = np.isin(gt_category, ids)
            if mc.sum() == 0:
                continue
            ins_c, speed_c, ref_c, gtref_c, pc_c = gt_instance[mc], gt_speed[mc], refine[mc], gt_refine[mc], pc[mc]
            for ins in np.unique(ins_c):
                m = ins_c == ins
                npts = int(m.sum())
                vel = speed_c[m].mean() / self.sensor_dt
                if npts < 10 or vel < self.min_vel:
                    continue
                dis = np.linalg.norm(pc_c[m], axis=1).mean()
                mpe = np.linalg.norm(gtref_c[m] - ref_c[m], axis=1).mean()
                todo.append((cname, npts, vel, dis, mpe, gtref_c[m], ref_c[m]))
        if self.device is not None:
            chams_all = chamfer_mean_nn_batched([(t[5], t[6]) for t in todo], self.device)
        else:
            chams_all = [chamfer_mean_nn(t[5], t[6]) for t in todo]
        for (cname, npts, vel, dis, mpe, _, _), cham in zip(todo, chams_all):
            for metric, val in (("vel", vel), ("dis", dis)):
                r = _bucket(val)
                if r is None:
                    continue
                frame[cname][metric][r]["num_pts"].append(npts)
                frame[cname][metric][r]["mpe"].append(mpe)
                frame[cname][metric][r]["cham"].append(cham)
        for cname in frame:
            tot, mpes, chams = [], [], []
            for metric in ("vel", "dis"):
                for r in RANGES:
                    cell = frame[cname][metric][r]
                    if not cell["num_pts"]:
                        continue
                    for k in ("num_pts", "mpe", "cham"):
                        self.evaluate_data[cname][metric][r][k] += cell[k]
                    if metric == "vel":
                        mpes.append(np.average(cell["mpe"], weights=cell["num_pts"]))
                        chams.append(np.average(cell["cham"], weights=cell["num_pts"]))
                        tot.append(sum(cell["num_pts"]))
            if sum(tot) == 0:
                continue
            mean = self.evaluate_data[cname]["mean"]
            mean["num_pts"].append(sum(tot))
            mean["mpe"].append(np.nanmean(mpes)); mean["cham"].append(np.nanmean(chams))
            mean["std_mpe"].append(np.nanstd(mpes)); mean["std_cham"].append(np.nanstd(chams))
        self.frame_cnt += 1

    # ---- aggregation (eval.py:151-268)
    def summary(self) -> Dict:
        avg = lambda v, w: float(np.average(v, weights=w)) if len(v) > 0 and np.sum(w) > 0 else 0.0
        std = lambda v: float(np.std(v)) if len(v) > 0 else 0.0
        out, tot = {}, {"mpe": [], "cham": [], "num_pts": []}
        for c in ("CAR", "OTHER_VEHICLES"):
            mean = self.evaluate_data[c]["mean"]
            if not mean["num_pts"]:
                continue
            entry = {"overall": {"mpe": avg(mean["mpe"], mean["num_pts"]), "cd": avg(mean["cham"], mean["num_pts"]),
                                 "std_mpe": std(mean["std_mpe"]), "std_cd": std(mean["std_cham"]),
                                 "num_pts": int(np.sum(mean["num_pts"])), "num_obj": len(mean["num_pts"])},
                     "velocity": {}, "distance": {}}
            for key, metric in (("velocity", "vel"), ("distance", "dis")):
                for r in RANGES:
                    cell = self.evaluate_data[c][metric][r]
                    entry[key][r] = {"mpe": avg(cell["mpe"], cell["num_pts"]), "cd": avg(cell["cham"], cell["num_pts"]),
                                     "num_pts": int(np.sum(cell["num_pts"])), "num_obj": len(cell["num_pts"])}
            out[c] = entry
            for k in ("mpe", "cham", "num_pts"):
                tot[k] += mean[k]
        if tot["num_pts"]:
            out["Total"] = {"mpe": avg(tot["mpe"], tot["num_pts"]), "cd": avg(tot["cham"], tot["num_pts"]),
                            "num_pts": int(np.sum(tot["num_pts"])), "num_obj": len(tot["num_pts"])}
        return out

    def merge(self, other: "InstanceMetrics") -> None:
        """Combine per-rank metrics (the reference merges gathered OfficialMetrics, OSF/src/runner.py:262-289)."""
        for c in self.evaluate_data:
            for metric in ("vel", "dis"):
                for r in RANGES:
                    for k in ("num_pts", "mpe", "cham"):
                        self.evaluate_data[c][metric][r][k] += other.evaluate_data[c][metric][r][k]
            for k in self.evaluate_data[c]["mean"]:
                self.evaluate_data[c]["mean"][k] += other.evaluate_data[c]["mean"][k]
        self.frame_cnt += other.frame_cnt

    def print(self, res_name="flow", file_name="result_av2.json") -> Dict:
        from tabulate import tabulate
        s = self.summary()
        data = {}
        if os.path.exists(file_name):
            try:
                data = json.load(open(file_name))
            except json.JSONDecodeError:
                data = {}
        node = data.setdefault(self.data_name, {}).setdefault(res_name, {})
        rows = []
        print(f"\nHiMo refinement metrics for {res_name} in {self.data_name}:")
        for c, disp in (("CAR", "CAR"), ("OTHER_VEHICLES", "OTHERS")):
            if c not in s:
                continue
            node[c] = s[c]
            o = s[c]["overall"]
            rows.append([disp, f"{o['cd']:.3f} ± {o['std_cd']:.2f}", f"{o['mpe']:.3f} ± {o['std_mpe']:.2f}",
                         o["num_pts"], o["num_obj"]])
        if "Total" in s:
            t = s["Total"]
            rows.insert(0, ["Total", f"{t['cd']:.3f}", f"{t['mpe']:.3f}", t["num_pts"], t["num_obj"]])
        json.dump(data, open(file_name, "w"), indent=4)
        print(tabulate(rows, headers=["Class", "CDE (Chamfer) ↓", "MPE (Point Err) ↓", "# Points", "# Objs"],
                       tablefmt="fancy_grid", stralign="center"))
        print(f"Total frames processed: {self.frame_cnt}")
        print(f"Results saved to {file_name}\n")
        return s


def eval_masks(data: Dict, data_name: str) -> np.ndarray:
    """eval.py:288-296: <= 35 m in xy, not ground, outside the ego box (+ flow_is_valid for Scania)."""
    pc0 = data["pc0"]
    m = (np.linalg.norm(pc0[:, :2], axis=1) <= CLOSE_DISTANCE_THRESHOLD) & ~data["gm0"]
    if data_name == "scania":
        return m & data["flow_is_valid"] & ego_pts_mask(pc0)
    return m & ego_pts_mask(pc0, min_bound=[-1.5, -1.5, -2.0], max_bound=[1.5, 1.5, 2.0])
