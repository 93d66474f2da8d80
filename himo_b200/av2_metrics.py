"""OpenSceneFlow / AV2 scene-flow evaluation metrics, host side like the reference.

Work-alike of OSF/src/utils/eval_metric.py (`evaluate_leaderboard`, `evaluate_leaderboard_v2`, `evaluate_ssf`,
`OfficialMetrics`) and of the metric kernels of OSF/src/utils/av2_eval.py (`compute_metrics` :460-552,
`compute_bucketed_epe` :839-870, `compute_ssf_metrics` :872-908).  The reference imports `av2` and `rich` at module
level, neither of which is needed for the arithmetic; the category table is the one HiMo's scorer carries
(tools/test/score.py:29-94, shared with himo_b200.himo).  Inputs may be numpy arrays or torch tensors (any device);
everything is evaluated in float64 on the host exactly as the reference does after its `.cpu().numpy().astype(float)`.

tests/test_av2_metrics.py pins every function against the reference's own module (imported through shims of `av2` /
`rich`) on seeded random frames, and against committed golden values.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .himo import CATEGORY_TO_INDEX

# av2_eval.py:32-75 (HiMo's own scorer only keeps the two vehicle groups; OpenSceneFlow evaluates all five)
BUCKETED_METACATAGORIES = {
    "BACKGROUND": ["NONE"],
    "CAR": ["REGULAR_VEHICLE"],
    "PEDESTRIAN": ["PEDESTRIAN", "STROLLER", "WHEELCHAIR", "OFFICIAL_SIGNALER"],
    "WHEELED_VRU": ["BICYCLE", "BICYCLIST", "MOTORCYCLE", "MOTORCYCLIST", "WHEELED_DEVICE", "WHEELED_RIDER"],
    "OTHER_VEHICLES": ["BOX_TRUCK", "LARGE_VEHICLE", "RAILED_VEHICLE", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER",
                       "ARTICULATED_BUS", "BUS", "SCHOOL_BUS"],
}
CLOSE_DISTANCE_THRESHOLD = 35.0          # av2_eval.py:30
SCENE_FLOW_DYNAMIC_THRESHOLD = 0.05      # av2_eval.py:28
_EPS_IOU = 1e-6                          # av2_eval.py:449 (the later of the two EPS definitions)

# av2_eval.py:155-232: foreground = every category of the four enums, in enum order
_FOREGROUND = [
    "BOLLARD", "CONSTRUCTION_BARREL", "CONSTRUCTION_CONE", "MOBILE_PEDESTRIAN_CROSSING_SIGN", "SIGN", "STOP_SIGN",
    "ANIMAL", "DOG", "OFFICIAL_SIGNALER", "PEDESTRIAN",
    "BICYCLE", "BICYCLIST", "MOTORCYCLE", "MOTORCYCLIST", "STROLLER", "WHEELCHAIR", "WHEELED_DEVICE", "WHEELED_RIDER",
    "ARTICULATED_BUS", "BOX_TRUCK", "BUS", "LARGE_VEHICLE", "MESSAGE_BOARD_TRAILER", "RAILED_VEHICLE",
    "REGULAR_VEHICLE", "SCHOOL_BUS", "TRAFFIC_LIGHT_TRAILER", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER",
]
FOREGROUND_BACKGROUND_BREAKDOWN = {"Background": [0], "Foreground": [CATEGORY_TO_INDEX[c] for c in _FOREGROUND]}


def _np(a, dtype=None):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.asarray(a)
    return a.astype(dtype) if dtype is not None else a


@dataclass(frozen=True, eq=True, repr=True)
class BaseSplitValue:                     # av2_eval.py:828-837
    name: str
    avg_epe: float
    avg_range: float
    thresholds_range: Tuple[float, float]
    count: int


# ------------------------------------------------------------------------------------------ metric kernels
def compute_metrics(pred_flow, pred_dynamic, gts, category_indices, is_dynamic, is_close, is_valid) -> Dict[str, float]:
    """EPE three-way + dynamic IoU of one frame (av2_eval.py:460-552)."""
    is_valid = _np(is_valid, bool)
    pred_flow = _np(pred_flow, np.float64)[is_valid]
    pred_dynamic = _np(pred_dynamic, bool)[is_valid]
    gts = _np(gts, np.float64)[is_valid]
    category_indices = _np(category_indices).astype(int)[is_valid]
    is_dynamic = _np(is_dynamic, bool)[is_valid]
    is_close = _np(is_close, bool)[is_valid]
    counts, epes = [], []
    tp = fp = fn = 0
    for _cls, idxs in FOREGROUND_BACKGROUND_BREAKDOWN.items():
        cat = np.isin(category_indices, np.asarray(idxs))
        for m_mask in (is_dynamic, ~is_dynamic):
            for d_mask in (is_close, ~is_close):
                mask = cat & m_mask & d_mask
                n = int(mask.sum())
                counts.append(n)
                if n > 0:
                    epes.append(np.linalg.norm(pred_flow[mask] - gts[mask], axis=-1).astype(np.float64).mean())
                    pd_, gd_ = pred_dynamic[mask], is_dynamic[mask]
                    tp += int(np.logical_and(pd_, gd_).sum())
                    fp += int(np.logical_and(pd_, ~gd_).sum())
                    fn += int(np.logical_and(~pd_, gd_).sum())
                else:
                    epes.append(np.nan)

    def epe(indices, eps=1e-8):            # av2_eval.py:450-458
        s, c = 0.0, 0
        for i in indices:
            if counts[i] != 0:
                s += epes[i] * counts[i]
                c += counts[i]
        return s / (c + eps) if c != 0 else 0.0

    # rows: Background {Dynamic,Static} x {Close,Far}, then Foreground likewise
    return {"EPE_BS": epe([2, 3]), "EPE_FD": epe([4, 5]), "EPE_FS": epe([6, 7]), "IoU": tp / (tp + fp + fn + _EPS_IOU)}


def speed_thresholds() -> List[Tuple[float, float]]:
    s = np.concatenate([np.linspace(0, 2.0, 51), [np.inf]])
    return list(zip(s, s[1:]))


def compute_bucketed_epe(pred_flow, gt_flow, category_indices, is_valid) -> List[BaseSplitValue]:
    """Per meta-category, per GT-speed bucket mean EPE of one frame (av2_eval.py:839-870)."""
    pred_flow, gt_flow = _np(pred_flow, np.float64), _np(gt_flow, np.float64)
    category_indices, is_valid = _np(category_indices), _np(is_valid, bool)
    out: List[BaseSplitValue] = []
    gt_speeds = np.linalg.norm(gt_flow, axis=-1)
    err = np.linalg.norm(pred_flow - gt_flow, axis=-1)
    for cname, members in BUCKETED_METACATAGORIES.items():
        cat = np.isin(category_indices, np.array([CATEGORY_TO_INDEX[c] for c in members]))
        if cname == "BACKGROUND":
            mask = cat & is_valid
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                out.append(BaseSplitValue(cname, err[mask].mean(), gt_speeds[mask].mean(), (0.0, 0.04), mask.sum()))
            continue
        for lo, hi in speed_thresholds():
            mask = cat & (gt_speeds >= lo) & (gt_speeds < hi) & is_valid
            n = mask.sum()
            if n == 0:
                continue
            out.append(BaseSplitValue(cname, err[mask].mean(), gt_speeds[mask].mean(), (lo, hi), n))
    return out


DISTANCE_SPLIT = [0, 35, 50, 75, 100, np.inf]


def compute_ssf_metrics(pc0_dis, pred_flow, gt_flow, is_valid, dynamic_speed: float = 1.4) -> List[BaseSplitValue]:
    """Range-wise static / dynamic EPE of one frame (av2_eval.py:872-908)."""
    pc0_dis, pred_flow, gt_flow = _np(pc0_dis, np.float64), _np(pred_flow, np.float64), _np(gt_flow, np.float64)
    is_valid = _np(is_valid, bool)
    out: List[BaseSplitValue] = []
    gt_speeds = np.linalg.norm(gt_flow, axis=-1) * 10
    for lo, hi in zip(DISTANCE_SPLIT, DISTANCE_SPLIT[1:]):
        mask = (pc0_dis >= lo) & (pc0_dis < hi) & is_valid
        sp, pf, gf, dis = gt_speeds[mask], pred_flow[mask], gt_flow[mask], pc0_dis[mask]
        dyn = sp >= dynamic_speed
        for motion, m in (("Dynamic", dyn), ("Static", ~dyn)):
            n = m.sum()
            if n == 0:
                continue
            out.append(BaseSplitValue(motion, np.linalg.norm(pf - gf, axis=-1)[m].mean(), dis[m].mean(), (lo, hi), n))
    return out


# ------------------------------------------------------------------------------------------ per-frame front ends
def _norm_like_reference(a: np.ndarray) -> np.ndarray:
    """Row norms evaluated by torch in the input precision -- the reference takes them with
    torch.linalg.vector_norm on the original (float32) tensors (eval_metric.py:29,34,48,63,85), and thresholds /
    bucket means depend on those roundings."""
    import torch
    return torch.linalg.vector_norm(torch.from_numpy(np.ascontiguousarray(a)), dim=-1).numpy()


def _finite_rows(*arrs):
    m = np.ones(arrs[0].shape[0], bool)
    for a in arrs:
        m &= ~np.isnan(a).any(axis=1)
    return m


def evaluate_leaderboard(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> Dict[str, float]:
    """eval_metric.py:28-58 (the norms are taken in the input precision, as torch does on the original tensors)."""
    est_flow, rigid_flow, pc0, gt_flow = _np(est_flow), _np(rigid_flow), _np(pc0), _np(gt_flow)
    is_valid, pts_ids = _np(is_valid), _np(pts_ids)
    gt_dyn = _norm_like_reference(gt_flow - rigid_flow) >= SCENE_FLOW_DYNAMIC_THRESHOLD
    m = _finite_rows(est_flow, rigid_flow, pc0[:, :3], gt_flow)
    m &= ~np.isnan(is_valid.astype(np.float64)) & ~np.isnan(pts_ids.astype(np.float64))
    m &= _norm_like_reference(pc0[:, :2]) <= 35.0
    est_flow, rigid_flow, pc0, gt_flow = est_flow[m], rigid_flow[m], pc0[m], gt_flow[m]
    est_dyn = _norm_like_reference(est_flow - rigid_flow) >= SCENE_FLOW_DYNAMIC_THRESHOLD
    is_close = np.all(np.abs(pc0[:, :2]) <= CLOSE_DISTANCE_THRESHOLD, axis=1)
    return compute_metrics(est_flow.astype(float), est_dyn, gt_flow.astype(float), pts_ids[m].astype(np.uint8),
                           gt_dyn[m], is_close, is_valid[m].astype(bool))


def evaluate_leaderboard_v2(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> List[BaseSplitValue]:
    """eval_metric.py:61-80: ego motion removed from both flows, points within 35 m (xy)."""
    est_flow, rigid_flow, pc0, gt_flow = _np(est_flow), _np(rigid_flow), _np(pc0), _np(gt_flow)
    is_valid, pts_ids = _np(is_valid), _np(pts_ids)
    m = _finite_rows(est_flow, rigid_flow, pc0[:, :3], gt_flow)
    m &= ~np.isnan(is_valid.astype(np.float64)) & ~np.isnan(pts_ids.astype(np.float64))
    m &= _norm_like_reference(pc0[:, :2]) <= CLOSE_DISTANCE_THRESHOLD
    rf = rigid_flow[m]
    return compute_bucketed_epe((est_flow[m] - rf).astype(float), (gt_flow[m] - rf).astype(float),
                                pts_ids[m].astype(np.uint8), is_valid[m].astype(bool))


def evaluate_ssf(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> List[BaseSplitValue]:
    """eval_metric.py:83-108: range-wise (3-D distance) static / dynamic EPE, ego motion removed."""
    est_flow, rigid_flow, pc0, gt_flow = _np(est_flow), _np(rigid_flow), _np(pc0), _np(gt_flow)
    is_valid, pts_ids = _np(is_valid), _np(pts_ids)
    dis = _norm_like_reference(pc0[:, :3])
    m = _finite_rows(est_flow, rigid_flow, pc0[:, :3], gt_flow)
    m &= ~np.isnan(is_valid.astype(np.float64)) & ~np.isnan(pts_ids.astype(np.float64))
    rf = rigid_flow[m]
    return compute_ssf_metrics(dis[m].astype(float), (est_flow[m] - rf).astype(float), (gt_flow[m] - rf).astype(float),
                               is_valid[m].astype(bool))


# ------------------------------------------------------------------------------------------ accumulation over frames
class BucketResultMatrix:
    """eval_metric.py:130-199: count-weighted running means per (class, bucket)."""

    def __init__(self, class_names: List[str], range_buckets: List[Tuple[float, float]]):
        assert class_names and range_buckets
        self.class_names, self.range_buckets = list(class_names), list(range_buckets)
        shape = (len(class_names), len(range_buckets))
        self.epe_storage_matrix = np.full(shape, np.nan)
        self.range_storage_matrix = np.full(shape, np.nan)
        self.count_storage_matrix = np.zeros(shape, dtype=np.int64)

    def accumulate_value(self, class_name, range_bucket, average_epe, average_range, count):
        if count == 0 or np.isnan(average_epe) or np.isnan(average_range):
            return
        ci, bi = self.class_names.index(class_name), self.range_buckets.index(range_bucket)
        pe, pr, pc = self.epe_storage_matrix[ci, bi], self.range_storage_matrix[ci, bi], self.count_storage_matrix[ci, bi]
        if np.isnan(pe):
            self.epe_storage_matrix[ci, bi], self.range_storage_matrix[ci, bi] = average_epe, average_range
            self.count_storage_matrix[ci, bi] = count
            return
        self.epe_storage_matrix[ci, bi] = np.average([pe, average_epe], weights=[pc, count])
        self.range_storage_matrix[ci, bi] = np.average([pr, average_range], weights=[pc, count])
        self.count_storage_matrix[ci, bi] += count

    def get_class_entries(self, class_name):
        ci = self.class_names.index(class_name)
        return self.epe_storage_matrix[ci], self.range_storage_matrix[ci], self.count_storage_matrix[ci]

    def merge(self, other: "BucketResultMatrix") -> None:
        """Fold another rank's matrix in (what OSF/src/runner.py:262-289 does after gather_object)."""
        for ci, cname in enumerate(other.class_names):
            for bi, bucket in enumerate(other.range_buckets):
                self.accumulate_value(cname, bucket, other.epe_storage_matrix[ci, bi], other.range_storage_matrix[ci, bi],
                                      int(other.count_storage_matrix[ci, bi]))


class BucketedSpeedMatrix(BucketResultMatrix):
    """eval_metric.py:203-238: speed-normalised dynamic error."""

    def get_normalized_error_matrix(self):
        e = self.epe_storage_matrix.copy()
        e[:, 1:] = e[:, 1:] / self.range_storage_matrix[:, 1:]
        return e

    def get_overall_class_errors(self, normalized: bool = True) -> Dict[str, Tuple[float, float]]:
        e = self.get_normalized_error_matrix() if normalized else self.epe_storage_matrix.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            dyn = np.nanmean(e[:, 1:], axis=1)
        return {c: (float(s), float(d)) for c, s, d in zip(self.class_names, e[:, 0], dyn)}

    def get_mean_average_values(self, normalized: bool = True) -> Tuple[float, float]:
        o = self.get_overall_class_errors(normalized)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return float(np.nanmean([v[0] for v in o.values()])), float(np.nanmean([v[1] for v in o.values()]))


class OfficialMetrics:
    """eval_metric.py:240-365: frame-wise accumulation, `normalize()`, `print()`."""

    CLASSES = ["BACKGROUND", "CAR", "OTHER_VEHICLES", "PEDESTRIAN", "WHEELED_VRU"]

    def __init__(self):
        self.bucketed = {k: {"Static": [], "Dynamic": []} for k in self.CLASSES + ["Mean"]}
        self.epe_3way = {"EPE_FD": [], "EPE_BS": [], "EPE_FS": [], "IoU": [], "Three-way": []}
        self.epe_ssf: Dict[str, Dict] = {}
        self.norm_flag = False
        self.bucketedMatrix = BucketedSpeedMatrix(self.CLASSES, speed_thresholds())
        buckets = list(zip(DISTANCE_SPLIT, DISTANCE_SPLIT[1:]))
        self.distanceMatrix = BucketResultMatrix(["Static", "Dynamic"], buckets)
        for lo, hi in buckets:
            name = f"{int(lo)}-{int(hi)}" if hi != np.inf else f"{int(lo)}-inf"
            self.epe_ssf[name] = {"Static": [], "Dynamic": [], "#Static": 0, "#Dynamic": 0}

    def step(self, epe_dict, bucket_dict, ssf_dict=None):
        for k in epe_dict:
            self.epe_3way[k].append(epe_dict[k])
        for it in bucket_dict:
            self.bucketedMatrix.accumulate_value(it.name, it.thresholds_range, it.avg_epe, it.avg_range, it.count)
        if ssf_dict is not None:
            for it in ssf_dict:
                self.distanceMatrix.accumulate_value(it.name, it.thresholds_range, it.avg_epe, it.avg_range, it.count)

    def merge(self, other: "OfficialMetrics") -> None:
        """Combine the un-normalised state of another rank (scene-sharded evaluation)."""
        for k in self.epe_3way:
            self.epe_3way[k].extend(other.epe_3way[k])
        self.bucketedMatrix.merge(other.bucketedMatrix)
        self.distanceMatrix.merge(other.distanceMatrix)

    def normalize(self):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            for k in self.epe_3way:
                self.epe_3way[k] = np.mean(self.epe_3way[k])
            self.epe_3way["Three-way"] = np.mean([self.epe_3way["EPE_FD"], self.epe_3way["EPE_BS"], self.epe_3way["EPE_FS"]])
        mean = self.bucketedMatrix.get_mean_average_values(True)
        cls_err = self.bucketedMatrix.get_overall_class_errors(True)
        for k in self.bucketed:
            src = mean if k == "Mean" else cls_err[k]
            self.bucketed[k]["Static"], self.bucketed[k]["Dynamic"] = src[0], src[1]
        self.norm_flag = True
        self.epe_ssf["Mean"] = {"Static": [], "Dynamic": [], "#Static": np.nan, "#Dynamic": np.nan}
        for motion in ("Static", "Dynamic"):
            epes, diss, cnts = self.distanceMatrix.get_class_entries(motion)
            for e, d, c in zip(epes, diss, cnts):
                for key in self.epe_ssf:
                    if key == "Mean":
                        continue
                    lo, hi = key.split("-")
                    lo, hi = int(lo), (int(hi) if hi != "inf" else np.inf)
                    if hi > d >= lo:
                        self.epe_ssf[key][motion] = e
                        self.epe_ssf[key]["#" + motion] += c
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                self.epe_ssf["Mean"][motion] = np.nanmean(epes)

    def print(self, ssf_metrics: bool = False):
        from tabulate import tabulate
        if not self.norm_flag:
            self.normalize()
        print("Version 1 Metric on EPE Three-way:")
        print(tabulate([[k, v] for k, v in self.epe_3way.items()]), "\n")
        print("Version 2 Metric on Normalized Category-based:")
        print(tabulate([[k, v["Static"], v["Dynamic"]] for k, v in self.bucketed.items()],
                       headers=["Class", "Static", "Dynamic"], tablefmt="orgtbl"), "\n")
        if ssf_metrics:
            rows = [[k, np.around(v["Static"], 4), np.around(v["Dynamic"], 4), v["#Static"], v["#Dynamic"]]
                    for k, v in self.epe_ssf.items()]
            print("Version 3 Metric on EPE Distance-based:")
            print(tabulate(rows, headers=["Distance", "Static", "Dynamic", "#Static", "#Dynamic"], tablefmt="orgtbl"), "\n")
