"""Seeded synthetic frame generators: the inputs of every parity test
and of bench.py.  Host-side numpy only; no dataset or network access is needed.

  uniform_frame   G-uniform: xyz ~ U([-60,60]^2 x [-2.5,2.5]); ~27 % of the points fall outside
                  the SeFlow++ range and exercise the -1 paths, ~63 k pillars per 73 k in-range pts
  LidarWorld      G-lidar ("Scania/AV2-shaped"): K spinning lidars ray-cast against a static
                  scene of building / wall boxes plus moving vehicle boxes; ground removed by
                  construction; per-point lidar_dt from the azimuth; rigid ego motion; GT flow,
                  instance ids and categories for the HiMo metrics
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

POINT_CLOUD_RANGE = [-51.2, -51.2, -3.0, 51.2, 51.2, 3.0]
VOXEL_SIZE = [0.2, 0.2, 6.0]


def uniform_frame(n: int, seed: int, nan_rows: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    pts = np.empty((n, 3), np.float32)
    pts[:, 0] = rng.uniform(-60, 60, n)
    pts[:, 1] = rng.uniform(-60, 60, n)
    pts[:, 2] = rng.uniform(-2.5, 2.5, n)
    if nan_rows:
        pts[rng.choice(n, nan_rows, replace=False)] = np.nan
    return pts


def pose_matrix(x: float, y: float, yaw: float, z: float = 0.0) -> np.ndarray:
    c, s = np.cos(yaw), np.sin(yaw)
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[:3, 3] = [x, y, z]
    return T


@dataclass
class Box:
    center: np.ndarray      # world xyz at t = 0
    size: np.ndarray        # lx, ly, lz
    yaw: float
    velocity: np.ndarray    # world m/s (xy), zero for static structure
    instance: int = 0
    category: int = 0       # AV2 category index (tools/test/score.py:29-64); 19 = REGULAR_VEHICLE


@dataclass
class LidarWorld:
    """A small deterministic world that can be observed at any time t (seconds)."""
    seed: int = 0
    n_lidars: int = 6               # Scania trucks carry 6 (dataprocess/extract_sca.py:170); AV2 has 2
    n_rings: int = 64
    n_azimuth: int = 2000
    n_vehicles: int = 32
    n_structures: int = 60
    n_clutter: int = 250
    ego_speed: float = 12.0         # m/s along +x
    clutter_fraction: float = 0.35  # share of low, static terrain-residual points (kerbs, grass)
    sweep_period: float = 0.1
    boxes: List[Box] = field(default_factory=list)

    def __post_init__(self):
        rng = np.random.default_rng(self.seed)
        boxes: List[Box] = []
        for i in range(self.n_structures):            # buildings / walls, static, tall
            r = rng.uniform(12, 70)
            a = rng.uniform(0, 2 * np.pi)
            size = np.array([rng.uniform(6, 30), rng.uniform(0.5, 12), rng.uniform(3, 9)])
            boxes.append(Box(np.array([r * np.cos(a), r * np.sin(a), size[2] / 2 - 1.7]), size,
                             rng.uniform(0, np.pi), np.zeros(2), 0, 0))
        for i in range(self.n_clutter):               # poles / bushes / signs: small static boxes
            r = rng.uniform(4, 60)
            a = rng.uniform(0, 2 * np.pi)
            size = np.array([rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5), rng.uniform(0.5, 4.0)])
            boxes.append(Box(np.array([r * np.cos(a), r * np.sin(a), size[2] / 2 - 1.7]), size,
                             rng.uniform(0, np.pi), np.zeros(2), 0, 0))
        cats = [19, 19, 19, 6, 7, 25]                  # CAR-heavy mix; others = OTHER_VEHICLES
        for i in range(self.n_vehicles):
            r = rng.uniform(5, 45)
            a = rng.uniform(0, 2 * np.pi)
            cat = cats[int(rng.integers(len(cats)))]
            size = np.array([4.5, 1.9, 1.6]) if cat == 19 else np.array([9.0, 2.5, 3.2])
            yaw = rng.uniform(0, 2 * np.pi)
            speed = rng.uniform(0, 35) if rng.random() < 0.7 else 0.0
            vel = speed * np.array([np.cos(yaw), np.sin(yaw)])
            boxes.append(Box(np.array([r * np.cos(a), r * np.sin(a), size[2] / 2 - 1.7]), size, yaw,
                             vel, i + 1, cat))
        self.boxes = boxes
        self._lidar_offsets = np.array(
            [[3.5, 0.0, 1.2], [3.5, 1.2, 0.4], [3.5, -1.2, 0.4], [-4.0, 0.0, 1.0], [0.0, 1.3, 0.8],
             [0.0, -1.3, 0.8]][: self.n_lidars], np.float64)
        self._elev = np.deg2rad(np.linspace(-24.0, 6.0, self.n_rings))

    # ------------------------------------------------------------------ observation
    def ego_pose(self, t: float) -> np.ndarray:
        return pose_matrix(self.ego_speed * t, 0.35 * np.sin(0.4 * t), 0.03 * np.sin(0.7 * t))

    def observe(self, t: float, n_points: int, seed: Optional[int] = None,
                fp16_quantise: bool = False) -> Dict[str, np.ndarray]:
        """One sweep starting at time t -> dict with the fields of the reference .h5 frame groups
        (OSF/dataprocess/extract_av2.py:225-238): lidar [N,4], ground_mask, pose, lidar_dt, lidar_id,
        flow (total flow incl. ego motion, to t + period), flow_is_valid, flow_category_indices,
        flow_instance_id, ego_motion."""
        rng = np.random.default_rng(self.seed * 1000003 + int(round(t * 1000)) if seed is None else seed)
        az = np.linspace(0, 2 * np.pi, self.n_azimuth, endpoint=False)
        dt_of_az = (az / (2 * np.pi) * self.sweep_period).astype(np.float32)
        pose0 = self.ego_pose(t)
        pose1 = self.ego_pose(t + self.sweep_period)
        R0, t0 = pose0[:3, :3], pose0[:3, 3]
        pts_all, dt_all, id_all, inst_all, cat_all, flow_all = [], [], [], [], [], []
        ce, se = np.cos(self._elev), np.sin(self._elev)
        dirs_l = np.stack([np.outer(ce, np.cos(az)), np.outer(ce, np.sin(az)),
                           np.outer(se, np.ones_like(az))], -1).reshape(-1, 3)     # ego frame
        dts = np.tile(dt_of_az, self.n_rings)
        dirs_w = dirs_l @ R0.T
        for li in range(self.n_lidars):
            org_w = R0 @ self._lidar_offsets[li] + t0
            n_az = self.n_azimuth
            best_t = np.full((self.n_rings, n_az), np.inf)
            best_b = np.full((self.n_rings, n_az), -1)
            dirs_w3 = dirs_w.reshape(self.n_rings, n_az, 3)
            for bi, b in enumerate(self.boxes):
                c = b.center.copy()
                c[:2] += b.velocity * t                     # box pose frozen at sweep start
                cy, sy = np.cos(b.yaw), np.sin(b.yaw)
                Rb = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
                # azimuth window of the box seen from this lidar (ego frame) -> candidate columns
                corners = (np.array([[sx, sy_, 0.0] for sx in (-0.5, 0.5) for sy_ in (-0.5, 0.5)])
                           * b.size) @ Rb.T + c
                ce_ = (corners - org_w) @ R0
                if np.linalg.norm((c - org_w)[:2]) < 0.75 * np.linalg.norm(b.size[:2]):
                    cols = np.arange(n_az)
                else:
                    a = np.arctan2(ce_[:, 1], ce_[:, 0])
                    a0 = np.arctan2((c - org_w) @ R0[:, 1], (c - org_w) @ R0[:, 0])
                    rel = (a - a0 + np.pi) % (2 * np.pi) - np.pi
                    lo = int(np.floor((a0 + rel.min()) / (2 * np.pi) * n_az)) - 1
                    hi = int(np.ceil((a0 + rel.max()) / (2 * np.pi) * n_az)) + 1
                    cols = np.arange(lo, hi + 1) % n_az
                d = dirs_w3[:, cols].reshape(-1, 3) @ Rb
                o = (org_w - c) @ Rb                         # into box frame
                with np.errstate(divide="ignore", invalid="ignore"):
                    inv = 1.0 / d
                    t1 = (-b.size / 2 - o) * inv
                    t2 = (b.size / 2 - o) * inv
                tn = np.nanmax(np.minimum(t1, t2), axis=1).reshape(self.n_rings, -1)
                tf = np.nanmin(np.maximum(t1, t2), axis=1).reshape(self.n_rings, -1)
                cur = best_t[:, cols]
                hit = (tf >= tn) & (tn > 0.5) & (tn < cur) & (tn < 120.0)
                cur[hit] = tn[hit]
                best_t[:, cols] = cur
                curb = best_b[:, cols]
                curb[hit] = bi
                best_b[:, cols] = curb
            best_t = best_t.reshape(-1)
            best_b = best_b.reshape(-1)
            ok = best_b >= 0
            p_w = org_w + dirs_w[ok] * best_t[ok, None]
            p_w += rng.normal(0, 0.01, p_w.shape)
            bsel = best_b[ok]
            vel = np.array([self.boxes[k].velocity for k in bsel]).reshape(-1, 2)
            inst = np.array([self.boxes[k].instance for k in bsel])
            cat = np.array([self.boxes[k].category for k in bsel])
            # ego frame of the sweep; total flow to the next sweep = object motion + ego motion
            p_e = (p_w - t0) @ R0
            p_w1 = p_w.copy()
            p_w1[:, :2] += vel * self.sweep_period
            p_e1 = (p_w1 - pose1[:3, 3]) @ pose1[:3, :3]
            pts_all.append(p_e)
            flow_all.append(p_e1 - p_e)
            dt_all.append(dts[ok])
            id_all.append(np.full(p_e.shape[0], li))
            inst_all.append(inst)
            cat_all.append(cat)
        if self.clutter_fraction > 0:
            # terrain residue that ground segmentation leaves behind: low static points whose
            # density falls off with range like lidar ground rings do
            n_hit = sum(p.shape[0] for p in pts_all)
            n_cl = int(n_hit * self.clutter_fraction / max(1e-6, 1 - self.clutter_fraction))
            a = rng.uniform(0, 2 * np.pi, n_cl)
            r = 3.0 + 55.0 * rng.uniform(0, 1, n_cl) ** 1.5
            p_e = np.stack([r * np.cos(a), r * np.sin(a),
                            -1.65 + 0.35 * np.abs(rng.normal(0, 1, n_cl))], 1)
            p_w = p_e @ R0.T + t0
            p_e1 = (p_w - pose1[:3, 3]) @ pose1[:3, :3]
            pts_all.append(p_e)
            flow_all.append(p_e1 - p_e)
            dt_all.append((a / (2 * np.pi) * self.sweep_period).astype(np.float32))
            id_all.append(rng.integers(0, self.n_lidars, n_cl))
            inst_all.append(np.zeros(n_cl, np.int64))
            cat_all.append(np.zeros(n_cl, np.int64))
        pts = np.concatenate(pts_all)
        flow = np.concatenate(flow_all)
        dt = np.concatenate(dt_all)
        lid = np.concatenate(id_all)
        inst = np.concatenate(inst_all)
        cat = np.concatenate(cat_all)
        m = pts.shape[0]
        if m == 0:
            raise RuntimeError("empty synthetic sweep")
        sel = rng.choice(m, n_points, replace=m < n_points)
        sel.sort()
        pts, flow, dt, lid, inst, cat = pts[sel], flow[sel], dt[sel], lid[sel], inst[sel], cat[sel]
        if m < n_points:  # de-duplicate resampled points with a little range noise
            pts = pts + rng.normal(0, 0.02, pts.shape)
        pts32 = pts.astype(np.float16).astype(np.float32) if fp16_quantise else pts.astype(np.float32)
        lidar = np.concatenate([pts32, rng.uniform(0, 1, (n_points, 1)).astype(np.float32)], 1)
        ego_motion = (np.linalg.inv(pose1) @ pose0).astype(np.float32)
        return {
            "lidar": lidar,
            "ground_mask": np.zeros(n_points, bool),
            "pose": pose0.astype(np.float32),
            "lidar_dt": dt.astype(np.float32),
            "lidar_id": lid.astype(np.uint8),
            "flow": flow.astype(np.float32),
            "flow_is_valid": np.ones(n_points, bool),
            "flow_category_indices": cat.astype(np.uint8),
            "flow_instance_id": inst.astype(np.int16),
            "ego_motion": ego_motion,
        }


def world_for_points(n_points: int, seed: int, **kw) -> LidarWorld:
    """LidarWorld whose ray budget is scaled to the requested sweep size (keeps tests fast)."""
    if "n_rings" not in kw and "n_azimuth" not in kw:
        if n_points <= 4000:
            kw.update(n_rings=16, n_azimuth=360)
        elif n_points <= 20000:
            kw.update(n_rings=32, n_azimuth=720)
    return LidarWorld(seed=seed, **kw)


def lidar_triple(n_points: int, seed: int, t: float = 1.0, **world_kw) -> Dict[str, np.ndarray]:
    """A (t-1, t, t+1) frame triple in the layout DeFlowPP.forward consumes after ground removal
    (OSF/src/trainer.py:290-297): pch1/pc0/pc1 [N,3] f32 and poseh1/pose0/pose1 [4,4] f32."""
    w = world_for_points(n_points, seed, **world_kw)
    fh = w.observe(t - w.sweep_period, n_points)
    f0 = w.observe(t, n_points)
    f1 = w.observe(t + w.sweep_period, n_points)
    return {
        "pch1": fh["lidar"][:, :3].copy(), "pc0": f0["lidar"][:, :3].copy(),
        "pc1": f1["lidar"][:, :3].copy(),
        "poseh1": fh["pose"], "pose0": f0["pose"], "pose1": f1["pose"],
        "frames": (fh, f0, f1),
    }


def uniform_triple(n_points: int, seed: int) -> Dict[str, np.ndarray]:
    """G-uniform triple with a small rigid ego step between frames."""
    return {
        "pch1": uniform_frame(n_points, seed * 3 + 0), "pc0": uniform_frame(n_points, seed * 3 + 1),
        "pc1": uniform_frame(n_points, seed * 3 + 2),
        "poseh1": pose_matrix(-1.2, 0.02, -0.004).astype(np.float32),
        "pose0": pose_matrix(0.0, 0.0, 0.0).astype(np.float32),
        "pose1": pose_matrix(1.2, 0.03, 0.005).astype(np.float32),
    }
