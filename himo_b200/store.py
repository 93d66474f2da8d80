"""Frame storage either side of the hot path: the reference's per-scene `.h5` files
(`<dir>/<scene_id>.h5` -> group `str(timestamp)` -> datasets `lidar`, `ground_mask`, `pose`, `lidar_dt`,
`flow`, ..., and the result dataset `<res_name>`; OSF/dataprocess/extract_av2.py:225-238,
OSF/src/trainer.py:337-343) plus `index_total.pkl` / `index_eval.pkl` (OSF/dataprocess/misc_data.py:32-55).

Two backends behind one interface:
  * H5Store    -- the real format, through h5py when it is importable (it is not in the build image;
                  there is no libhdf5 here), semantics of the reference: open 'r+' per write,
                  `del f[key][name]` then `create_dataset` (idempotent re-runs).
  * NpyStore   -- `<dir>/<scene_id>.frames/<timestamp>/<name>.npy`; same keys, same dtypes; used by the
                  tests, the synthetic benchmarks and anywhere h5py is absent.
`open_store(dir)` picks H5Store when the directory holds `.h5` files and h5py imports, else NpyStore.
"""
from __future__ import annotations

import os
import pickle
import threading
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_H5_WRITE_LOCK = threading.Lock()


class FrameStore:
    def scenes(self) -> List[str]:
        raise NotImplementedError

    def has(self, scene: str, ts, name: str) -> bool:
        raise NotImplementedError

    def read(self, scene: str, ts, name: str) -> np.ndarray:
        raise NotImplementedError

    def write(self, scene: str, ts, name: str, data: np.ndarray) -> None:
        raise NotImplementedError

    def timestamps(self, scene: str) -> List[str]:
        raise NotImplementedError

    def names(self, scene: str, ts) -> List[str]:
        raise NotImplementedError


class NpyStore(FrameStore):
    def __init__(self, directory: str):
        self.dir = directory

    def _frame(self, scene, ts):
        return os.path.join(self.dir, f"{scene}.frames", str(ts))

    def scenes(self):
        return sorted(d[:-7] for d in os.listdir(self.dir) if d.endswith(".frames"))

    def timestamps(self, scene):
        return sorted(os.listdir(os.path.join(self.dir, f"{scene}.frames")), key=int)

    def has(self, scene, ts, name):
        return os.path.exists(os.path.join(self._frame(scene, ts), name + ".npy"))

    def names(self, scene, ts):
        return sorted(f[:-4] for f in os.listdir(self._frame(scene, ts)) if f.endswith(".npy"))

    def read(self, scene, ts, name):
        return np.load(os.path.join(self._frame(scene, ts), name + ".npy"))

    def write(self, scene, ts, name, data):
        d = self._frame(scene, ts)
        os.makedirs(d, exist_ok=True)
        tmp = os.path.join(d, f".{name}.{os.getpid()}.tmp.npy")
        np.save(tmp, np.asarray(data))
        os.replace(tmp, os.path.join(d, name + ".npy"))     # atomic: a re-run simply replaces the result


def _h5_backend():
    """h5py when it is importable (the reference's own library), else the self-contained subset codec `h5lite`
    (superblock v0, symbol-table groups, contiguous datasets: what h5py writes by default and what the pipeline uses)."""
    try:
        import h5py
        return h5py
    except ImportError:
        from . import h5lite
        return h5lite


class H5Store(FrameStore):
    """`<scene_id>.h5` files in the reference's schema: group `str(timestamp)` -> datasets (OSF/src/dataset.py:313-364 reads,
    OSF/src/trainer.py:337-343 writes)."""

    def __init__(self, directory: str, backend=None):
        self.dir = directory
        self.h5 = backend or _h5_backend()

    def _path(self, scene):
        return os.path.join(self.dir, f"{scene}.h5")

    def scenes(self):
        return sorted(f[:-3] for f in os.listdir(self.dir) if f.endswith(".h5"))

    def timestamps(self, scene):
        with self.h5.File(self._path(scene), "r") as f:
            return sorted(f.keys(), key=int)

    def has(self, scene, ts, name):
        with self.h5.File(self._path(scene), "r") as f:
            return str(ts) in f and name in f[str(ts)]

    def names(self, scene, ts):
        with self.h5.File(self._path(scene), "r") as f:
            return sorted(f[str(ts)].keys())

    def read(self, scene, ts, name):
        with self.h5.File(self._path(scene), "r") as f:
            return f[str(ts)][name][:]

    def write(self, scene, ts, name, data):
        with _H5_WRITE_LOCK:                                   # one writer per process (the runner's writer thread); scene
            with self.h5.File(self._path(scene), "a") as f:     # sharding keeps processes out of each other's files
                g = f.require_group(str(ts))                   # OSF/src/trainer.py:339-343
                if name in g:
                    del g[name]
                g.create_dataset(name, data=np.asarray(data))


def open_store(directory: str, prefer: Optional[str] = None) -> FrameStore:
    has_h5 = any(f.endswith(".h5") for f in os.listdir(directory)) if os.path.isdir(directory) else False
    if prefer == "npy" or not has_h5:
        return NpyStore(directory)
    return H5Store(directory)


def read_index(directory: str, name: str = "index_total.pkl") -> List[List]:
    with open(os.path.join(directory, name), "rb") as f:
        return pickle.load(f)


def write_index(directory: str, rows: Sequence[Sequence], name: str = "index_total.pkl") -> None:
    with open(os.path.join(directory, name), "wb") as f:
        pickle.dump([list(r) for r in rows], f)


def create_reading_index(directory: str, flow_inside_check: bool = False, store: Optional[FrameStore] = None) -> List[List]:
    """OSF/dataprocess/misc_data.py:32-55: scan the scenes and write `index_total.pkl` (every frame) or, with
    `flow_inside_check`, `index_flow.pkl` (frames that carry a ground-truth `flow`); rows are [scene_id, timestamp]
    with the timestamps of a scene in numeric order."""
    st = store or open_store(directory)
    rows = []
    for scene in st.scenes():
        for ts in st.timestamps(scene):
            if not flow_inside_check or st.has(scene, ts, "flow"):
                rows.append([scene, ts])
    write_index(directory, rows, "index_flow.pkl" if flow_inside_check else "index_total.pkl")
    return rows


def subset_index(full_pkl_path: str, new_folder: str, store: Optional[FrameStore] = None) -> List[List]:
    """tools/pkl_extract.py: keep the rows of an index whose scene is present in `new_folder` and write the result
    there under the same file name (how the reference cuts the 13-scene demo split out of the full index)."""
    with open(full_pkl_path, "rb") as f:
        rows = pickle.load(f)
    present = set((store or open_store(new_folder)).scenes())
    kept = [list(r) for r in rows if r[0] in present]
    write_index(new_folder, kept, os.path.basename(full_pkl_path))
    return kept


def write_synthetic_dataset(directory: str, n_scenes: int = 2, n_frames: int = 6, n_points: int = 4000,
                            seed: int = 0, eval_every: int = 2, ground_fraction: float = 0.25,
                            store: Optional[FrameStore] = None) -> FrameStore:
    """A small dataset in the reference's schema from `frames.LidarWorld` (one world per scene)."""
    from . import frames
    os.makedirs(directory, exist_ok=True)
    st = store or NpyStore(directory)
    index, eval_index = [], []
    rng = np.random.default_rng(seed)
    for s in range(n_scenes):
        scene = f"scene{seed:02d}_{s:03d}"
        world = frames.world_for_points(n_points, seed * 100 + s)
        for k in range(n_frames):
            t = 1.0 + k * world.sweep_period
            ts = str(int(1_000_000_000 + (seed * 100 + s) * 10_000_000 + k * 100_000))
            fr = world.observe(t, n_points)
            n_g = int(n_points * ground_fraction)
            if n_g:     # add labelled ground returns so the ground-removal path is exercised
                a = rng.uniform(0, 2 * np.pi, n_g); r = rng.uniform(3, 50, n_g)
                g = np.stack([r * np.cos(a), r * np.sin(a), np.full(n_g, -1.72), rng.uniform(0, 1, n_g)], 1).astype(np.float32)
                pose0, pose1 = world.ego_pose(t), world.ego_pose(t + world.sweep_period)
                gw = g[:, :3].astype(np.float64) @ pose0[:3, :3].T + pose0[:3, 3]
                g1 = (gw - pose1[:3, 3]) @ pose1[:3, :3]
                order = rng.permutation(n_points + n_g)
                cat = lambda a_, b_: np.concatenate([a_, b_])[order]
                fr = {
                    "lidar": cat(fr["lidar"], g), "ground_mask": cat(fr["ground_mask"], np.ones(n_g, bool)),
                    "pose": fr["pose"], "lidar_dt": cat(fr["lidar_dt"], (a / (2 * np.pi) * 0.1).astype(np.float32)),
                    "lidar_id": cat(fr["lidar_id"], np.zeros(n_g, np.uint8)),
                    "flow": cat(fr["flow"], (g1 - g[:, :3]).astype(np.float32)),
                    "flow_is_valid": cat(fr["flow_is_valid"], np.ones(n_g, bool)),
                    "flow_category_indices": cat(fr["flow_category_indices"], np.zeros(n_g, np.uint8)),
                    "flow_instance_id": cat(fr["flow_instance_id"], np.zeros(n_g, np.int16)),
                    "ego_motion": fr["ego_motion"],
                }
            for name, arr in fr.items():
                st.write(scene, ts, name, arr)
            index.append([scene, ts])
            if 0 < k < n_frames - 1 and k % eval_every == 0:
                eval_index.append([scene, ts])
    write_index(directory, index, "index_total.pkl")
    write_index(directory, eval_index, "index_eval.pkl")
    return st


def write_replicated_dataset(directory: str, triples: Sequence[Dict], n_scenes: int, n_frames: int, seed: int = 0,
                             eval_every: int = 10, store: Optional[FrameStore] = None) -> FrameStore:
    """A dataset of `n_scenes` x `n_frames` sweeps in the reference's schema built by cycling through the sweeps of a few
    already generated `frames.lidar_triple` results (their "frames" entry: the three full sweep dicts).  Generating a
    lidar-shaped 100 k-point sweep costs seconds of host ray casting, so throughput benchmarks of the drivers (bench.py
    `pipeline`) replicate a handful of sweeps instead; every array is still written and read per frame."""
    os.makedirs(directory, exist_ok=True)
    st = store or NpyStore(directory)
    sweeps = [fr for tr in triples for fr in tr["frames"]]
    index, eval_index = [], []
    for s in range(n_scenes):
        scene = f"scene{seed:02d}_{s:03d}"
        for k in range(n_frames):
            ts = str(int(1_000_000_000 + (seed * 100 + s) * 10_000_000 + k * 100_000))
            for name, arr in sweeps[(s * 3 + k) % len(sweeps)].items():
                st.write(scene, ts, name, arr)
            index.append([scene, ts])
            if 0 < k < n_frames - 1 and k % eval_every == 0:
                eval_index.append([scene, ts])
    write_index(directory, index, "index_total.pkl")
    write_index(directory, eval_index, "index_eval.pkl")
    return st
