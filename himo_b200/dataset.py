"""`HDF5Dataset` work-alike over a FrameStore: same constructor arguments (subset), same index and
frame-assembly rules, same keys in the returned dict as OSF/src/dataset.py:186-368.

  * `index_total.pkl` rows `[scene_id, timestamp]`; with eval=True the items come from `index_eval.pkl`
    (fallback `index_flow.pkl`)                                            (dataset.py:206-230)
  * pc1 = next row of the index; the last frame of a scene is clamped to the previous pair;
    pch1 = previous row clamped at the scene start                        (dataset.py:273-301, 321-348)
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from typing import Dict, List

import numpy as np

from .store import FrameStore, open_store, read_index

_EXTRA = ["ego_motion", "lidar_dt", "flow", "flow_is_valid", "flow_category_indices", "flow_instance_id", "dufo"]


class HDF5Dataset:
    def __init__(self, directory: str, n_frames: int = 2, eval: bool = False, vis_name="", store: FrameStore = None):
        self.directory = str(directory)
        self.store = store or open_store(self.directory)
        self.data_index = read_index(self.directory, "index_total.pkl")
        self.history_frames = n_frames - 2
        self.vis_name = vis_name if isinstance(vis_name, list) else [vis_name]
        self.eval_index = False
        if eval:
            name = "index_eval.pkl"
            if not os.path.exists(os.path.join(self.directory, name)):
                name = "index_flow.pkl"
                if not os.path.exists(os.path.join(self.directory, name)):
                    raise Exception(f"No any eval index file found! Please check {self.directory}")
            self.eval_index = True
            self.eval_data_index = read_index(self.directory, name)
        self.scene_id_bounds: Dict[str, Dict] = {}
        for idx, (scene_id, ts) in enumerate(self.data_index):
            b = self.scene_id_bounds.setdefault(scene_id, {"min_timestamp": ts, "max_timestamp": ts,
                                                           "min_index": idx, "max_index": idx})
            if ts < b["min_timestamp"]:
                b["min_timestamp"], b["min_index"] = ts, idx
            if ts > b["max_timestamp"]:
                b["max_timestamp"], b["max_index"] = ts, idx
        self._pos = {(s, str(t)): i for i, (s, t) in enumerate(self.data_index)}
        # consecutive items share clouds (frame t is pc1 of item t-1, pc0 of item t and pch1 of item t+1): a small LRU
        # over (scene, timestamp, name) turns three reads per item into one
        self._cache: "OrderedDict" = OrderedDict()
        self._cache_cap = 48
        self._cache_lock = threading.Lock()

    def _read(self, scene_id, ts, name):
        key = (scene_id, str(ts), name)
        with self._cache_lock:
            hit = self._cache.get(key)
            if hit is not None:
                self._cache.move_to_end(key)
                return hit
        arr = self.store.read(scene_id, ts, name)
        with self._cache_lock:
            self._cache[key] = arr
            while len(self._cache) > self._cache_cap:
                self._cache.popitem(last=False)
        return arr

    def __len__(self):
        return len(self.eval_data_index) if self.eval_index else len(self.data_index)

    def valid_index(self, index_):
        if self.eval_index:
            eval_i = index_
            scene_id, ts = self.eval_data_index[eval_i]
            index_ = self._pos[(scene_id, str(ts))]
            if index_ >= self.scene_id_bounds[scene_id]["max_index"]:
                _, index_ = self.valid_index(eval_i - 1)
            return True, index_
        scene_id, _ = self.data_index[index_]
        b = self.scene_id_bounds[scene_id]
        lo, hi = b["min_index"] + max(self.history_frames, 0), b["max_index"] - 1
        return False, max(lo, min(hi, index_))

    def __getitem__(self, index_) -> Dict:
        eval_flag, index_ = self.valid_index(index_)
        scene_id, ts = self.data_index[index_]
        st = self.store
        d = {"scene_id": scene_id, "timestamp": ts, "eval_flag": eval_flag}
        rd = self._read
        d["pc0"] = rd(scene_id, ts, "lidar")[:, :3]
        d["gm0"] = rd(scene_id, ts, "ground_mask")
        d["pose0"] = rd(scene_id, ts, "pose")
        nts = self.data_index[index_ + 1][1]
        d["pose1"] = rd(scene_id, nts, "pose")
        d["pc1"] = rd(scene_id, nts, "lidar")[:, :3]
        d["gm1"] = rd(scene_id, nts, "ground_mask")
        for i in range(1, self.history_frames + 1):
            fi = max(index_ - i, self.scene_id_bounds[scene_id]["min_index"])
            pts = self.data_index[fi][1]
            d[f"pch{i}"] = rd(scene_id, pts, "lidar")[:, :3]
            d[f"gmh{i}"] = rd(scene_id, pts, "ground_mask")
            d[f"poseh{i}"] = rd(scene_id, pts, "pose")
        for key in [v for v in self.vis_name if v] + _EXTRA:
            if st.has(scene_id, ts, key):
                d[key] = st.read(scene_id, ts, key)
        if self.eval_index:
            if st.has(scene_id, ts, "eval_mask"):
                d["eval_mask"] = st.read(scene_id, ts, "eval_mask")
            else:
                d["eval_mask"] = ~d["gm0"]
        return d
