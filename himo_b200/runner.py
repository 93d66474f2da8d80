"""Drivers behind the command lines kept from the reference: `save.py` (flow -> <res_name> in the frame
store), `save_zip.py` (flow -> compensation-distance zip) and `eval.py` (HiMo instance metrics).

Multi-GPU: one process per GPU -- under torchrun, or `gpus=N` on the command line, which spawns N ranks on this node
the way the reference's `launch_runner` does with mp.spawn (OSF/src/runner.py:316-343).  Frames are dealt in balanced
contiguous blocks (`shard=frame`) or whole scenes like `SceneDistributedSampler` (`shard=scene`, OSF/src/runner.py:38-88);
the only collective is the metric gather at the end (runner.py:249-256).

Per rank the frame loop is a three-stage pipeline: a reader thread prefetches frames from the store, the engine's
`infer_stream` overlaps H2D / network / D2H on three CUDA streams, and a writer thread puts results into the store --
the same `infer_stream` call bench.py's `e2e` arm times.
"""
from __future__ import annotations

import os
import queue
import sys
import threading
import time
from collections import deque
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch


def parse_overrides(argv: List[str], aliases: Optional[Dict[str, str]] = None) -> Dict[str, str]:
    """Accept both spellings the reference uses: hydra `key=value` (OSF/save.py) and fire `--key value` /
    `--key=value` (HiMo eval.py, save_zip.py)."""
    out: Dict[str, str] = {}
    aliases = aliases or {}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a.startswith("--"):
            a = a[2:]
            if "=" in a:
                k, v = a.split("=", 1)
            elif i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                k, v = a, argv[i + 1]
                i += 1
            else:
                k, v = a, "true"
        elif "=" in a:
            k, v = a.split("=", 1)
        else:
            raise SystemExit(f"cannot parse argument {a!r}")
        k = k.replace("-", "_")
        out[aliases.get(k, k)] = v
        i += 1
    return out


class _Prefetch:
    """Reader pool: `fetch(i)` for i in `indices` on `workers` threads, results handed over IN ORDER, at most `depth`
    frames ahead of the consumer (the reference gets the same overlap from DataLoader workers, OSF/src/runner.py:123-127;
    np.load / h5 reads release the GIL, so threads are enough)."""

    def __init__(self, fetch: Callable[[int], Dict], indices: Iterable[int], depth: int = 8, workers: int = 4):
        from concurrent.futures import ThreadPoolExecutor
        self.fetch, self.indices, self.depth = fetch, list(indices), max(1, depth)
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="himo-read")

    def __iter__(self) -> Iterator[Tuple[int, Dict]]:
        pending: deque = deque()
        it = iter(self.indices)
        try:
            for i in it:
                pending.append((i, self.pool.submit(self.fetch, i)))
                if len(pending) >= self.depth:
                    j, fut = pending.popleft()
                    yield j, fut.result()          # a reader exception surfaces here, in the consumer
            while pending:
                j, fut = pending.popleft()
                yield j, fut.result()
        finally:
            for _, fut in pending:
                fut.cancel()
            self.pool.shutdown(wait=False)


class _Writer:
    """Writer thread: store writes leave the frame loop (the reference writes synchronously, OSF/src/trainer.py:337-343)."""

    def __init__(self, write: Callable, depth: int = 8):
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.err: Optional[BaseException] = None
        self.write = write
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while True:
            it = self.q.get()
            if it is None:
                return
            if self.err is None:
                try:
                    self.write(*it)
                except BaseException as e:
                    self.err = e

    def put(self, *args):
        if self.err is not None:
            raise self.err
        self.q.put(args)

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err is not None:
            raise self.err


def infer_many(engine, items: Iterable[Dict]) -> Iterator[Tuple[Dict, np.ndarray]]:
    """(item, final_flow) for every item, in order, through the engine's pipelined `infer_stream` when it has one."""
    if not hasattr(engine, "infer_stream"):
        for item in items:
            yield item, engine.infer(item)
        return
    seen: deque = deque()

    def tee():
        for item in items:
            seen.append(item)
            yield item
    for final in engine.infer_stream(tee()):
        yield seen.popleft(), final


def _ensure_group(local: int, need_cuda: bool):
    """Process group for the end-of-run metric gather / barrier.  NCCL with this rank's own device when the run uses
    GPUs (the device is set BEFORE the group exists, otherwise every rank would sit on cuda:0), gloo for host-only runs."""
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    if need_cuda and torch.cuda.is_available() and local < torch.cuda.device_count():
        dev = torch.device("cuda", local)
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    return True


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_scenes(scenes: List[str], rank: int, world: int) -> List[str]:
    """SceneDistributedSampler (OSF/src/runner.py:74-80): sorted scenes, rank r takes scenes[r::world]."""
    return sorted(scenes)[rank::world]


def _build_engine(cfg: Dict[str, str], device):
    from . import weights
    model = cfg.get("model", "deflowpp")
    ckpt = cfg.get("checkpoint", "")
    if model in ("fastnsf",):
        from .engine import FastNSFEngine
        kw = {}
        for k, cast in (("itr_num", int), ("early_patience", int), ("lr", float), ("min_delta", float)):
            if k in cfg:
                kw[k] = cast(cfg[k])
        kw.setdefault("early_patience", 10)          # conf/model/fastnsf.yaml:13
        return FastNSFEngine(device=device, precision=cfg.get("precision", "fp32"), **kw), 2
    if model == "nsfp":
        from .engine import NSFPEngine
        kw = {}
        for k, cast in (("itr_num", int), ("early_patience", int), ("lr", float), ("min_delta", float)):
            if k in cfg:
                kw[k] = cast(cfg[k])
        return NSFPEngine(device=device, **kw), 2              # conf/model/nsfp.yaml: patience 30
    from .engine import SeFlowPPEngine
    if ckpt.startswith("synthetic"):
        seed = int(ckpt.split(":")[1]) if ":" in ckpt else 0
        sd = weights.synth_deflowpp_state_dict(seed)
    elif ckpt:
        sd = weights.load_deflowpp_checkpoint(ckpt)
    else:
        raise SystemExit("save.py: checkpoint=<path to seflowpp .ckpt> (or checkpoint=synthetic:<seed>) is required")
    return SeFlowPPEngine(sd, device=device, precision=cfg.get("precision", "fp32")), 3


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous frame blocks balanced to +-1 frame (2040 frames / 8 ranks = 255 each), where the
    reference can only deal whole scenes (13 scenes over 8 ranks: 2 + 2 + 2 + 2 + 2 + 1 + 1 + 1)."""
    return range(rank * n_frames // world, (rank + 1) * n_frames // world)


def run_save(cfg: Dict[str, str], engine=None, n_frames: Optional[int] = None,
             dist_env: Optional[Tuple[int, int, int]] = None) -> int:
    """`python save.py checkpoint=... dataset_path=... [res_name=...] [shard=frame|scene]` / `python save.py model=fastnsf
    dataset_path=...` (README.md:47-54; OSF/save.py:25-58, OSF/src/runner.py:316-343).  The reference shards by scene
    only to keep two processes out of one .h5 file (runner.py:74-80); the per-frame store has no such constraint, so
    frames are dealt in balanced contiguous blocks unless the store is .h5-backed or `shard=scene` is given.
    `engine` is injectable for the host tests; `dist_env` = (rank, world, local_rank) overrides the environment (bench.py
    gives every rank a private store).  Returns the number of frames this rank wrote."""
    from .dataset import HDF5Dataset
    from .store import H5Store
    rank, world, local = dist_env if dist_env is not None else _dist_env()
    data_dir = cfg.get("dataset_path") or cfg.get("data_dir")
    if not data_dir:
        raise SystemExit("dataset_path=<dir with index_total.pkl and scene files> is required")
    dev = None
    if engine is None:
        if not torch.cuda.is_available():
            raise SystemExit("himo_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", local)
        torch.cuda.set_device(dev)
        engine, n_frames = _build_engine(cfg, dev)
    ds = HDF5Dataset(data_dir, n_frames=n_frames or 2)
    res_name = cfg.get("res_name") or (cfg.get("model", "deflowpp") if "model" in cfg else "seflowpp_best")
    shard = cfg.get("shard", "scene" if isinstance(ds.store, H5Store) else "frame")
    if shard not in ("frame", "scene"):
        raise SystemExit(f"shard={shard}: frame or scene")
    if shard == "scene":
        scenes = set(shard_scenes(list(ds.scene_id_bounds.keys()), rank, world))
        mine = [i for i, (scene, _) in enumerate(ds.data_index) if scene in scenes]
    else:
        mine = shard_frames(len(ds.data_index), rank, world)
    t0, done = time.time(), 0

    def own_frames():
        for i, item in _Prefetch(ds.__getitem__, mine):
            if (item["scene_id"], item["timestamp"]) == tuple(ds.data_index[i]):
                yield item      # else: clamped duplicate of the neighbouring pair (last frame of a scene)

    writer = _Writer(ds.store.write)
    try:
        for item, final in infer_many(engine, own_frames()):
            writer.put(item["scene_id"], item["timestamp"], res_name, np.asarray(final, dtype=np.float32))
            done += 1
    finally:
        writer.close()
    if world > 1:
        import torch.distributed as dist
        own = _ensure_group(local, dev is not None)
        dist.barrier()
        if own:
            dist.destroy_process_group()
    if rank == 0:
        print(f"[save] wrote '{res_name}' for {done} frames on rank 0 of {world} ({shard} shards) in {time.time() - t0:.1f}s -> {data_dir}")
    return done


def run_save_zip(cfg: Dict[str, str]) -> str:
    """`python save_zip.py --data_dir ... --res_name ...` (save_zip.py:102-125)."""
    from . import himo
    from .dataset import HDF5Dataset
    data_dir = cfg["data_dir"]
    res_name = cfg.get("res_name", "seflowpp_best")
    out_dir = os.path.join(data_dir, "results")
    os.makedirs(out_dir, exist_ok=True)
    ds = HDF5Dataset(data_dir, vis_name=res_name, eval=True)
    for i in range(len(ds)):
        data = ds[i]
        comp = himo.comp_dis_from_total_flow(data, res_name)
        himo.write_output_file(comp, (data["scene_id"], str(data["timestamp"])), out_dir)
    path = himo.zip_res(out_dir, output_file=os.path.join(out_dir, f"{res_name}-submit.zip"))
    print(f"Zipped results into {path}")
    return path


def run_eval(cfg: Dict[str, str]):
    """`python eval.py --data_dir ... --res_name ... | --comp_dis_zip ...` (eval.py:270-312)."""
    from . import himo
    from .dataset import HDF5Dataset
    data_dir = cfg["data_dir"]
    res_name = cfg.get("res_name", "")
    zip_path = cfg.get("comp_dis_zip", "")
    data_name, flag = himo.check_valid(data_dir, res_name, zip_path)
    rank, world, local = _dist_env()
    # the per-instance Chamfer of every frame runs as one batched device launch when a GPU is there (himo_segmented_nn);
    # `metrics_device=host` forces the reference's scipy path
    mdev = cfg.get("metrics_device", f"cuda:{local}" if torch.cuda.is_available() and local < torch.cuda.device_count() else "host")
    metrics = himo.InstanceMetrics(data_name, device=None if mdev == "host" else mdev)
    ds = HDF5Dataset(data_dir, vis_name=res_name if flag == 2 else "", eval=True)
    for i in range(rank, len(ds), world):
        data = ds[i]
        pc0 = data["pc0"]
        pf = himo.pose_flow_np(pc0, data["pose0"], data["pose1"])
        gt_flow = data["flow"] - pf
        m = himo.eval_masks(data, data_name)
        dt0 = max(data["lidar_dt"]) - data["lidar_dt"]
        if flag == 2:
            est = np.zeros_like(pf) if res_name == "raw" else (data[res_name] - pf)
            metrics.step_eval(pc0[m], gt_flow[m], dt0[m], data["flow_category_indices"][m],
                              data["flow_instance_id"][m], est_flow=est[m])
        else:
            comp = himo.read_output_zip(zip_path, (data["scene_id"], str(data["timestamp"])))
            metrics.step_eval(pc0[m], gt_flow[m], dt0[m], data["flow_category_indices"][m],
                              data["flow_instance_id"][m], est_dis=comp[m])
    if world > 1:
        import torch.distributed as dist
        _ensure_group(local, True)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(metrics, gathered, dst=0)
        if rank == 0:
            for other in gathered[1:]:
                metrics.merge(other)
        dist.barrier()
    if rank == 0:
        return metrics.print(res_name=res_name or "zip", file_name=cfg.get("out_json", f"res-{data_name}.json"))
    return None


class _StoredFlow:
    def __init__(self, name: str):
        self.name = name

    def infer(self, item):
        if self.name not in item:
            raise SystemExit(f"frame {item['scene_id']}/{item['timestamp']} holds no '{self.name}'")
        return item[self.name]


def run_validate(cfg: Dict[str, str], engine=None, n_frames: Optional[int] = None):
    """OSF `python eval.py checkpoint=... dataset_path=... [data_mode=val]` (OSF/eval.py:27-72): run the model over the
    eval index and score it with the AV2 / OpenSceneFlow metrics (ModelWrapper.eval_only_step_, OSF/src/trainer.py:
    251-278; on_validation_epoch_end :226-249).  Frames of the eval index are dealt round-robin over the ranks; the
    un-normalised `OfficialMetrics` are gathered on rank 0 and merged.  `engine` is injectable for the host tests."""
    from . import av2_metrics as M, himo
    from .dataset import HDF5Dataset
    rank, world, local = _dist_env()
    data_dir = cfg.get("dataset_path") or cfg.get("data_dir")
    if not data_dir:
        raise SystemExit("dataset_path=<dir with index_eval.pkl and scene files> is required")
    stored = cfg.get("res_name", "") if cfg.get("model") == "stored" else ""
    if cfg.get("model") == "stored":
        # score flows a previous save.py run left in the store (no model, no GPU): `model=stored res_name=<name>`
        if not stored:
            raise SystemExit("model=stored needs res_name=<dataset name written by save.py>")
        engine, n_frames = _StoredFlow(stored), 2
    elif engine is None:
        if not torch.cuda.is_available():
            raise SystemExit("himo_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", local)
        torch.cuda.set_device(dev)
        engine, n_frames = _build_engine(cfg, dev)
    ds = HDF5Dataset(data_dir, n_frames=n_frames or 2, eval=True, vis_name=stored)
    truthy = lambda v: str(v).lower() in ("1", "true", "yes")
    data_mode = cfg.get("data_mode", "val")
    if data_mode not in ("val", "test"):
        raise SystemExit(f"data_mode={data_mode}: val or test")
    version = int(cfg.get("leaderboard_version", 1))
    if version not in (1, 2):
        raise ValueError(f"Leaderboard version {version} is not valid. Please set it to 1 or 2.")     # OSF/eval.py:29-30
    save_res = data_mode == "test" or truthy(cfg.get("save_res", "false"))         # trainer.py:281
    res_dir = cfg.get("save_res_path") or os.path.join(os.path.dirname(os.path.abspath(data_dir)), "results",
                                                       cfg.get("output", f"{cfg.get('model', 'deflowpp')}-{data_mode}-v{version}"))
    metrics = M.OfficialMetrics()
    t0, done = time.time(), 0
    frames_it = (item for _, item in _Prefetch(ds.__getitem__, range(rank, len(ds), world)))
    for item, final in infer_many(engine, frames_it):
        final = np.asarray(final, np.float32)
        pc0 = np.asarray(item["pc0"], np.float32)[:, :3]
        pose_flow = himo.pose_flow_np(pc0, item["pose0"], item["pose1"]).astype(np.float32)
        m = np.asarray(item["eval_mask"], bool).squeeze()
        if data_mode == "val":                  # only val carries ground truth (trainer.py:269)
            gt, valid, cats = item["flow"], item["flow_is_valid"], item["flow_category_indices"]
            metrics.step(M.evaluate_leaderboard(final[m], pose_flow[m], pc0[m], gt[m], valid[m], cats[m]),
                         M.evaluate_leaderboard_v2(final[m], pose_flow[m], pc0[m], gt[m], valid[m], cats[m]),
                         M.evaluate_ssf(final, pose_flow, pc0, gt, valid, cats))
        if save_res:
            from . import av2_submit
            flow_out, is_dyn = av2_submit.leaderboard_arrays(final, pose_flow, m, version)
            av2_submit.write_output_file(flow_out, is_dyn, (item["scene_id"], item["timestamp"]), res_dir, version)
        done += 1
    if world > 1:
        import torch.distributed as dist
        _ensure_group(local, True)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(metrics, gathered, dst=0)
        if rank == 0:
            for other in gathered[1:]:
                metrics.merge(other)
        dist.barrier()
    if rank != 0:
        return None
    print(f"[eval] {done} frames on rank 0 of {world} in {time.time() - t0:.1f}s")
    if data_mode == "test":                     # trainer.py:213-224: zip for the online leaderboard, no metrics
        from . import av2_submit
        path = av2_submit.zip_res(res_dir, output_file=res_dir.rstrip("/") + ".zip", leaderboard_version=version,
                                  is_supervised=truthy(cfg.get("supervised_flag", "true")))
        print(f"Test results saved in: {res_dir}, zipped into {path}")
        return path
    metrics.normalize()
    metrics.print(ssf_metrics=cfg.get("ssf_metrics", "false").lower() in ("1", "true", "yes"))
    out = cfg.get("out_json")
    if out:
        import json
        clean = lambda v: None if isinstance(v, float) and np.isnan(v) else v
        json.dump({"epe_3way": {k: clean(float(v)) for k, v in metrics.epe_3way.items()},
                   "bucketed": {k: {t: clean(float(v[t])) for t in ("Static", "Dynamic")} for k, v in metrics.bucketed.items()}},
                  open(out, "w"), indent=1)
    return metrics


def _spawn_rank(rank: int, world: int, port: int, fn_name: str, cfg: Dict[str, str]):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    globals()[fn_name](cfg)


def _launch(fn_name: str, cfg: Dict[str, str]):
    """`gpus=N` (N > 1) outside torchrun: spawn N ranks on this node, one per GPU (launch_runner, OSF/src/runner.py:
    316-343: mp.spawn over torch.cuda.device_count()).  Under torchrun, or with gpus=1 / absent, run in this process."""
    gpus = cfg.pop("gpus", None)
    n = 1
    if gpus is not None:
        n = torch.cuda.device_count() if str(gpus) in ("all", "-1") else int(gpus)
        if n < 1:
            raise SystemExit(f"gpus={gpus}: need a positive GPU count")
    if n > 1 and "WORLD_SIZE" not in os.environ:
        if torch.cuda.is_available() and n > torch.cuda.device_count():
            raise SystemExit(f"gpus={n} but this node has {torch.cuda.device_count()} CUDA device(s)")
        import socket
        import torch.multiprocessing as mp
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        mp.spawn(_spawn_rank, args=(n, port, fn_name, cfg), nprocs=n, join=True)
        return None
    return globals()[fn_name](cfg)


def main_save(argv=None):
    return _launch("run_save", parse_overrides(sys.argv[1:] if argv is None else argv))


def main_save_zip(argv=None):
    run_save_zip(parse_overrides(sys.argv[1:] if argv is None else argv))


def main_eval(argv=None):
    """HiMo's eval.py (`--data_dir/--res_name/--comp_dis_zip`: compensation-distance metrics) and, when a model is named
    the hydra way (`checkpoint=` / `model=`), OpenSceneFlow's eval.py (scene-flow metrics of a fresh model run)."""
    cfg = parse_overrides(sys.argv[1:] if argv is None else argv, aliases={"flow_mode": "res_name"})
    if "checkpoint" in cfg or "model" in cfg:
        return _launch("run_validate", cfg)
    return _launch("run_eval", cfg)
