"""CPU: host logic around the path -- frame store, HDF5Dataset work-alike, HiMo compensation / metrics
against the reference's own Python (utils/__init__.py, tools/test/score.py) when /root/reference is
present, CLI argument parsing, scene sharding, and the world_size-2 gloo metric gather."""
import os
import sys

import numpy as np
import pytest

from himo_b200 import himo, runner, store
from himo_b200.dataset import HDF5Dataset

REF = "/root/reference"


@pytest.fixture(scope="module")
def dataset_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("himo_av2_synth")
    store.write_synthetic_dataset(str(d), n_scenes=2, n_frames=5, n_points=2500, seed=3)
    return str(d)


def test_store_roundtrip_and_replace(tmp_path):
    st = store.NpyStore(str(tmp_path))
    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    st.write("s", "100", "flow", a)
    st.write("s", "100", "flow", a + 1)             # del + create semantics: re-runs replace
    assert (st.read("s", "100", "flow") == a + 1).all() and st.has("s", "100", "flow")
    assert st.scenes() == ["s"] and st.names("s", "100") == ["flow"]


def test_dataset_pair_assembly(dataset_dir):
    ds = HDF5Dataset(dataset_dir, n_frames=3)
    assert len(ds) == 10
    first, last = ds[0], ds[4]
    # history clamps at the scene start; the last frame of a scene is served by the previous pair
    assert first["timestamp"] == ds.data_index[1][1] or first["timestamp"] == ds.data_index[0][1]
    assert last["timestamp"] == ds.data_index[3][1]
    it = ds[2]
    assert it["pc0"].shape[1] == 3 and it["gm0"].dtype == bool and it["pose0"].shape == (4, 4)
    assert (it["pose1"] == ds.store.read(it["scene_id"], ds.data_index[3][1], "pose")).all()
    assert (it["poseh1"] == ds.store.read(it["scene_id"], ds.data_index[1][1], "pose")).all()
    ev = HDF5Dataset(dataset_dir, eval=True)
    assert len(ev) == 2 and ev[0]["eval_flag"] and "eval_mask" in ev[0] and "lidar_dt" in ev[0]


def test_parse_overrides_and_sharding():
    cfg = runner.parse_overrides(["checkpoint=a.ckpt", "dataset_path=/d", "--res_name", "x", "--flow-mode=y"],
                                 aliases={"flow_mode": "res_name2"})
    assert cfg == {"checkpoint": "a.ckpt", "dataset_path": "/d", "res_name": "x", "res_name2": "y"}
    scenes = [f"s{i}" for i in range(13)]
    parts = [runner.shard_scenes(scenes, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == sorted(scenes) and [len(p) for p in parts] == [2, 2, 2, 2, 2, 1, 1, 1]


def test_compdis_and_zip_roundtrip(dataset_dir, tmp_path):
    ds = HDF5Dataset(dataset_dir, eval=True)
    data = ds[0]
    data["gt"] = data["flow"]
    comp = himo.comp_dis_from_total_flow(data, "gt")
    dt0 = data["lidar_dt"].max() - data["lidar_dt"]
    pf = himo.pose_flow_np(data["pc0"], data["pose0"], data["pose1"])
    np.testing.assert_allclose(comp, (data["flow"] - pf) / 0.1 * dt0[:, None], rtol=1e-6, atol=1e-7)
    himo.write_output_file(comp, (data["scene_id"], str(data["timestamp"])), tmp_path / "results")
    z = himo.zip_res(tmp_path / "results", str(tmp_path / "x-submit.zip"))
    back = himo.read_output_zip(z, (data["scene_id"], str(data["timestamp"])))
    assert back.dtype == np.float32 and (back == comp.astype(np.float32)).all()


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference")
def test_against_reference_utils_and_scorer(dataset_dir):
    """flow2compDis / ego_pts_mask against HiMo's own utils, and InstanceMetrics against the reference's
    stand-alone scorer (tools/test/score.py, 'matching eval.py exactly')."""
    sys.path.insert(0, REF)
    import importlib
    ref_utils = importlib.import_module("utils")
    spec = importlib.util.spec_from_file_location("himo_ref_score", os.path.join(REF, "tools", "test", "score.py"))
    score = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(score)
    sys.path.remove(REF)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-12, 12, (5000, 3)).astype(np.float32)
    assert (himo.ego_pts_mask(pts) == ref_utils.ego_pts_mask(pts)).all()
    fl = rng.normal(size=(5000, 3)).astype(np.float32)
    dt = rng.uniform(0, 0.1, 5000).astype(np.float32)
    assert (himo.flow2compDis(fl, dt, 0.1) == ref_utils.flow2compDis(fl, dt, 0.1)).all()
    assert himo.CATEGORY_TO_INDEX == score.CATEGORY_TO_INDEX
    ours, ref = himo.InstanceMetrics("av2"), score.ScoreMetrics()
    ds = HDF5Dataset(dataset_dir, eval=True)
    for i in range(len(ds)):
        d = ds[i]
        pf = himo.pose_flow_np(d["pc0"], d["pose0"], d["pose1"])
        gt = d["flow"] - pf
        est = gt + rng.normal(0, 0.05, gt.shape).astype(np.float32)
        m = himo.eval_masks(d, "av2")
        dt0 = d["lidar_dt"].max() - d["lidar_dt"]
        ours.step_eval(d["pc0"][m], gt[m], dt0[m], d["flow_category_indices"][m], d["flow_instance_id"][m], est_flow=est[m])
        ref.step(himo.flow2compDis(gt, dt0, 0.1), himo.flow2compDis(est, dt0, 0.1), m, d["flow_category_indices"],
                 d["flow_instance_id"].astype(np.uint32), np.linalg.norm(gt, axis=1), d["pc0"])
    s, r = ours.summary(), ref.compute_scores()
    assert "CAR" in s or "OTHER_VEHICLES" in s, "synthetic world produced no moving instances"
    for c, key in (("CAR", "car"), ("OTHER_VEHICLES", "others")):
        if c in s:
            assert abs(s[c]["overall"]["mpe"] - r[f"{key}_mpe"]) < 1e-6
            assert abs(s[c]["overall"]["cd"] - r[f"{key}_cde"]) < 1e-6
            assert s[c]["overall"]["num_pts"] == r[f"{key}_num_pts"]
    assert abs(s["Total"]["mpe"] - r["mpe"]) < 1e-6 and abs(s["Total"]["cd"] - r["chamfer"]) < 1e-6


def _gather_worker(rank, world, port, data_dir, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = runner.run_eval({"data_dir": data_dir, "res_name": "flow", "out_json": out})
    dist.destroy_process_group()
    if rank == 0:
        assert res is not None


def test_eval_two_ranks_matches_one(dataset_dir, tmp_path):
    """world_size-2 gloo run of the eval driver: the merged metrics equal the single-process ones."""
    import json
    import torch.multiprocessing as mp
    d = dataset_dir  # name contains 'av2'
    one = runner.run_eval({"data_dir": d, "res_name": "flow", "out_json": str(tmp_path / "one.json")})
    mp.spawn(_gather_worker, args=(2, 29613, d, str(tmp_path / "two.json")), nprocs=2, join=True)
    a = json.load(open(tmp_path / "one.json"))["av2"]["flow"]
    b = json.load(open(tmp_path / "two.json"))["av2"]["flow"]
    assert a.keys() == b.keys() and len(a) > 0
    for c in a:
        assert a[c]["overall"]["num_pts"] == b[c]["overall"]["num_pts"]
        assert abs(a[c]["overall"]["mpe"] - b[c]["overall"]["mpe"]) < 1e-9


class _NoisyGtEngine:
    """Stand-in model for the host tests of the validation driver: ground truth plus seeded per-frame noise."""

    def infer(self, item):
        rng = np.random.default_rng(int(item["timestamp"]) % (2 ** 31))
        return (item["flow"] + rng.normal(0, 0.03, item["flow"].shape)).astype(np.float32)


def _validate_worker(rank, world, port, data_dir, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = runner.run_validate({"dataset_path": data_dir, "out_json": out}, engine=_NoisyGtEngine())
    dist.destroy_process_group()
    assert (res is not None) == (rank == 0)


def test_validate_driver_and_two_rank_merge(dataset_dir, tmp_path):
    """OSF eval.py work-alike (runner.run_validate): per-frame AV2 metrics of a model run over the eval index; the
    world_size-2 gloo run merges to the single-process result; a perfect model scores zero."""
    import json
    import torch.multiprocessing as mp
    from himo_b200 import av2_metrics as M
    one = runner.run_validate({"dataset_path": dataset_dir, "out_json": str(tmp_path / "v1.json")}, engine=_NoisyGtEngine())
    assert isinstance(one, M.OfficialMetrics) and one.norm_flag
    assert 0 < one.epe_3way["Three-way"] < 0.1
    mp.spawn(_validate_worker, args=(2, 29617, dataset_dir, str(tmp_path / "v2.json")), nprocs=2, join=True)
    a, b = json.load(open(tmp_path / "v1.json")), json.load(open(tmp_path / "v2.json"))
    for k in ("EPE_FD", "EPE_BS", "EPE_FS", "IoU", "Three-way"):       # per-frame means: order-independent up to rounding
        assert a["epe_3way"][k] == pytest.approx(b["epe_3way"][k], rel=1e-9, abs=1e-12)
    for k, v in a["bucketed"].items():
        for t in ("Static", "Dynamic"):
            assert (v[t] is None and b["bucketed"][k][t] is None) or v[t] == pytest.approx(b["bucketed"][k][t], rel=1e-9, abs=1e-12)

    # `model=stored res_name=flow`: score what is already in the store (here the ground truth itself) without a model
    p = runner.run_validate({"dataset_path": dataset_dir, "model": "stored", "res_name": "flow"})
    assert p.epe_3way["Three-way"] == pytest.approx(0.0, abs=1e-9) and p.epe_3way["IoU"] == pytest.approx(1.0)
    with pytest.raises(SystemExit):
        runner.run_validate({"dataset_path": dataset_dir, "model": "stored", "res_name": "not_there"})


class _ScaledGtEngine:
    def infer(self, item):
        return item["flow"] * np.float32(0.5)


def _save_worker(rank, world, port, data_dir, shard, counts):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = runner.run_save({"dataset_path": data_dir, "res_name": f"half_{shard}", "shard": shard}, engine=_ScaledGtEngine())
    out = [None] * world
    dist.all_gather_object(out, n)
    if rank == 0:
        np.save(counts, np.array(out))
    dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["frame", "scene"])
def test_save_driver_shards_frames_over_two_ranks(tmp_path, shard):
    """save.py driver with an injected engine: every writable frame is written exactly once across two ranks; frame
    shards are balanced to one frame where scene shards are as uneven as the scenes (3 scenes over 2 ranks)."""
    import torch.multiprocessing as mp
    d = str(tmp_path / "av2_synth")
    st = store.write_synthetic_dataset(d, n_scenes=3, n_frames=4, n_points=600, seed=13)
    counts = str(tmp_path / "counts.npy")
    mp.spawn(_save_worker, args=(2, 29621 + (shard == "scene"), d, shard, counts), nprocs=2, join=True)
    per_rank = np.load(counts)
    ds = HDF5Dataset(d, store=st)
    written = [(s, t) for s, t in ds.data_index if st.has(s, t, f"half_{shard}")]
    assert per_rank.sum() == len(written) == 9              # the last frame of each scene has no successor
    for s, t in written:
        np.testing.assert_array_equal(st.read(s, t, f"half_{shard}"), st.read(s, t, "flow") * np.float32(0.5))
    if shard == "frame":
        assert abs(int(per_rank[0]) - int(per_rank[1])) <= 2 and list(runner.shard_frames(2040, 3, 8)) == list(range(765, 1020))
    else:
        assert sorted(per_rank.tolist()) == [3, 6]


def test_create_reading_index_rebuilds_the_indices(tmp_path):
    """store.create_reading_index (OSF/dataprocess/misc_data.py:32-55): the rebuilt index_total equals the one the
    writer produced; index_flow lists only frames that carry ground-truth flow; timestamps sort numerically."""
    d = str(tmp_path / "av2_synth")
    st = store.write_synthetic_dataset(d, n_scenes=2, n_frames=4, n_points=300, seed=2)
    want = store.read_index(d, "index_total.pkl")
    os.remove(os.path.join(d, "index_total.pkl"))
    assert store.create_reading_index(d, store=st) == sorted(want) == store.read_index(d, "index_total.pkl")
    scene = want[0][0]
    st.write(scene, "999", "lidar", np.zeros((1, 4), np.float32))        # a frame without flow, short timestamp
    total = store.create_reading_index(d, store=st)
    assert total[0] == [scene, "999"] and len(total) == len(want) + 1     # 999 < 10-digit stamps numerically
    flow = store.create_reading_index(d, flow_inside_check=True, store=st)
    assert [scene, "999"] not in flow and len(flow) == len(want) and store.read_index(d, "index_flow.pkl") == flow
    # tools/pkl_extract.py: cut an index down to the scenes present in another folder
    import shutil
    demo = str(tmp_path / "demo")
    os.makedirs(demo)
    shutil.copytree(os.path.join(d, f"{scene}.frames"), os.path.join(demo, f"{scene}.frames"))
    kept = store.subset_index(os.path.join(d, "index_total.pkl"), demo)
    assert kept == [r for r in total if r[0] == scene] == store.read_index(demo, "index_total.pkl") and 0 < len(kept) < len(total)


def test_config4_index_shapes_and_shard_balance():
    """BASELINE config 4 (the reference's AV2 demo split): 2040 frames in 13 scenes, 70 eval frames (SURVEY.md 8(c)
    item 3; tests/golden/av2_demo_index_shape.json from /root/reference/assets/docs/av2/index_*.pkl).  Frame shards put
    255 frames on each of 8 ranks; the reference's scene shards put between 156 and 314."""
    import json
    from conftest import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "av2_demo_index_shape.json")))
    scenes = sorted(g["frames_per_scene"])
    assert g["total_rows"] == sum(g["frames_per_scene"].values()) == 2040 and len(scenes) == 13
    assert g["eval_rows"] == sum(g["eval_frames_per_scene"].values()) == 70
    if os.path.isdir(os.path.join(REF, "assets", "docs", "av2")):
        rows = store.read_index(os.path.join(REF, "assets", "docs", "av2"), "index_total.pkl")
        assert len(rows) == 2040 and {s for s, _ in rows} == set(scenes)
        assert all(isinstance(t, str) and int(t) > 0 for _, t in rows)
    per_rank_frame = [len(runner.shard_frames(2040, r, 8)) for r in range(8)]
    per_rank_scene = [sum(g["frames_per_scene"][s] for s in runner.shard_scenes(scenes, r, 8)) for r in range(8)]
    assert per_rank_frame == [255] * 8
    assert sum(per_rank_scene) == 2040 and max(per_rank_scene) >= 313 and min(per_rank_scene) <= 159
    covered = sorted(i for r in range(8) for i in runner.shard_frames(2040, r, 8))
    assert covered == list(range(2040))
