"""GPU: the kept command lines end to end on a synthetic AV2-shaped dataset:
save.py (SeFlow++ and FastNSF) -> <res_name> in the frame store -> save_zip.py -> eval.py (flow and zip)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from himo_b200 import himo, store, weights
from himo_b200.dataset import HDF5Dataset
from oracle import deflowpp_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli_av2_synth")
    store.write_synthetic_dataset(str(d), n_scenes=2, n_frames=5, n_points=3000, seed=5)
    return str(d)


def _run(args):
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_save_seflowpp_then_zip_then_eval(data_dir):
    _run(["save.py", "checkpoint=synthetic:4", f"dataset_path={data_dir}", "res_name=seflowpp_synth"])
    ds = HDF5Dataset(data_dir, n_frames=3, vis_name="seflowpp_synth")
    item = ds[2]
    got = item["seflowpp_synth"]
    assert got.shape == item["pc0"].shape and got.dtype == np.float32
    # the stored array is what ModelWrapper.test_step writes: check it against the CPU oracle
    sd = weights.synth_deflowpp_state_dict(4)
    ng = lambda pc, gm: pc[~gm]
    res = deflowpp_ref.deflowpp_forward(sd, ng(item["pch1"], item["gmh1"]), ng(item["pc0"], item["gm0"]),
                                        ng(item["pc1"], item["gm1"]), item["poseh1"], item["pose0"], item["pose1"])
    ref = deflowpp_ref.final_flow(torch.from_numpy(item["pc0"]), torch.from_numpy(item["gm0"]), item["pose0"],
                                  item["pose1"], res).numpy()
    assert np.abs(got - ref).max() <= 1e-4
    out = _run(["save_zip.py", "--data_dir", data_dir, "--res_name", "seflowpp_synth"])
    z = os.path.join(data_dir, "results", "seflowpp_synth-submit.zip")
    assert os.path.exists(z), out
    j1 = os.path.join(data_dir, "flow.json")
    j2 = os.path.join(data_dir, "zip.json")
    _run(["eval.py", "--data_dir", data_dir, "--res_name", "seflowpp_synth", "--out_json", j1])
    _run(["eval.py", "--data_dir", data_dir, "--flow_mode", "seflowpp_synth", "--comp_dis_zip", z, "--out_json", j2])
    a = json.load(open(j1))["av2"]["seflowpp_synth"]
    b = json.load(open(j2))["av2"]["seflowpp_synth"]
    assert a.keys() == b.keys() and len(a) > 0
    for c in a:   # the zip carries float32 compensation distances: same metrics up to that rounding
        assert abs(a[c]["overall"]["mpe"] - b[c]["overall"]["mpe"]) < 1e-5


def test_save_fastnsf_plumbing(data_dir):
    """BASELINE configs[0]: FastNSF through the save.py command line (few iterations: plumbing)."""
    _run(["save.py", "model=fastnsf", f"dataset_path={data_dir}", "itr_num=6", "res_name=fastnsf"])
    ds = HDF5Dataset(data_dir, vis_name="fastnsf")
    item = ds[1]
    f = item["fastnsf"]
    assert f.shape == item["pc0"].shape and np.isfinite(f).all()
    pf = himo.pose_flow_np(item["pc0"], item["pose0"], item["pose1"])
    # ground points carry the ego-motion flow only (OSF/src/runner.py:149-155)
    np.testing.assert_allclose(f[item["gm0"]], pf[item["gm0"]], rtol=0, atol=1e-4)


def test_save_nsfp_plumbing(data_dir):
    """NSFP through the save.py command line (few iterations: plumbing)."""
    _run(["save.py", "model=nsfp", f"dataset_path={data_dir}", "itr_num=4", "res_name=nsfp"])
    ds = HDF5Dataset(data_dir, vis_name="nsfp")
    item = ds[1]
    f = item["nsfp"]
    assert f.shape == item["pc0"].shape and np.isfinite(f).all()
    pf = himo.pose_flow_np(item["pc0"], item["pose0"], item["pose1"])
    np.testing.assert_allclose(f[item["gm0"]], pf[item["gm0"]], rtol=0, atol=1e-4)
    assert np.abs(f[~item["gm0"]] - pf[~item["gm0"]]).max() > 1e-4       # the optimised flow was added


def test_validate_cli_matches_stored_flows(data_dir, tmp_path):
    """OSF eval.py work-alike: `eval.py checkpoint=... dataset_path=...` runs the model over the eval index and prints
    the AV2 metrics; the same metrics come out of scoring the flows save.py stored for the same checkpoint."""
    _run(["save.py", "checkpoint=synthetic:4", f"dataset_path={data_dir}", "res_name=seflowpp_val"])
    j1, j2 = str(tmp_path / "val.json"), str(tmp_path / "stored.json")
    out = _run(["eval.py", "checkpoint=synthetic:4", f"dataset_path={data_dir}", f"out_json={j1}"])
    assert "Three-way" in out
    _run(["eval.py", "model=stored", "res_name=seflowpp_val", f"dataset_path={data_dir}", f"out_json={j2}"])
    got, ref = json.load(open(j1)), json.load(open(j2))
    for k in ("EPE_FD", "EPE_BS", "EPE_FS", "IoU", "Three-way"):
        assert np.isfinite(got["epe_3way"][k])
        assert got["epe_3way"][k] == pytest.approx(ref["epe_3way"][k], abs=2e-5)


def test_nnd_autolabel_matches_bruteforce(tmp_path):
    """himo_b200.autolabel.run_nnd (OSF/process.py:106-172) on a synthetic two-scene store: the labels written under
    `nnd` equal the reference rule applied to brute-force nearest-neighbour distances (oracle/leaf_ops.c)."""
    import numpy as np
    from himo_b200 import autolabel, store
    from himo_b200.dataset import HDF5Dataset
    from oracle import leaf
    d = str(tmp_path / "ds")
    st = store.write_synthetic_dataset(d, n_scenes=2, n_frames=4, n_points=3000, seed=5)
    n = autolabel.run_nnd(d, min_nnd=0.14, store=st)
    ds = HDF5Dataset(d, store=st)
    assert n == len(ds.data_index) == 8
    moving = 0
    for scene_id, b in ds.scene_id_bounds.items():
        norm = st.read(scene_id, ds.data_index[b["min_index"]][1], "pose")
        for i in range(b["min_index"], b["max_index"] + 1):
            ts = ds.data_index[i][1]
            j = i - 1 if i == b["max_index"] else i + 1
            ts1 = ds.data_index[j][1]
            pc0 = st.read(scene_id, ts, "lidar")[:, :3]
            pose0 = np.linalg.inv(norm) @ st.read(scene_id, ts, "pose")
            pose1 = np.linalg.inv(norm) @ st.read(scene_id, ts1, "pose")
            ego = np.linalg.inv(pose1) @ pose0
            tr0 = (pc0 @ ego[:3, :3].T + ego[:3, 3]).astype(np.float32)
            pc1 = np.ascontiguousarray(st.read(scene_id, ts1, "lidar")[:, :3]).astype(np.float32)
            d0 = leaf.nn_bruteforce(np.ascontiguousarray(tr0), pc1)[0]
            ref = ((d0 >= pow(0.14, 2)) & (d0 < pow(4.4, 2))).astype(np.uint8)      # process.py:124
            got = st.read(scene_id, ts, "nnd")
            assert got.dtype == np.uint8 and (got == ref).all()
            moving += int(ref.sum())
    assert moving > 0


def test_instance_chamfer_on_device_equals_host_path(data_dir):
    """(f)3: InstanceMetrics with the batched device Chamfer (himo_segmented_nn) against the reference's scipy path
    (eval.py:50-62) -- per-instance values to 1e-9, the whole eval.py result table equal."""
    rng = np.random.default_rng(3)
    pairs = [(rng.normal(size=(n, 3)) * 2 + k, rng.normal(size=(m, 3)) * 2 + k)
             for k, (n, m) in enumerate([(10, 10), (1, 7), (300, 257), (4000, 3500), (513, 12)])]
    got = himo.chamfer_mean_nn_batched(pairs, "cuda")
    ref = [himo.chamfer_mean_nn(a, b) for a, b in pairs]
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)
    from himo_b200 import runner
    j_host, j_dev = os.path.join(data_dir, "m_host.json"), os.path.join(data_dir, "m_dev.json")
    store.write_synthetic_dataset  # (the module fixture already holds a scored dataset with a stored flow)
    _run(["save.py", "checkpoint=synthetic:4", f"dataset_path={data_dir}", "res_name=seflowpp_synth"])
    runner.run_eval({"data_dir": data_dir, "res_name": "seflowpp_synth", "out_json": j_host, "metrics_device": "host"})
    runner.run_eval({"data_dir": data_dir, "res_name": "seflowpp_synth", "out_json": j_dev, "metrics_device": "cuda:0"})
    a, b = json.load(open(j_host)), json.load(open(j_dev))

    def walk(x, y):
        assert type(x) is type(y)
        if isinstance(x, dict):
            assert x.keys() == y.keys()
            for k in x:
                walk(x[k], y[k])
        elif isinstance(x, float):
            assert abs(x - y) <= 1e-9 * max(1.0, abs(x)), (x, y)
        else:
            assert x == y
    walk(a, b)
    assert len(a["av2"]["seflowpp_synth"]) > 0


def test_save_into_h5_scene_files_equals_npy_store(tmp_path):
    """(f)2: save.py on `.h5` scene files (the reference's on-disk contract, served by himo_b200.h5lite here) writes the same
    flow as on the per-array store -- the result goes INTO the scene file next to `lidar` / `pose`, like trainer.py:337-343."""
    from himo_b200 import h5lite
    d_npy, d_h5 = str(tmp_path / "av2_npy"), str(tmp_path / "av2_h5")
    os.makedirs(d_h5)
    store.write_synthetic_dataset(d_npy, n_scenes=1, n_frames=4, n_points=3000, seed=9)
    store.write_synthetic_dataset(d_h5, n_scenes=1, n_frames=4, n_points=3000, seed=9, store=store.H5Store(d_h5, backend=h5lite))
    for d in (d_npy, d_h5):
        _run(["save.py", "checkpoint=synthetic:4", f"dataset_path={d}", "res_name=seflowpp_synth"])
    a, b = HDF5Dataset(d_npy, n_frames=3, vis_name="seflowpp_synth"), HDF5Dataset(d_h5, n_frames=3, vis_name="seflowpp_synth")
    assert isinstance(b.store, store.H5Store)
    for i in (1, 2):
        assert np.array_equal(a[i]["seflowpp_synth"], b[i]["seflowpp_synth"])
    (scene,) = b.store.scenes()
    with h5lite.File(os.path.join(d_h5, scene + ".h5")) as f:
        ts = f.keys()[1]
        assert {"lidar", "pose", "ground_mask", "seflowpp_synth"} <= set(f[ts].keys())
        assert f[ts]["seflowpp_synth"].dtype == np.float32 and f[ts]["seflowpp_synth"].shape[1] == 3
