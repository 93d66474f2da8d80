"""CPU: frame / pose bookkeeping of the `nnd` auto-label pass (himo_b200/autolabel.py, OSF/process.py:106-172) with the
device search swapped for the brute-force oracle; the CUDA path itself is covered by tests/test_gpu_cli.py."""
import numpy as np
import pytest
import torch

from himo_b200 import autolabel, store
from himo_b200.dataset import HDF5Dataset
from oracle import leaf


def _oracle_nnd(pc0, pc1, moving_threshold=0.14, truncated=4.4):
    d0 = leaf.nn_bruteforce(pc0.cpu().numpy(), pc1.cpu().numpy())[0]
    return ((d0 >= pow(moving_threshold, 2)) & (d0 < pow(truncated, 2))).astype(np.uint8)


def test_run_nnd_refuses_without_cuda(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    with pytest.raises(EnvironmentError):
        autolabel.run_nnd(str(tmp_path))


def test_run_nnd_bookkeeping(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(autolabel, "cuda_nnd", _oracle_nnd)
    d = str(tmp_path / "ds")
    st = store.write_synthetic_dataset(d, n_scenes=3, n_frames=3, n_points=800, seed=9)
    ds = HDF5Dataset(d, store=st)
    scenes = list(ds.scene_id_bounds)
    # scene range [1, 2): only the middle scene is touched
    assert autolabel.run_nnd(d, scene_range=[1, 2], store=st, device="cpu") == 3
    for s, ts in ds.data_index:
        assert st.has(s, ts, "nnd") == (s == scenes[1])
    # the last frame of a scene is matched against its predecessor, every other one against its successor
    b = ds.scene_id_bounds[scenes[1]]
    norm = st.read(scenes[1], ds.data_index[b["min_index"]][1], "pose")
    for i in range(b["min_index"], b["max_index"] + 1):
        j = i - 1 if i == b["max_index"] else i + 1
        ts, ts1 = ds.data_index[i][1], ds.data_index[j][1]
        p0 = np.linalg.inv(norm) @ st.read(scenes[1], ts, "pose")
        p1 = np.linalg.inv(norm) @ st.read(scenes[1], ts1, "pose")
        ego = np.linalg.inv(p1) @ p0
        tr0 = (st.read(scenes[1], ts, "lidar")[:, :3] @ ego[:3, :3].T + ego[:3, 3]).astype(np.float32)
        ref = _oracle_nnd(torch.from_numpy(tr0), torch.from_numpy(np.ascontiguousarray(st.read(scenes[1], ts1, "lidar")[:, :3])))
        got = st.read(scenes[1], ts, "nnd")
        assert got.dtype == np.uint8 and got.shape == (tr0.shape[0],) and (got == ref).all()
    # overwrite=False skips finished scenes and fills in the rest
    assert autolabel.run_nnd(d, store=st, device="cpu", overwrite=False) == 6
    assert autolabel.run_nnd(d, store=st, device="cpu", overwrite=False) == 0
    # a larger threshold can only remove labels
    before = st.read(scenes[0], ds.data_index[0][1], "nnd").copy()
    autolabel.run_nnd(d, scene_range=[0, 1], store=st, device="cpu", min_nnd=0.32)
    after = st.read(scenes[0], ds.data_index[0][1], "nnd")
    assert (after <= before).all() and after.sum() < before.sum()


def _label_frame(seed, n=4000, n_clusters=40):
    rng = np.random.default_rng(seed)
    cluster = rng.integers(-1, n_clusters, n).astype(np.int16)       # <= 0: no cluster
    p_dufo, p_nnd = rng.random(n_clusters + 1), rng.random(n_clusters + 1) * 0.6
    c = np.clip(cluster, 0, None)
    return {"dufo": (rng.random(n) < p_dufo[c]).astype(np.uint8), "nnd": (rng.random(n) < p_nnd[c]).astype(np.uint8),
            "cluster": cluster, "dufocluster": cluster}


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_label_strategies_match_reference(seed):
    """seflow_auto / seflowpp_auto (OSF/src/autolabel.py:32-62) against the reference's own functions, plus hand cases."""
    fr = _label_frame(seed)
    ours = autolabel.seflowpp_auto(fr)
    assert ours.dtype == np.uint8 and set(np.unique(ours)) - {0} and (ours[fr["cluster"] <= 0] == 0).all()
    from oracle import ref_shims
    if ref_shims.reference_available():
        import importlib
        ref_shims.install()
        ref = importlib.import_module("src.autolabel")
        for kw in ({}, {"tau1": 0.01, "tau2": 0.05}, {"tau1": 0.4, "tau2": 0.9}):
            np.testing.assert_array_equal(autolabel.seflowpp_auto(fr, **kw), ref.seflowpp_auto(fr, **kw))
        np.testing.assert_array_equal(autolabel.seflow_auto(fr), ref.seflow_auto(fr))
        np.testing.assert_array_equal(autolabel.shift_cluster_id(fr["cluster"]), ref.shiftClusterid(fr["cluster"]))


def test_label_strategy_hand_case():
    # cluster ids 1..3 -> shifted 2..4.  id 2: dufo 2/4, nnd 1/4 -> dynamic; id 3: nnd 0 -> min fails; id 4: both 1/10 -> max fails
    cluster = np.array([1] * 4 + [2] * 4 + [3] * 10 + [0, -1], np.int16)
    dufo = np.array([1, 1, 0, 0] + [1, 1, 1, 1] + [1] + [0] * 9 + [1, 1], np.uint8)
    nnd = np.array([1, 0, 0, 0] + [0, 0, 0, 0] + [0, 1] + [0] * 8 + [1, 1], np.uint8)
    out = autolabel.seflowpp_auto({"dufo": dufo, "nnd": nnd, "cluster": cluster})
    assert out.tolist() == [2] * 4 + [0] * 4 + [0] * 10 + [0, 0]
    assert autolabel.seflow_auto({"dufo": dufo, "dufocluster": cluster}).tolist() == [2, 2, 0, 0, 3, 3, 3, 3, 4] + [0] * 9 + [0, 0]
    assert autolabel.seflowpp_auto({"dufo": dufo[:0], "nnd": nnd[:0], "cluster": cluster[:0]}).shape == (0,)
