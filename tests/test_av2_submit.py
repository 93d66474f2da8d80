"""CPU: AV2 leaderboard submission files (himo_b200/av2_submit.py) against the reference's OWN writers
(OSF/src/utils/av2_eval.py:758-801, OSF/src/utils/mics.py:312-344) and through the `data_mode=test` driver."""
import io
import json
import os
from pathlib import Path
from zipfile import ZipFile

import numpy as np
import pandas as pd
import pytest

from himo_b200 import av2_submit, runner, store
from oracle import ref_shims


def _sweeps(seed=0, logs=("logB", "logA"), per_log=3, n=500):
    rng = np.random.default_rng(seed)
    out = []
    for log in logs:
        for k in range(per_log):
            flow = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
            out.append((flow, rng.random(n) < 0.3, (log, 315969904359876000 + k * 100000000)))
    return out


def _zip_tables(path):
    with ZipFile(path) as z:
        names = z.namelist()
        return names, {n: (pd.read_feather(io.BytesIO(z.read(n))) if n.endswith(".feather") else json.loads(z.read(n))) for n in names}


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("version", [1, 2])
def test_files_and_zip_equal_the_reference_writers(tmp_path, version):
    ref_write, ref_zip = ref_shims.import_submit_writers()
    a, b = tmp_path / "ours", tmp_path / "ref"
    for flow, dyn, uuid in _sweeps():
        av2_submit.write_output_file(flow, dyn, uuid, a, leaderboard_version=version)
        ref_write(flow, dyn, uuid, Path(b), leaderboard_version=version)
    za = av2_submit.zip_res(a, output_file=str(tmp_path / "ours.zip"), leaderboard_version=version, is_supervised=False)
    zb = ref_zip(str(b), output_file=str(tmp_path / "ref.zip"), leaderboard_version=version, is_supervised=False)
    na, ta = _zip_tables(za)
    nb, tb = _zip_tables(zb)
    assert sorted(na) == sorted(nb)
    for n in na:
        if n.endswith(".feather"):
            assert list(ta[n].columns) == list(tb[n].columns) and (ta[n].dtypes == tb[n].dtypes).all()
            pd.testing.assert_frame_equal(ta[n], tb[n], check_exact=True)
        else:
            assert ta[n] == tb[n]


def test_version_2_renames_sweeps_in_timestamp_order(tmp_path):
    for flow, dyn, uuid in _sweeps(per_log=3):
        av2_submit.write_output_file(flow, dyn, uuid, tmp_path / "r", leaderboard_version=2)
    names, tabs = _zip_tables(av2_submit.zip_res(tmp_path / "r", output_file=str(tmp_path / "s.zip"), leaderboard_version=2, is_supervised=True))
    assert names[0] == "metadata.json" and tabs["metadata.json"] == {"Is Supervised?": True}
    assert sorted(n for n in names if n.startswith("logA/")) == [f"logA/{k:010d}.feather" for k in (0, 5, 10)]
    t = tabs["logA/0000000005.feather"]
    assert list(t.columns) == ["is_valid", "flow_tx_m", "flow_ty_m", "flow_tz_m"] and t["flow_tx_m"].dtype == np.float16
    with pytest.raises(ValueError):
        av2_submit.write_output_file(np.zeros((1, 3)), np.zeros(1, bool), ("l", 1), tmp_path, leaderboard_version=3)


@pytest.mark.parametrize("version", [1, 2])
def test_test_mode_driver_writes_the_submission(tmp_path, version):
    d = str(tmp_path / "data" / "av2_synth")
    st = store.write_synthetic_dataset(d, n_scenes=2, n_frames=5, n_points=1200, seed=7)
    out = runner.run_validate({"dataset_path": d, "model": "stored", "res_name": "flow", "data_mode": "test",
                               "leaderboard_version": str(version), "output": "sub", "supervised_flag": "false"})
    assert out == os.path.join(str(tmp_path / "data"), "results", "sub.zip") and os.path.exists(out)
    names, tabs = _zip_tables(out)
    from himo_b200.dataset import HDF5Dataset
    from himo_b200 import himo
    ds = HDF5Dataset(d, eval=True, store=st)
    feathers = [n for n in names if n.endswith(".feather")]
    assert len(feathers) == len(ds) > 0
    item = ds[0]
    pf = himo.pose_flow_np(item["pc0"][:, :3], item["pose0"], item["pose1"]).astype(np.float32)
    m = np.asarray(item["eval_mask"], bool)
    if version == 1:
        t = tabs[f"{item['scene_id']}/{item['timestamp']}.feather"]
        np.testing.assert_array_equal(t["flow_tx_m"].values, item["flow"][m, 0].astype(np.float16))
        np.testing.assert_array_equal(t["is_dynamic"].values, np.linalg.norm(item["flow"][m] - pf[m], axis=1) >= 0.05)
    else:
        assert tabs["metadata.json"] == {"Is Supervised?": False}
        first = sorted(n for n in feathers if n.startswith(item["scene_id"] + "/"))[0]
        assert first.endswith("0000000000.feather") and len(tabs[first]) == item["pc0"].shape[0]
