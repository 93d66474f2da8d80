"""CPU: offline benchmark scorer (himo_b200/scoring.py): our ground-truth zip is read and scored by the reference's OWN
tools/test/score.py, and our `score()` returns the same numbers from the same two zips."""
import contextlib
import importlib.util
import io
import os

import numpy as np
import pytest

from himo_b200 import runner, scoring, store
from himo_b200.dataset import HDF5Dataset

REF_SCORE = "/root/reference/tools/test/score.py"


@pytest.fixture(scope="module")
def zips(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("bench") / "av2_synth")
    st = store.write_synthetic_dataset(d, n_scenes=2, n_frames=6, n_points=3000, seed=21)
    ds = HDF5Dataset(d, store=st)
    rng = np.random.default_rng(0)
    for scene, ts in ds.data_index:                      # a plausible estimate: ground truth + noise
        f = st.read(scene, ts, "flow")
        st.write(scene, ts, "est", (f + rng.normal(0, 0.05, f.shape)).astype(np.float32))
    pred = runner.run_save_zip({"data_dir": d, "res_name": "est"})
    gt = scoring.save_zip_gt(d, os.path.join(d, "gt"), res_name="flow", store=st)
    return d, gt, pred


def test_gt_zip_layout(zips):
    d, gt, pred = zips
    uu = scoring.list_sweep_uuids(gt)
    assert len(uu) > 0 and sorted(uu) == sorted(scoring.list_sweep_uuids(pred))
    comp, mask, cat, inst, speed, pc0 = scoring.read_data_file(gt, uu[0])
    assert comp.dtype == np.float32 and mask.dtype == bool and cat.dtype == np.uint8 and inst.dtype == np.uint32
    assert speed.dtype == np.float32 and pc0.shape == comp.shape and mask.any() and not mask.all()
    est, m2, c2, i2, s2, p2 = scoring.read_data_file(pred, uu[0])
    assert est.shape == comp.shape and m2.all() and c2 is None and i2 is None and s2 is None and p2 is None


def test_score_of_ground_truth_against_itself_is_zero(zips):
    _, gt, _ = zips
    s = scoring.score(gt, gt)
    assert s["num_instances"] > 0 and s["mpe"] == 0.0 and s["chamfer"] == 0.0


@pytest.mark.skipif(not os.path.exists(REF_SCORE), reason="needs /root/reference")
def test_score_equals_reference_scorer(zips, tmp_path):
    _, gt, pred = zips
    spec = importlib.util.spec_from_file_location("himo_ref_score", REF_SCORE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref.score(gt, pred, output_dir=str(tmp_path / "ref"))
    s = scoring.score(gt, pred, output_dir=str(tmp_path / "ours"))
    assert s["num_instances"] > 0 and s["mpe"] > 0
    for k in ("num_frames", "num_instances", "total_points", "car_num_objs", "car_num_pts", "others_num_objs", "others_num_pts"):
        assert s[k] == r[k], k
    # score.py takes the MPE on the float32 distance difference, eval.py (and InstanceMetrics) on the refined float32
    # points pc0 + distance: the two reference copies themselves agree to float32 rounding of the coordinates only
    tol = dict(rel=0, abs=1e-6)
    for k in ("mpe", "chamfer", "car_cde", "car_mpe", "others_cde", "others_mpe"):
        assert s[k] == pytest.approx(r[k], **tol), k
    for c in ("CAR", "OTHER_VEHICLES"):
        for k in ("mpe_mean", "mpe_std", "cham_mean", "cham_std"):
            assert s["per_category"][c][k] == pytest.approx(r["per_category"][c][k], **tol), (c, k)
        for rng_ in ("0-10", "10-20", "20-30", "30+"):
            for k in ("mpe", "cd", "num_pts", "num_obj"):
                assert s["per_category"][c]["velocity"][rng_][k] == pytest.approx(r["per_category"][c]["velocity"][rng_][k], **tol)
    assert os.path.exists(tmp_path / "ours" / "scores.json")


def test_score_reports_missing_and_mismatched_sweeps(zips, tmp_path):
    import shutil
    from zipfile import ZipFile
    _, gt, pred = zips
    ext = tmp_path / "pred_dir"
    with ZipFile(pred) as z:
        z.extractall(ext)
    feathers = sorted(ext.rglob("*.feather"))
    feathers[0].unlink()                                            # one missing prediction
    import pandas as pd
    df = pd.read_feather(feathers[1]); df.iloc[:-3].reset_index(drop=True).to_feather(feathers[1])     # one truncated
    s = scoring.score(gt, str(ext))
    assert len(s["missing_predictions"]) == 1 and len(s["point_count_mismatches"]) == 1
    full = scoring.score(gt, pred)
    assert s["num_frames"] == full["num_frames"] - 2
