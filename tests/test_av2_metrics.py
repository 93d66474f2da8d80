"""CPU: himo_b200.av2_metrics (OpenSceneFlow / AV2 evaluation metrics) pinned against the
reference's own OSF/src/utils/eval_metric.py + av2_eval.py (imported through oracle/ref_shims.py when
/root/reference exists) and against golden values those functions produced (tests/golden/av2_metrics_*.json)."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import av2_metrics as M
from oracle import ref_shims


def synth_frame(seed: int, n: int = 20000):
    """One frame of eval inputs: categories over the whole AV2 table, moving objects, NaN rows, invalid points."""
    rng = np.random.default_rng(seed)
    pc0 = np.stack([rng.uniform(-90, 90, n), rng.uniform(-90, 90, n), rng.uniform(-2, 4, n)], 1).astype(np.float32)
    ids = rng.integers(0, 31, n).astype(np.uint8)
    ids[rng.random(n) < 0.5] = 0                                   # half background
    rigid = np.tile(np.array([[0.9, 0.02, 0.0]], np.float32), (n, 1)) + rng.normal(0, 0.01, (n, 3)).astype(np.float32)
    speed = np.where(ids > 0, rng.choice([0.0, 0.03, 0.2, 0.8, 1.5, 2.4], n), 0.0).astype(np.float32)
    direc = rng.normal(size=(n, 3)).astype(np.float32)
    direc /= np.linalg.norm(direc, axis=1, keepdims=True)
    gt = rigid + direc * speed[:, None]
    est = gt + rng.normal(0, 0.04, (n, 3)).astype(np.float32)
    est[rng.choice(n, 50, replace=False)] = np.nan
    valid = rng.random(n) < 0.93
    t = lambda a: torch.from_numpy(a)
    return dict(est_flow=t(est), rigid_flow=t(rigid), pc0=t(pc0), gt_flow=t(gt.astype(np.float32)),
                is_valid=t(valid), pts_ids=t(ids))


def _split_key(v):
    return (v.name, float(v.thresholds_range[0]), float(v.thresholds_range[1]))


def _close(a, b):
    a, b = float(a), float(b)
    return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-12 * max(1.0, abs(b))


def _assert_splits_equal(ours, ref):
    ours, ref = {_split_key(v): v for v in ours}, {_split_key(v): v for v in ref}
    assert ours.keys() == ref.keys()
    for k in ref:
        assert int(ours[k].count) == int(ref[k].count)
        assert _close(ours[k].avg_epe, ref[k].avg_epe) and _close(ours[k].avg_range, ref[k].avg_range), k


def _summary(metrics) -> dict:
    """normalised OfficialMetrics state as plain floats (shared by both implementations)."""
    metrics.normalize()
    f = lambda v: None if (isinstance(v, list) or v is None) else float(v)
    return {"epe_3way": {k: f(v) for k, v in metrics.epe_3way.items()},
            "bucketed": {k: {s: f(v[s]) for s in ("Static", "Dynamic")} for k, v in metrics.bucketed.items()},
            "ssf": {k: {s: f(v[s]) for s in ("Static", "Dynamic", "#Static", "#Dynamic")} for k, v in metrics.epe_ssf.items()}}


def _assert_summary_equal(a, b):
    assert a.keys() == b.keys()
    for sec in a:
        assert a[sec].keys() == b[sec].keys(), sec
        for k in a[sec]:
            va, vb = a[sec][k], b[sec][k]
            if isinstance(va, dict):
                for s in va:
                    assert (va[s] is None and vb[s] is None) or _close(va[s], vb[s]), (sec, k, s, va[s], vb[s])
            else:
                assert (va is None and vb is None) or _close(va, vb), (sec, k, va, vb)


def _run(mod, seeds):
    om = mod.OfficialMetrics()
    for s in seeds:
        fr = synth_frame(s)
        om.step(mod.evaluate_leaderboard(**fr), mod.evaluate_leaderboard_v2(**fr), mod.evaluate_ssf(**fr))
    return om


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_per_frame_metrics_match_live_reference():
    ref = ref_shims.import_eval_metric()
    for seed in (1, 2):
        fr = synth_frame(seed)
        a, b = M.evaluate_leaderboard(**fr), ref.evaluate_leaderboard(**fr)
        assert a.keys() == b.keys() and all(_close(a[k], b[k]) for k in b), (a, b)
        _assert_splits_equal(M.evaluate_leaderboard_v2(**fr), ref.evaluate_leaderboard_v2(**fr))
        _assert_splits_equal(M.evaluate_ssf(**fr), ref.evaluate_ssf(**fr))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_official_metrics_match_live_reference():
    ref = ref_shims.import_eval_metric()
    _assert_summary_equal(_summary(_run(M, (3, 4, 5))), _summary(_run(ref, (3, 4, 5))))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "av2_metrics_*.json"))))
def test_official_metrics_match_golden(path):
    g = json.load(open(path))
    _assert_summary_equal(_summary(_run(M, g["seeds"])), g["summary"])


def test_rank_merge_equals_single_pass():
    """Scene-sharded evaluation: two ranks' un-normalised states merged == one pass over all frames."""
    one = _summary(_run(M, (6, 7, 8, 9)))
    a, b = _run(M, (6, 7)), _run(M, (8, 9))
    a.merge(b)
    _assert_summary_equal(_summary(a), one)


def test_empty_and_all_invalid_frames():
    fr = synth_frame(10, 500)
    fr["is_valid"] = torch.zeros(500, dtype=torch.bool)
    r = M.evaluate_leaderboard(**fr)
    assert r == {"EPE_BS": 0.0, "EPE_FD": 0.0, "EPE_FS": 0.0, "IoU": 0.0}
    assert [v for v in M.evaluate_leaderboard_v2(**fr) if v.name != "BACKGROUND"] == []
    assert M.evaluate_ssf(**fr) == []
