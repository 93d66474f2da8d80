"""GPU: the public engine API -- synchronous `infer` and the pipelined `infer_stream` give identical
results, with and without ground points."""
import numpy as np
import pytest

from himo_b200 import frames, weights
from himo_b200.engine import SeFlowPPEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_slots", [1, 2, 3])
def test_stream_equals_sync_and_handles_ground(n_slots):
    """n_slots networks in flight on their own streams and workspaces (shared weights): results are those of the
    one-at-a-time call, bit for bit."""
    eng = SeFlowPPEngine(weights.synth_deflowpp_state_dict(1), max_points=8192, n_slots=n_slots)
    fr = []
    for k in range(8):
        tr = frames.lidar_triple(3000 + 200 * k, 40 + k)
        f = {key: tr[key] for key in ("pc0", "pc1", "pch1", "pose0", "pose1", "poseh1")}
        if k % 2:
            rng = np.random.default_rng(k)
            f["gm0"] = rng.random(f["pc0"].shape[0]) < 0.2
            f["gm1"] = rng.random(f["pc1"].shape[0]) < 0.2
        fr.append(f)
    sync = [eng.infer(f) for f in fr]
    stream = list(eng.infer_stream(iter(fr)))
    assert len(stream) == len(sync)
    for a, b in zip(sync, stream):
        assert a.shape == b.shape
        np.testing.assert_array_equal(a, b)
    # ground points carry pose flow only
    g = fr[1]["gm0"]
    assert np.abs(sync[1][g]).max() < 5.0
