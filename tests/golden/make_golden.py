"""Generate tests/golden/* in the BUILD CONTAINER (needs /root/reference; never run on the GPU box).

    python tests/golden/make_golden.py

What is pinned by what
  av2_fixture_clouds.npz   xyz of the reference's own test clouds OSF/assets/tests/test_pc{0,1}.npy
                           (fp16-quantised in the reference, stored losslessly as float16) + the
                           reference's recorded known answer, Chamfer loss 0.1710
                           (OSF/assets/tests/chamferdis_speed_test.py:113-126)
  deflowpp_n*.npz          outputs of the reference's OWN `src.models.DeFlowPP` class (imported from
                           /root/reference through oracle/ref_shims.py) on seeded synthetic triples with
                           himo_b200.weights.synth_deflowpp_state_dict(seed) loaded strictly
  fastnsf_n*.npz           output of the reference's OWN `src.models.FastNSF` class (its Neural_Prior, EarlyStopping,
                           Adam, grid_sample lookup) with the initial weights it drew, on a cropped synthetic pair;
                           the FastGeodis call is served by the restated transform (parity unpinned for that piece)
  nsfp_n*.npz              output of the reference's OWN `src.models.NSFP` class with the network it drew, on a cropped synthetic pair
  seflow_loss_*.npz        loss terms + flow gradient of the reference's OWN OSF/src/lossfuncs/selfsupervise.py on seeded frames
  av2_metrics_*.json       normalised OfficialMetrics of the reference's OWN OSF/src/utils/eval_metric.py on seeded frames
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from himo_b200 import frames, weights  # noqa: E402
from oracle import ref_shims  # noqa: E402


def fixture_clouds():
    base = os.path.join(ref_shims.OSF_ROOT, "assets", "tests")
    pc0 = np.load(os.path.join(base, "test_pc0.npy"))[:, :3]
    pc1 = np.load(os.path.join(base, "test_pc1.npy"))[:, :3]
    assert (pc0.astype(np.float16).astype(np.float32) == pc0).all()
    assert (pc1.astype(np.float16).astype(np.float32) == pc1).all()
    np.savez_compressed(os.path.join(HERE, "av2_fixture_clouds.npz"), pc0=pc0.astype(np.float16),
                        pc1=pc1.astype(np.float16), chamfer_known_answer=np.float32(0.1710))


def deflowpp(models, n, seed, kind):
    sd = weights.synth_deflowpp_state_dict(seed)
    net = models.DeFlowPP().eval()
    net.load_state_dict(sd, strict=True)
    tr = frames.lidar_triple(n, seed) if kind == "lidar" else frames.uniform_triple(n, seed)
    if kind == "uniform":  # exercise the NaN-padding path of DynamicVoxelizer (encoder.py:576-578)
        tr["pc0"][5::97] = np.nan
    batch = {k: torch.from_numpy(tr[k])[None] for k in ("pc0", "pc1", "pch1")}
    batch.update({k: [torch.from_numpy(tr[k])] for k in ("pose0", "pose1", "poseh1")})
    torch.set_num_threads(1)  # fixed reduction order inside the CPU convolutions
    with torch.no_grad():
        out = net(batch)
    np.savez_compressed(
        os.path.join(HERE, f"deflowpp_{kind}_n{n}_s{seed}.npz"),
        pc0=tr["pc0"], pc1=tr["pc1"], pch1=tr["pch1"], pose0=tr["pose0"], pose1=tr["pose1"],
        poseh1=tr["poseh1"], weight_seed=np.int64(seed),
        flow=out["flow"][0].numpy(), pose_flow=out["pose_flow"][0].numpy(),
        pc0_valid_point_idxes=out["pc0_valid_point_idxes"][0].numpy(),
        pc1_valid_point_idxes=out["pc1_valid_point_idxes"][0].numpy(),
        pch1_valid_point_idxes=out["pch1_valid_point_idxes"][0].numpy())


def fastnsf(models, n, seed, itr_num, patience, half_extent):
    """The reference's OWN `src.models.FastNSF` (fastnsf.py:83-222; FastGeodis replaced by the restated transform of
    oracle/ref_shims.py -- the one unpinned piece) on a cropped synthetic pair.  The initial weights are whatever
    `Neural_Prior()` + `init_weights()` draw from the global torch RNG after torch.manual_seed(seed) (fastnsf.py:110-113);
    they are stored so that the oracle and the CUDA path can start from exactly the same network."""
    import importlib
    npm = importlib.import_module("src.models.basic.nsfp_module")
    tr = frames.lidar_triple(n, seed)
    crop = lambda a: np.ascontiguousarray(a[(np.abs(a[:, 0]) < half_extent) & (np.abs(a[:, 1]) < half_extent)])
    pc0, pc1 = crop(tr["pc0"]), crop(tr["pc1"])
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    net = npm.Neural_Prior(filter_size=128, act_fn="relu", layer_size=8)
    net.init_weights()
    sd0 = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    model = models.FastNSF(itr_num=itr_num, early_patience=patience)     # conf/model/fastnsf.yaml:13: patience 10
    torch.manual_seed(seed)                                              # optimize() re-draws the same network
    batch = {"pc0": [torch.from_numpy(pc0)], "pc1": [torch.from_numpy(pc1)],
             "pose0": [torch.from_numpy(tr["pose0"])], "pose1": [torch.from_numpy(tr["pose1"])]}
    out = model(batch)
    np.savez_compressed(
        os.path.join(HERE, f"fastnsf_n{pc0.shape[0]}_s{seed}_k{itr_num}.npz"),
        pc0=pc0, pc1=pc1, pose0=tr["pose0"], pose1=tr["pose1"], itr_num=np.int64(itr_num), patience=np.int64(patience),
        flow=out["flow"][0].detach().numpy(), pose_flow=out["pose_flow"][0].detach().numpy(),
        **{"w::" + k: v for k, v in sd0.items()})


def nsfp(models, n, seed, itr_num, patience, half_extent):
    """The reference's OWN `src.models.NSFP` (nsfp.py:28-186, chamfer3D served by the brute-force shim) on a cropped
    synthetic pair; the default-initialised network it draws after torch.manual_seed(seed) is stored alongside."""
    import importlib
    npm = importlib.import_module("src.models.basic.nsfp_module")
    tr = frames.lidar_triple(n, seed)
    crop = lambda a: np.ascontiguousarray(a[(np.abs(a[:, 0]) < half_extent) & (np.abs(a[:, 1]) < half_extent)])
    pc0, pc1 = crop(tr["pc0"]), crop(tr["pc1"])
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    sd0 = {k: v.detach().clone().numpy() for k, v in npm.Neural_Prior(filter_size=128, act_fn="relu", layer_size=8).state_dict().items()}
    model = models.NSFP(itr_num=itr_num, early_patience=patience)
    torch.manual_seed(seed)
    batch = {"pc0": [torch.from_numpy(pc0)], "pc1": [torch.from_numpy(pc1)],
             "pose0": [torch.from_numpy(tr["pose0"])], "pose1": [torch.from_numpy(tr["pose1"])]}
    out = model(batch)
    np.savez_compressed(
        os.path.join(HERE, f"nsfp_n{pc0.shape[0]}_s{seed}_k{itr_num}.npz"),
        pc0=pc0, pc1=pc1, pose0=tr["pose0"], pose1=tr["pose1"], itr_num=np.int64(itr_num), patience=np.int64(patience),
        flow=out["flow"][0].detach().numpy(), pose_flow=out["pose_flow"][0].detach().numpy(),
        **{"w::" + k: v for k, v in sd0.items()})


def av2_metrics(seeds):
    """Normalised `OfficialMetrics` state of the reference's OWN OSF/src/utils/eval_metric.py (three-way EPE, bucketed
    normalised EPE, range-wise SSF EPE) over seeded synthetic frames (tests/test_av2_metrics.py::synth_frame)."""
    import json
    sys.path.insert(0, os.path.dirname(HERE))
    import test_av2_metrics as T
    ref = ref_shims.import_eval_metric()
    summary = T._summary(T._run(ref, seeds))
    name = "av2_metrics_s" + "_".join(str(s) for s in seeds) + ".json"
    json.dump({"seeds": list(seeds), "generator": "tests/golden/make_golden.py::av2_metrics (reference eval_metric.py)",
               "summary": summary}, open(os.path.join(HERE, name), "w"), indent=1)


def seflow_losses(cases):
    """Terms and d(sum of terms)/d(est_flow) of the reference's OWN seflowLoss / seflowppLoss
    (OSF/src/lossfuncs/selfsupervise.py:25-190, chamfer3D served by the brute-force shim) on the seeded clustered
    frames of tests/test_lossfuncs.py::synth_loss_frame."""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_lossfuncs as T
    ref = ref_shims.import_lossfuncs()
    for seed, n, frac in cases:
        frame = T.synth_loss_frame(seed, n, dynamic_fraction=frac)
        out = {"seed": np.int64(seed), "n": np.int64(n), "frac": np.float64(frac)}
        for name in ("seflowLoss", "seflowppLoss"):
            v, g = T.run_loss(getattr(ref, name), frame)
            out[name + "_terms"] = np.array([v[k] for k in T.TERMS], np.float64)
            out[name + "_grad"] = g.astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"seflow_loss_s{seed}_n{n}.npz"), **out)


def demo_index_shape():
    """Per-scene row counts of the reference's AV2 demo indices (assets/docs/av2/index_{total,eval}.pkl): config-4 shapes."""
    import collections
    import json
    import pickle
    base = os.path.join(ref_shims.REFERENCE_ROOT, "assets", "docs", "av2")
    t = pickle.load(open(os.path.join(base, "index_total.pkl"), "rb"))
    e = pickle.load(open(os.path.join(base, "index_eval.pkl"), "rb"))
    c, ce = collections.Counter(s for s, _ in t), collections.Counter(s for s, _ in e)
    json.dump({"source": "/root/reference/assets/docs/av2/index_{total,eval}.pkl", "total_rows": len(t), "eval_rows": len(e),
               "frames_per_scene": dict(sorted(c.items())), "eval_frames_per_scene": dict(sorted(ce.items()))},
              open(os.path.join(HERE, "av2_demo_index_shape.json"), "w"), indent=1)


def main():
    assert ref_shims.reference_available(), "needs /root/reference"
    fixture_clouds()
    models = ref_shims.import_models()
    deflowpp(models, 2000, 11, "lidar")
    deflowpp(models, 3000, 12, "uniform")
    fastnsf(models, 12000, 13, 15, 10, 12.0)
    fastnsf(models, 12000, 13, 3, 10, 12.0)     # short horizon: before the optimiser's chaotic divergence sets in
    nsfp(models, 6000, 14, 3, 30, 12.0)
    nsfp(models, 6000, 14, 12, 30, 12.0)
    av2_metrics([11, 12, 13])
    demo_index_shape()
    seflow_losses([(31, 2400, 0.35), (33, 900, 0.5)])
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".npz"):
            print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
