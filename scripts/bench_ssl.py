"""Self-supervision consumers of the 1-NN kernel at dataset scale: the `nnd` labelling rule per frame pair and the
SeFlow / SeFlow++ losses (forward + backward to the flow) on a lidar-shaped triple.  python scripts/bench_ssl.py [n]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from himo_b200 import autolabel, frames, lossfuncs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
tr = frames.lidar_triple(n, 5)
pc0, pc1, pch1 = (torch.from_numpy(tr[k][:, :3].copy()).cuda() for k in ("pc0", "pc1", "pch1"))


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_nnd = timed(lambda: autolabel.cuda_nnd(pc0, pc1))            # includes the uint8 labels' device->host copy
lab = autolabel.cuda_nnd(pc0, pc1)

# cluster labels: moving points grouped on a 4 m grid (a few hundred clusters), label 1 for isolated movers
rng = np.random.default_rng(0)
def labels(pc, moving):
    cell = torch.floor(pc[:, :2] / 4.0).long()
    key = (cell[:, 0] + 64) * 128 + (cell[:, 1] + 64)
    _, inv = torch.unique(key, return_inverse=True)
    l = torch.where(moving, inv + 2, torch.zeros_like(inv))
    return l
m0 = torch.from_numpy(lab.astype(bool)).cuda()
l0 = labels(pc0, m0)
l1 = labels(pc1, torch.from_numpy(autolabel.cuda_nnd(pc1, pc0).astype(bool)).cuda())
lh = labels(pch1, torch.from_numpy(autolabel.cuda_nnd(pch1, pc0).astype(bool)).cuda())
est = (0.05 * torch.randn_like(pc0)).requires_grad_(True)
d = {"pc0": pc0, "pc1": pc1, "pch1": pch1, "est_flow": est, "pc0_labels": l0, "pc1_labels": l1, "pch1_labels": lh}


def step(fn):
    out = fn(d)
    sum(out.values()).backward()
    est.grad = None


res = {"bench": "ssl_consumers", "n_points": int(pc0.shape[0]), "moving_fraction": float(lab.mean()),
       "clusters": int(torch.unique(l0).numel()), "nnd_label_ms_per_frame": round(t_nnd, 3),
       "seflow_loss_fwd_bwd_ms": round(timed(lambda: step(lossfuncs.seflowLoss), reps=10), 3),
       "seflowpp_loss_fwd_bwd_ms": round(timed(lambda: step(lossfuncs.seflowppLoss), reps=10), 3)}
print(json.dumps(res))
