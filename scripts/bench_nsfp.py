"""NSFP iteration time on a lidar-shaped pair: total ms/iteration and the share spent in the four 1-NN searches +
gradient scatters (our kernels) versus the torch MLP/Adam part.  python scripts/bench_nsfp.py [n_points] [iters]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from himo_b200 import chamfer3d, frames, nsfp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tr = frames.lidar_triple(n, 3)
pc0 = torch.from_numpy(tr["pc0"][:, :3].copy()).cuda()
pc1 = torch.from_numpy(tr["pc1"][:, :3].copy()).cuda()
m = nsfp.NSFP(itr_num=5, early_patience=0)
torch.manual_seed(0)
m.optimize(pc0, pc1)                                  # warm-up
m.iteration_num = iters
torch.cuda.synchronize(); t0 = time.perf_counter()
torch.manual_seed(0)
r = m.optimize(pc0, pc1)
torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) / r["iterations"] * 1e3

ch = chamfer3d.nnChamferDis()
a = (pc0 + 0.05 * torch.randn_like(pc0)).requires_grad_(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    (ch.truncated_dis(a, pc1) + ch.truncated_dis(a, pc0)).backward()
e0.record()
for _ in range(20):
    (ch.truncated_dis(a, pc1) + ch.truncated_dis(a, pc0)).backward()
e1.record(); torch.cuda.synchronize()
t_ch = e0.elapsed_time(e1) / 20
print(json.dumps({"bench": "nsfp_iteration", "n0": pc0.shape[0], "n1": pc1.shape[0], "iterations": r["iterations"],
                  "ms_per_iteration": round(t_all, 3), "chamfer_loss_fwd_bwd_ms": round(t_ch, 3), "best_loss": r["loss"]}))
