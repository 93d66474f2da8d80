"""FastNSF (H3) timing: iterations/s of the device optimisation loop on a 100k-point pair, next to the
CPU oracle (torch autograd, all host threads)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import fastnsf, frames, weights
from oracle import fastnsf_ref
from oracle.deflowpp_ref import pose0to1

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tr = frames.lidar_triple(n, 5)
pc0, pc1 = torch.from_numpy(tr["pc0"]), torch.from_numpy(tr["pc1"])
T = pose0to1(torch.from_numpy(tr["pose0"]), torch.from_numpy(tr["pose1"]))
sel0 = pc0[fastnsf_ref.range_mask(pc0)]
tr0 = (sel0 @ T[:3, :3].T + T[:3, 3]).contiguous()
sel1 = pc1[fastnsf_ref.range_mask(pc1)].contiguous()
sd = weights.synth_neural_prior_state_dict(1)
res = {"n_points": int(tr0.shape[0])}
for prec in ("fp32", "bf16"):
    net = fastnsf.FastNSF(itr_num=iters, early_patience=10 ** 6, precision=prec)
    d0, d1 = tr0.cuda(), sel1.cuda()
    torch.cuda.synchronize(); t = time.perf_counter()
    lo, dims = fastnsf.volume_geometry(d0, d1, 10.0)
    D = fastnsf.dt_build(d1, lo, dims, 10.0)
    torch.cuda.synchronize(); t_dt = time.perf_counter() - t
    net.optimize(d0, d1, init_state_dict=sd, D=D, lo=lo, dims=dims)          # warm-up
    torch.cuda.synchronize(); t = time.perf_counter()
    out = net.optimize(d0, d1, init_state_dict=sd, D=D, lo=lo, dims=dims)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    res[prec] = {"dt_build_ms": 1e3 * t_dt, "dims": list(dims), "ms_per_iter": 1e3 * dt / out["iterations"],
                 "iterations": out["iterations"], "loss": out["loss"]}
torch.set_num_threads(os.cpu_count())
t = time.perf_counter()
ref = fastnsf_ref.optimize(sd, tr0, sel1, itr_num=3, patience=10 ** 6, Dvol=D.cpu())
res["cpu"] = {"ms_per_iter": 1e3 * (time.perf_counter() - t) / 3, "cores": os.cpu_count()}
print(json.dumps(res))
