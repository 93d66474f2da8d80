"""Where does a FastNSF pair's wall time go in the engine (save.py model=fastnsf)?  Sequential vs streamed, with a
breakdown of the sequential form."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from himo_b200.engine import FastNSFEngine  # noqa: E402

dev = torch.device("cuda", 0)
frame = bench.make_frames(0, 1)[0]
out = {}
for mode in ("sequential", "stream", "breakdown"):
    eng = FastNSFEngine(device=dev, itr_num=5000, early_patience=10)
    eng.infer(frame)
    torch.cuda.synchronize()
    its, t0 = [], time.perf_counter()
    if mode == "sequential":
        for _ in range(6):
            eng.infer(frame); its.append(eng.net.last_info["iterations"])
    elif mode == "stream":
        for _ in eng.infer_stream(frame for _ in range(6)):
            its.append(eng.net.last_info["iterations"])
    else:
        prep_s = fin_s = 0.0
        for _ in range(6):
            a = time.perf_counter()
            p = eng._prepare(frame); torch.cuda.synchronize()
            b = time.perf_counter()
            eng._finish(p); torch.cuda.synchronize()
            c = time.perf_counter()
            prep_s += b - a; fin_s += c - b
            its.append(eng.net.last_info["iterations"])
        out["prepare_ms_per_pair"] = prep_s / 6 * 1e3
        out["finish_ms_per_pair"] = fin_s / 6 * 1e3
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[mode] = {"ms_per_pair": dt / 6 * 1e3, "iterations": its, "ms_per_iteration_all_in": dt * 1e3 / sum(its)}
print(json.dumps(out))
