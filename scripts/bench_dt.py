"""Raster distance-transform passes at the bench volume (1040 x 1030 x 52): tiled multi-launch passes vs the cluster sweep."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from himo_b200 import _lib, fastnsf
L = _lib.lib()
dims = (1040, 1030, 52)
dims_c = (ctypes.c_int32 * 3)(*dims)
g = torch.Generator().manual_seed(0)
D0 = torch.full(dims, 1e10, dtype=torch.float32)
idx = torch.randint(0, D0.numel(), (90000,), generator=g)
D0.view(-1)[idx] = 0
D0 = D0.cuda()
out = {}
res = {}
for sweep in (0, 1):
    D = D0.clone()
    for axis in (0, 1):
        for dr in (1, -1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            st = L.himo_nsf_dt_pass(_lib.ptr(D), dims_c, ctypes.c_float(10.0), axis, dr, sweep, _lib.stream_ptr(D.device))
            e1.record(); torch.cuda.synchronize()
            assert st == 0, st
            out[f"{'sweep' if sweep else 'tiled'}_axis{axis}_dir{dr}_ms"] = round(e0.elapsed_time(e1), 3)
    res[sweep] = D
for big in (0, 1, 2, 3):
    L.himo_nsf_set_dt_big_tiles(big)
    D = D0.clone()
    ts = []
    for dr in (1, -1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        L.himo_nsf_dt_pass(_lib.ptr(D), dims_c, ctypes.c_float(10.0), 2, dr, 0, _lib.stream_ptr(D.device))
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    out[f"axis2_variant{big}_ms"] = ts
    res[10 + big] = D
out["axis2_identical"] = bool(torch.equal(res[10], res[11]) and torch.equal(res[10], res[12]) and torch.equal(res[10], res[13]))
L.himo_nsf_set_dt_big_tiles(0)
dbg = torch.zeros(16 * 4 * 4, dtype=torch.int64, device="cuda")
L.himo_nsf_set_dt_debug_buffer.argtypes = [ctypes.c_void_p]
L.himo_nsf_set_dt_debug_buffer(dbg.data_ptr())
D = D0.clone()
L.himo_nsf_dt_pass(_lib.ptr(D), dims_c, ctypes.c_float(10.0), 0, 1, 1, _lib.stream_ptr(D.device))
torch.cuda.synchronize()
L.himo_nsf_set_dt_debug_buffer(None)
t = dbg.view(16, 4, 4).cpu().double() / (dims[0] - 1)
out["cycles_per_step_cta0_probe0"] = dict(zip(["cp_wait", "halo_wait", "compute", "syncthreads"], [round(float(v)) for v in t[0, 0]]))
out["cycles_per_step_mean"] = dict(zip(["cp_wait", "halo_wait", "compute", "syncthreads"], [round(float(v)) for v in t.mean((0, 1))]))
out["cycles_per_step_max_over_cta"] = dict(zip(["cp_wait", "halo_wait", "compute", "syncthreads"], [round(float(v)) for v in t.amax((0, 1))]))
out["identical"] = bool(torch.equal(res[0], res[1]))
print(json.dumps(out))
