"""Debug helper: per-stage error report of the fused SeFlow++ path vs the CPU oracle."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from himo_b200 import deflowpp, frames, weights
from oracle import deflowpp_ref

kind, n, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4 and sys.argv[4] != '-':
    from himo_b200 import _lib
    _lib.lib().himo_conv_set_flush_iters(int(sys.argv[4]))
    print("flush_iters", sys.argv[4])
sd = weights.synth_deflowpp_state_dict(seed)
tr = frames.lidar_triple(n, seed) if kind == "lidar" else frames.uniform_triple(n, seed)
ref = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"], tr["pose1"],
                                    accum="exact", keep=True)
net = deflowpp.DeFlowPP(max_points=max(n, 4096)).load_state_dict(sd)
b = {k: torch.from_numpy(tr[k])[None].cuda() for k in ("pc0", "pc1", "pch1")}
b.update({k: [torch.from_numpy(tr[k])] for k in ("pose0", "pose1", "poseh1")})
out = net(b)
torch.cuda.synchronize()
ws = net._ws
v = net.views()
def f32(ptr, cnt):
    off = ptr - ws.data_ptr()
    return ws[off:off + cnt * 4].view(torch.float32).cpu()
V = f32(v["V"], 512 * 512 * 96).view(512, 512, 96)
after = ref["after"].permute(1, 2, 0)
print("V   max|ref| %.3f  max err %.3e  rms err %.3e" % (after.abs().max(), (V - after).abs().max(), (V - after).pow(2).mean().sqrt()))
fl, rf = out["flow"][0].cpu(), ref["flow"]
e = (fl - rf).abs().max(1).values
print("flow max|ref| %.3f  max err %.3e  rms %.3e  n>1e-4: %d of %d" % (rf.abs().max(), e.max(), e.pow(2).mean().sqrt(), (e > 1e-4).sum(), e.numel()))
idx = e.argmax().item()
print("worst point", idx, fl[idx].tolist(), rf[idx].tolist())
# decoder alone on oracle inputs: feed the reference 'after' through the oracle decoder with OUR V
co = ref["info_0"]["voxel_coords"]
fl2 = deflowpp_ref.gru_decoder(sd, ref["before"], V.permute(2, 0, 1).contiguous(), ref["info_0"]["point_offsets"], co)
e2 = (fl2 - rf).abs().max(1).values
print("oracle decoder on OUR V: max err %.3e" % e2.max())
e3 = (fl - fl2).abs().max(1).values
print("our decoder vs oracle decoder on OUR V: max err %.3e" % e3.max())
