"""Per-tile clock64() timeline of k_conv_umma for one layer (himo_conv_set_debug_buffer): where a tile's time goes.
usage: python scripts/trace_conv_tiles.py [groups]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from himo_b200 import conv, _lib
L = _lib.lib()
g = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cin = cout = int(sys.argv[3]) if len(sys.argv) > 3 else 128
act = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x = torch.randn(H, W, g * cin, device="cuda"); xp = conv.split_planes(x, 2)
w = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
ws = conv.weight_prescale(w, 2); wp = conv.pack_conv_weight(w, 2, ws).cuda()
out = torch.zeros((2, H, W, g * cout), dtype=torch.bfloat16, device="cuda")
bias = torch.randn(cout, device="cuda") * 0.1
def run():
    conv.conv2d_nhwc(xp, wp, bias, out, ksize=3, act=act, stride=1, cin=cin, n_groups=g, cin_group_stride=cin,
                     cout_group_stride=cout, acc_scale=1.0 / ws)
for _ in range(3): run()
dbg = torch.zeros((148, 32, 8), dtype=torch.int64, device="cuda")
L.himo_conv_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(); torch.cuda.synchronize()
L.himo_conv_set_debug_buffer(ctypes.c_void_p(0))
d = dbg.cpu().numpy()
names = ["mma:tile start", "mma:cross free", "mma:chunk0 issued", "mma:all issued", "epi:ready", "epi:mma done", "epi:drained", "epi:stored"]
for cta in (0, 2, 74):
    t0 = d[cta, 0, 0]
    if t0 == 0: continue
    print(f"CTA {cta}: cycles relative to its first tile start")
    for t in range(32):
        if d[cta, t, 0] == 0: break
        row = d[cta, t] - t0
        print(f"  tile {t:2d}: " + "  ".join(f"{n.split(':')[1]}={v}" for n, v in zip(names, row)))
n_tiles = (d[:, :, 0] != 0).sum(1)
lead = d[n_tiles > 0]
per_tile = []
for c in range(lead.shape[0]):
    k = (lead[c, :, 0] != 0).sum()
    if k >= 3:
        per_tile.append((lead[c, k - 1, 0] - lead[c, 1, 0]) / (k - 2))
print("median cycles between tile starts (steady state):", np.median(per_tile) if per_tile else None)
