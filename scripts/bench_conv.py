"""Per-layer timing of the tcgen05 convolution (warm L2 for weights, activations > L2 where large)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from himo_b200 import conv, _lib

LAYERS = [  # name, H, W, Cin, Cout, ksize, stride, groups, act (1 = bias + GELU, as the encoder), fp32 out
    ("enc1.0 32->64 s2 @512 x3", 512, 512, 32, 64, 3, 2, 3, 1, 0),
    ("enc1.1 64->64 @256 x3", 256, 256, 64, 64, 3, 1, 3, 1, 0),
    ("enc2.1 128->128 @128 x3", 128, 128, 128, 128, 3, 1, 3, 1, 0),
    ("enc2.1 no-act (epilogue probe)", 128, 128, 128, 128, 3, 1, 3, 0, 0),
    ("enc2.1 no-act fp32 out (probe)", 128, 128, 128, 128, 3, 1, 3, 0, 1),
    ("enc3.1 256->256 @64 x3", 64, 64, 256, 256, 3, 1, 3, 1, 0),
    ("b1.u3 1x1 384->384 @128", 128, 128, 384, 384, 1, 1, 1, 0, 0),
    ("b1.u4 768->384 @128", 128, 128, 768, 384, 3, 1, 1, 0, 0),
    ("b2.u4 384->192 @256", 256, 256, 384, 192, 3, 1, 1, 0, 0),
    ("b3.u4 192->96 @512", 512, 512, 192, 96, 3, 1, 1, 0, 0),
    ("b3.u5 96->96 @512", 512, 512, 96, 96, 3, 1, 1, 0, 0),
    ("b2.u5 192->192 @256", 256, 256, 192, 192, 3, 1, 1, 0, 0),
    ("dec4 96->96 @512 fp32 out", 512, 512, 96, 96, 3, 1, 1, 0, 1),
    ("probe enc2.1 x6 frames", 128, 128, 128, 128, 3, 1, 6, 1, 0),
    ("probe enc2.1 x12 frames", 128, 128, 128, 128, 3, 1, 12, 1, 0),
    ("probe enc2.1 x24 frames", 128, 128, 128, 128, 3, 1, 24, 1, 0),
    ("probe enc1.1 x6 frames", 256, 256, 64, 64, 3, 1, 6, 1, 0),
    ("probe enc1.1 x12 frames", 256, 256, 64, 64, 3, 1, 12, 1, 0),
    ("probe enc3.1 x6 frames", 64, 64, 256, 256, 3, 1, 6, 1, 0),
    ("probe enc3.1 x12 frames", 64, 64, 256, 256, 3, 1, 12, 1, 0),
    ("mlp fwd 128->128 x100k relu", 782, 128, 128, 128, 1, 1, 1, 4, 0),
    ("gru zr 288->384 x100k", 782, 128, 288, 384, 1, 1, 1, 2, 1),
    ("gru q 288->192 x100k", 782, 128, 288, 192, 1, 1, 1, 3, 1),
]
WARM = int(os.environ.get("WARM", "3"))
REPS = int(os.environ.get("REPS", "20"))
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sel = sys.argv[2] if len(sys.argv) > 2 else ""
L = _lib.lib()
if os.environ.get("PERSIST") is not None:
    L.himo_conv_set_persistent(int(os.environ["PERSIST"]))
if os.environ.get("CTA2") is not None:
    L.himo_conv_set_2cta(int(os.environ["CTA2"]))
if os.environ.get("HALO") is not None:
    L.himo_conv_set_halo(int(os.environ["HALO"]))
if os.environ.get("MAXSMS") is not None:
    L.himo_conv_set_max_sms(int(os.environ["MAXSMS"]))
if os.environ.get("ATMEM") is not None:
    L.himo_conv_set_a_tmem(int(os.environ["ATMEM"]))
if os.environ.get("PAIRMIN") is not None:
    L.himo_conv_set_pair_min_mmas(int(os.environ["PAIRMIN"]))
if os.environ.get("WIDE") is not None:
    L.himo_conv_set_wide_tiles(int(os.environ["WIDE"]))
if os.environ.get("WRES") is not None:
    L.himo_conv_set_weights_resident(int(os.environ["WRES"]))
if os.environ.get("TSTORE") is not None:
    L.himo_conv_set_tma_store(int(os.environ["TSTORE"]))
if os.environ.get("ROWS2") is not None:
    L.himo_conv_set_rows2(int(os.environ["ROWS2"]))
if os.environ.get("PDL") is not None:
    L.himo_conv_set_pdl(int(os.environ["PDL"]))
if os.environ.get("FLUSH") is not None:
    L.himo_conv_set_flush_iters(int(os.environ["FLUSH"]))
for name, H, W, cin, cout, k, s, g, act, f32 in LAYERS:
    if sel and sel not in name:
        continue
    x = torch.randn(H, W, g * cin, device="cuda")
    xp = conv.split_planes(x, planes)
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    ws = conv.weight_prescale(w, planes)
    wp = conv.pack_conv_weight(w, planes, ws).cuda()
    Ho, Wo = H // s, W // s
    out = (torch.zeros((Ho, Wo, g * cout), device="cuda") if f32 else
           torch.zeros((planes, Ho, Wo, g * cout), dtype=torch.bfloat16, device="cuda"))
    bias = torch.randn(cout, device="cuda") * 0.1 if act else None
    def run():
        conv.conv2d_nhwc(xp, wp, bias, out, ksize=k, act=act, stride=s, cin=cin, n_groups=g, cin_group_stride=cin if g > 1 else 0,
                         cout_group_stride=cout if g > 1 else 0, acc_scale=1.0 / ws)
    for _ in range(WARM):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = REPS
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = 2.0 * g * Ho * Wo * cout * cin * k * k
    print(f"{name:28s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s alg  x{3 if planes == 2 else 1} = {fl * (3 if planes == 2 else 1) / us / 1e6:7.1f} tensor")
