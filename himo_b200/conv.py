"""Host-side helpers for the tcgen05 convolution (csrc/conv.cu): split-bf16 plane packing of
activations / weights and the ctypes call into `himo_conv2d_nhwc` / `himo_upsample2x_nhwc`."""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_longlong, c_void_p

import torch

from . import _lib


class ConvDesc(ctypes.Structure):
    """Mirror of `himo_conv_desc` (include/himo_b200.h)."""
    _fields_ = [
        ("in_", c_void_p), ("in_planes", c_int), ("in_plane_stride", c_longlong),
        ("H_in", c_int), ("W_in", c_int), ("Cin_total", c_int), ("cin_off", c_int), ("Cin", c_int),
        ("wgt", c_void_p), ("bias", c_void_p),
        ("Cout", c_int), ("ksize", c_int), ("stride", c_int),
        ("out", c_void_p), ("out_planes", c_int), ("out_plane_stride", c_longlong),
        ("Cout_total", c_int), ("cout_off", c_int), ("out_fp32", c_int), ("act", c_int),
        ("n_groups", c_int), ("cin_group_stride", c_int), ("cout_group_stride", c_int),
        ("acc_scale", ctypes.c_float),
        ("out_t", c_void_p), ("out_t_plane_stride", c_longlong), ("ld_t", c_int),
        ("mask_src", c_void_p), ("mask_plane_stride", c_longlong), ("mask_planes", c_int),
        ("b_group_k_stride", c_int), ("out_group_pix_stride", c_longlong), ("b_k_total", c_longlong),
        ("stop_flag", c_void_p),
        ("aux_h", c_void_p), ("aux_z", c_void_p), ("aux_ld", c_int),
        ("out2", c_void_p), ("out2_plane_stride", c_longlong), ("out2_ld", c_int),
        ("border_bias", c_void_p),
    ]


_lib.register("himo_conv2d_nhwc", c_int, [ctypes.POINTER(ConvDesc), c_void_p])
_lib.register("himo_upsample2x_nhwc", c_int,
              [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p, c_int, c_longlong, c_int, c_int,
               c_void_p])


def split_planes(x: torch.Tensor, planes: int) -> torch.Tensor:
    """fp32 [...] -> 16-bit [planes, ...] stored in a bf16-typed tensor.
    planes == 1: bf16(x).  planes == 2: split fp16 -- plane0 = fp16(x), plane1 = fp16(x - plane0),
    both bit-cast into the bf16 storage (11 + 11 mantissa bits)."""
    if planes == 1:
        return x.to(torch.bfloat16).unsqueeze(0).contiguous()
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return torch.stack([hi, lo], 0).contiguous().view(torch.bfloat16)


def merge_planes(p: torch.Tensor) -> torch.Tensor:
    if p.shape[0] == 1:
        return p[0].float()
    h = p.contiguous().view(torch.float16)
    return h[0].float() + h[1].float()


def weight_prescale(w: torch.Tensor, planes: int) -> float:
    """Power of two that brings max|w| to [256, 512) in split mode (keeps the low fp16 plane of the
    weights in the normal range); 1 for the single-plane bf16 mode."""
    if planes == 1:
        return 1.0
    m = float(w.abs().max())
    if m == 0.0:
        return 1.0
    import math
    return 2.0 ** (8 - math.floor(math.log2(m)))


def pack_conv_weight(w: torch.Tensor, planes: int, prescale: float = 1.0) -> torch.Tensor:
    """[Cout, Cin, kh, kw] fp32 -> [planes, Cout, kh*kw*Cin] 16-bit, K index = (ky*kw + kx)*Cin + ci."""
    cout, cin, kh, kw = w.shape
    k = (w * prescale).permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
    return split_planes(k, planes)


def conv2d_nhwc(x_planes: torch.Tensor, w_planes: torch.Tensor, bias, out: torch.Tensor, *, ksize: int,
                stride: int = 1, act: int = 0, cin_off: int = 0, cin: int | None = None, cout_off: int = 0,
                n_groups: int = 1, cin_group_stride: int = 0, cout_group_stride: int = 0,
                acc_scale: float = 1.0, border_bias: torch.Tensor | None = None) -> None:
    """x_planes [P,H,W,Cin_total] bf16; w_planes [P,Cout,K] bf16; out [Po,Ho,Wo,Cout_total] bf16 or
    [Ho,Wo,Cout_total] fp32 (written in place)."""
    P, H, W, Ct = x_planes.shape
    cout = w_planes.shape[1]
    cin = cin if cin is not None else w_planes.shape[2] // (ksize * ksize)
    d = ConvDesc()
    d.in_ = x_planes.data_ptr(); d.in_planes = P; d.in_plane_stride = H * W * Ct
    d.H_in, d.W_in, d.Cin_total, d.cin_off, d.Cin = H, W, Ct, cin_off, cin
    d.wgt = w_planes.data_ptr(); d.bias = bias.data_ptr() if bias is not None else None
    d.Cout, d.ksize, d.stride = cout, ksize, stride
    d.out = out.data_ptr()
    if out.dtype == torch.float32:
        d.out_fp32 = 1; d.out_planes = 1; d.out_plane_stride = 0
        d.Cout_total = out.shape[-1]
    else:
        d.out_fp32 = 0; d.out_planes = out.shape[0]
        d.out_plane_stride = out.shape[1] * out.shape[2] * out.shape[3]
        d.Cout_total = out.shape[3]
    d.cout_off = cout_off; d.act = act; d.acc_scale = acc_scale
    d.n_groups = n_groups; d.cin_group_stride = cin_group_stride; d.cout_group_stride = cout_group_stride
    d.border_bias = border_bias.data_ptr() if border_bias is not None else None     # [3,3,Cout] f32 (composed 1x1 -> 3x3)
    dev = x_planes.device
    with torch.cuda.device(dev):
        st = _lib.lib().himo_conv2d_nhwc(ctypes.byref(d), _lib.stream_ptr(dev))
    _lib.check(st, "conv2d_nhwc")


def upsample2x_nhwc(x_planes: torch.Tensor, out: torch.Tensor, cout_off: int = 0) -> None:
    P, h, w, c = x_planes.shape
    Po, H2, W2, Ct = out.shape
    assert H2 == 2 * h and W2 == 2 * w
    dev = x_planes.device
    with torch.cuda.device(dev):
        st = _lib.lib().himo_upsample2x_nhwc(_lib.ptr(x_planes), P, h * w * c, h, w, c, _lib.ptr(out), Po,
                                             H2 * W2 * Ct, Ct, cout_off, _lib.stream_ptr(dev))
    _lib.check(st, "upsample2x_nhwc")
