// himo_b200/csrc/deflowpp.cu -- H4: the whole SeFlow++ network forward as one C call.
//
// Replaces DeFlowPP.forward (OSF/src/models/deflow.py:115-158) for one frame triple: embedder x3 ->
// UNetThreeFrame (OSF/src/models/basic/unet.py:131-166) -> ConvGRUDecoder(96, 2 iterations)
// (OSF/src/models/basic/decoder.py:195-253).  ~60 kernel launches on one stream, no host sync, no
// allocation: every intermediate lives in the caller's workspace, activations as NHWC split-bf16
// planes with the three frames side by side in the channel dimension, so the reference's torch.cat
// calls (unet.py:152-159, deflow.py:141) cost nothing -- producers write straight into channel slices.
#include "common.cuh"
#include "dec.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

struct Act {
  __nv_bfloat16* p;
  int H, W, C;
  long long plane_stride() const { return (long long)H * W * C; }
};

struct NetBuffers {
  Act B, Fa, Fb, La, Lb, Ra, Rb, T1, CAT1, S, T2, CAT2, Tt, T3, CAT3, X3, U;
  float* V;               // [512*512][96] fp32
  void* embed_ws; size_t embed_ws_bytes;
  float* h32; __nv_bfloat16* hx; __nv_bfloat16* rhx; float* zr; float* q; float* y;
  int* pos; void* scan_scratch;
};

static size_t net_layout(int n_max, int planes, const float* vs, const float* cr, NetBuffers* nb, void* base) {
  Arena A(base, (size_t)-1);
  auto act = [&](int H, int C) {
    Act a; a.H = H; a.W = H; a.C = C;
    a.p = A.take<__nv_bfloat16>((size_t)planes * H * H * C);
    return a;
  };
  NetBuffers b;
  b.B = act(512, 96);
  b.Fa = act(256, 192); b.Fb = act(256, 192);
  b.La = act(128, 384); b.Lb = act(128, 384);
  b.Ra = act(64, 768); b.Rb = act(64, 768);
  b.T1 = act(64, 384); b.CAT1 = act(128, 768); b.S = act(128, 384);
  b.T2 = act(128, 192); b.CAT2 = act(256, 384); b.Tt = act(256, 192);
  b.T3 = act(256, 96); b.CAT3 = act(512, 192); b.X3 = act(512, 96); b.U = act(512, 96);
  b.V = A.take<float>((size_t)512 * 512 * 96);
  b.embed_ws_bytes = himo_embed_workspace_bytes(3, n_max, vs, cr);
  b.embed_ws = A.take<char>(b.embed_ws_bytes);
  const size_t n_pad = (size_t)ceil_div(n_max > 0 ? n_max : 1, 256) * 256;
  b.h32 = A.take<float>(n_pad * 192);
  b.hx = A.take<__nv_bfloat16>((size_t)planes * n_pad * 288);
  b.rhx = A.take<__nv_bfloat16>((size_t)planes * n_pad * 288);
  b.zr = A.take<float>(n_pad * 384);
  b.q = A.take<float>(n_pad * 192);
  b.y = A.take<float>(n_pad * 64);
  b.pos = A.take<int>(n_pad);
  b.scan_scratch = A.take<char>(ScanScratch::bytes((long long)n_pad + 1));
  if (nb) *nb = b;
  return A.off + 1024;
}

static int conv(const Act& in, int cin_off, int cin, int groups, const void* w, const float* bias, int cout,
                int ksize, int stride, int act, const Act& out, int cout_off, int planes, cudaStream_t stream,
                float acc_scale, float* out_f32 = nullptr, const float* border_bias = nullptr) {
  himo_conv_desc d = {};
  d.in = in.p; d.in_planes = planes; d.in_plane_stride = in.plane_stride();
  d.H_in = in.H; d.W_in = in.W; d.Cin_total = in.C; d.cin_off = cin_off; d.Cin = cin;
  d.wgt = w; d.bias = bias; d.Cout = cout; d.ksize = ksize; d.stride = stride;
  if (out_f32) { d.out = out_f32; d.out_fp32 = 1; d.out_planes = 1; d.out_plane_stride = 0; d.Cout_total = cout; }
  else { d.out = out.p; d.out_fp32 = 0; d.out_planes = planes; d.out_plane_stride = out.plane_stride(); d.Cout_total = out.C; }
  d.cout_off = cout_off; d.act = act; d.acc_scale = acc_scale;
  d.n_groups = groups; d.cin_group_stride = groups > 1 ? cin : 0; d.cout_group_stride = groups > 1 ? cout : 0;
  d.border_bias = border_bias;
  return himo_conv2d_nhwc(&d, stream);
}

// ConvGRU GEMM with the gate math fused into the epilogue (conv.cu act 5 / 6); z lives in zbuf [n][192].
static int gru_gemm(const Act& in, const void* w, const float* bias, int cout, int act, float acc_scale, float* h32,
                    float* zbuf, __nv_bfloat16* out2, long long ps, int planes, cudaStream_t stream) {
  himo_conv_desc d = {};
  d.in = in.p; d.in_planes = planes; d.in_plane_stride = in.plane_stride();
  d.H_in = in.H; d.W_in = in.W; d.Cin_total = in.C; d.Cin = in.C;
  d.wgt = w; d.bias = bias; d.Cout = cout; d.ksize = 1; d.stride = 1;
  d.out = nullptr; d.out_planes = planes; d.Cout_total = cout; d.act = act; d.acc_scale = acc_scale; d.n_groups = 1;
  d.aux_h = h32; d.aux_z = zbuf; d.aux_ld = 192; d.out2 = out2; d.out2_plane_stride = ps; d.out2_ld = 288;
  return himo_conv2d_nhwc(&d, stream);
}

#define HIMO_RET(expr) do { int _s = (expr); if (_s != HIMO_OK) return _s; } while (0)

static int g_dec_fused = 1;

}  // namespace himo

using namespace himo;

// A/B knob: 0 runs the ConvGRU decoder as separate GEMM + element-wise launches (the pre-fusion path).
extern "C" int himo_deflowpp_set_fused_decoder(int enable) { g_dec_fused = enable ? 1 : 0; return HIMO_OK; }

extern "C" size_t himo_deflowpp_workspace_bytes(int n_max, int planes) {
  if (n_max < 0 || (planes != 1 && planes != 2)) return 0;
  const float vs[3] = {0.2f, 0.2f, 6.f}, cr[6] = {-51.2f, -51.2f, -3.f, 51.2f, 51.2f, 3.f};
  return net_layout(n_max, planes, vs, cr, nullptr, nullptr);
}

extern "C" int himo_deflowpp_forward(const himo_deflowpp_weights* w, const himo_deflowpp_io* io, void* stream_) {
  if (!w || !io || !io->workspace) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int P = w->planes;
  if (P != 1 && P != 2) return HIMO_ERR_ARG;
  const float vs[3] = {0.2f, 0.2f, 6.f}, cr[6] = {-51.2f, -51.2f, -3.f, 51.2f, 51.2f, 3.f};
  const int n_in[3] = {io->n_h1, io->n0, io->n1};
  for (int f = 0; f < 3; ++f) if (n_in[f] < 0 || n_in[f] > io->n_max) return HIMO_ERR_ARG;
  NetBuffers nb;
  if (net_layout(io->n_max, P, vs, cr, &nb, io->workspace) > io->workspace_bytes) return HIMO_ERR_WORKSPACE;

  auto mark = [&](int k) {
    if (io->stage_events[k]) cudaEventRecord((cudaEvent_t)io->stage_events[k], stream);
  };
  mark(0);
  // ---- embedder x3 (deflow.py:129-131); frame order (pch1, pc0, pc1) = channel slices 0,1,2
  himo_embed_desc e;
  e.n_frames = 3; e.n_max = io->n_max;
  e.points[0] = io->pch1; e.points[1] = io->pc0; e.points[2] = io->pc1;
  for (int f = 0; f < 3; ++f) e.num_points[f] = n_in[f];
  e.has_transform[0] = 1; e.has_transform[1] = 1; e.has_transform[2] = 0;
  for (int k = 0; k < 12; ++k) { e.transform[0][k] = io->T_h1[k]; e.transform[1][k] = io->T_0[k]; e.transform[2][k] = 0.f; }
  for (int k = 0; k < 3; ++k) { e.voxel_size[k] = vs[k]; }
  for (int k = 0; k < 6; ++k) { e.coors_range[k] = cr[k]; }
  e.voxel_size_f64[0] = 0.2; e.voxel_size_f64[1] = 0.2; e.voxel_size_f64[2] = 6.0;
  const double crd[6] = {-51.2, -51.2, -3.0, 51.2, 51.2, 3.0};
  for (int k = 0; k < 6; ++k) e.coors_range_f64[k] = crd[k];
  e.pfn_weight = w->pfn_w; e.pfn_bias = w->pfn_b;
  // composed u3 -> u4 (himo_deflowpp_weights::dec_bb): the skip tensors are consumed by u4 directly, so every skip
  // producer writes into the second half of its block's concatenation buffer and the three u3 launches disappear
  const bool composed = w->dec_bb[0] && w->dec_bb[1] && w->dec_bb[2];
  e.canvas_planes = P; e.skip_canvas_clear = 0;
  if (composed) { e.canvas = nb.CAT3.p; e.canvas_ld = nb.CAT3.C; e.canvas_ch_off = 96; }
  else { e.canvas = nb.B.p; e.canvas_ld = 0; e.canvas_ch_off = 0; }
  e.workspace = nb.embed_ws; e.workspace_bytes = nb.embed_ws_bytes;
  HIMO_RET(himo_embed_frames(&e, stream));

  mark(1);
  // ---- shared encoder on the three pseudo-images (unet.py:139-150), frames = conv groups
  int li = 0;
  auto enc = [&](const Act& in, int cin_off, int cin, const Act& out, int cout_off, int cout, int stride) {
    int s = conv(in, cin_off, cin, 3, w->enc_w[li], w->enc_b[li], cout, 3, stride, 1, out, cout_off, P, stream, w->enc_s[li]);
    ++li;
    return s;
  };
  // skip tensors: Bstar [512^2,96] (canvas), Fstar [256^2,192], Lstar [128^2,384], Rstar = Rb [64^2,768]
  const Act& Bs = composed ? nb.CAT3 : nb.B;   const int b_off = composed ? 96 : 0;
  const Act& Fs = composed ? nb.CAT2 : nb.Fb;  const int f_off = composed ? 192 : 0;
  const Act& Ls = composed ? nb.CAT1 : nb.Lb;  const int l_off = composed ? 384 : 0;
  HIMO_RET(enc(Bs, b_off, 32, nb.Fa, 0, 64, 2));
  HIMO_RET(enc(nb.Fa, 0, 64, nb.Fb, 0, 64, 1)); HIMO_RET(enc(nb.Fb, 0, 64, nb.Fa, 0, 64, 1)); HIMO_RET(enc(nb.Fa, 0, 64, Fs, f_off, 64, 1));
  HIMO_RET(enc(Fs, f_off, 64, nb.La, 0, 128, 2));
  HIMO_RET(enc(nb.La, 0, 128, nb.Lb, 0, 128, 1)); HIMO_RET(enc(nb.Lb, 0, 128, nb.La, 0, 128, 1)); HIMO_RET(enc(nb.La, 0, 128, nb.Lb, 0, 128, 1));
  HIMO_RET(enc(nb.Lb, 0, 128, nb.La, 0, 128, 1)); HIMO_RET(enc(nb.La, 0, 128, Ls, l_off, 128, 1));
  HIMO_RET(enc(Ls, l_off, 128, nb.Ra, 0, 256, 2));
  HIMO_RET(enc(nb.Ra, 0, 256, nb.Rb, 0, 256, 1)); HIMO_RET(enc(nb.Rb, 0, 256, nb.Ra, 0, 256, 1)); HIMO_RET(enc(nb.Ra, 0, 256, nb.Rb, 0, 256, 1));
  HIMO_RET(enc(nb.Rb, 0, 256, nb.Ra, 0, 256, 1)); HIMO_RET(enc(nb.Ra, 0, 256, nb.Rb, 0, 256, 1));

  // ---- UpsampleSkip x3 (unet.py:31-35) + decoder_step4
  auto up_block = [&](int bi, const Act& a, const Act& skip, int skip_off, int latent, int out_c, const Act& T,
                      const Act& CAT, const Act& X, const Act& Y) {
    HIMO_RET(conv(a, 0, a.C, 1, w->dec_w[bi][0], w->dec_b[bi][0], latent, 1, 1, 0, T, 0, P, stream, w->dec_s[bi][0]));
    HIMO_RET(himo_upsample2x_nhwc(T.p, P, T.plane_stride(), T.H, T.W, latent, CAT.p, P, CAT.plane_stride(), CAT.C, 0, stream));
    if (!composed)
      HIMO_RET(conv(skip, skip_off, latent, 1, w->dec_w[bi][1], w->dec_b[bi][1], latent, 1, 1, 0, CAT, latent, P, stream, w->dec_s[bi][1]));
    HIMO_RET(conv(CAT, 0, CAT.C, 1, w->dec_w[bi][2], w->dec_b[bi][2], out_c, 3, 1, 0, X, 0, P, stream, w->dec_s[bi][2],
                  nullptr, composed ? w->dec_bb[bi] : nullptr));
    HIMO_RET(conv(X, 0, X.C, 1, w->dec_w[bi][3], w->dec_b[bi][3], out_c, 3, 1, 0, Y, 0, P, stream, w->dec_s[bi][3]));
    return HIMO_OK;
  };
  HIMO_RET(up_block(0, nb.Rb, Ls, l_off, 384, 384, nb.T1, nb.CAT1, nb.La, nb.S));
  HIMO_RET(up_block(1, nb.S, Fs, f_off, 192, 192, nb.T2, nb.CAT2, nb.Fa, nb.Tt));
  HIMO_RET(up_block(2, nb.Tt, Bs, b_off, 96, 96, nb.T3, nb.CAT3, nb.X3, nb.U));
  HIMO_RET(conv(nb.U, 0, 96, 1, w->dec4_w, w->dec4_b, 96, 3, 1, 0, nb.U, 0, P, stream, w->dec4_s, nb.V));

  mark(2);
  // ---- ConvGRU decoder on the pc0 points (decoder.py:210-237)
  const int n0 = io->n0;
  if (n0 > 0) {
    himo_embed_view ev;
    HIMO_RET(himo_embed_views(3, io->n_max, vs, cr, nb.embed_ws, &ev));
    const bool fused = g_dec_fused && P == 2;
    const int n_pad = fused ? ceil_div(n0, 256) * 256 : ceil_div(n0, 128) * 128;
    const long long ps = (long long)n_pad * 288;
    DecGatherArgs g;
    g.pt4 = (const float4*)ev.pt4 + (size_t)1 * io->n_max;     // frame 1 = pc0
    g.n = n0; g.n_pad = n_pad; g.n_frames = 3;
    g.bitmap = ev.bitmap; g.word_prefix = ev.word_prefix; g.voxel_feats = ev.voxel_feats;
    g.n_words = ev.n_words; g.n_max = io->n_max;
    g.after = nb.V; g.c_after = 96; g.w_off = w->off_w; g.b_off = w->off_b;
    g.vx = vs[0]; g.vy = vs[1]; g.vz = vs[2]; g.x_min = cr[0]; g.y_min = cr[1]; g.z_min = cr[2];
    g.hx = vs[0] / 2; g.hy = vs[1] / 2; g.hz = vs[2] / 2; g.gx = 512;
    g.h32 = fused ? nullptr : nb.h32; g.hx_planes = nb.hx; g.rhx_planes = fused ? nullptr : nb.rhx;
    g.planes = P; g.plane_stride = ps;
    HIMO_RET(dec_gather(g, stream));
    if (fused) {
      // the whole ConvGRU + head in one persistent kernel (csrc/decfused.cu)
      HIMO_RET(dec_fused(nb.hx, n0, n_pad, io->num_iters, w, g.pt4, io->flow_all, stream));
    } else {
    Act HX{nb.hx, n_pad / 128, 128, 288}, RHX{nb.rhx, n_pad / 128, 128, 288}, none{nullptr, 0, 0, 0};
    for (int it = 0; it < io->num_iters; ++it) {
      // The gate math could ride in the GEMM epilogue (conv.cu act 5/6, kept for reference) but measured
      // 2x slower: the per-row global reads of h and z sit on the tile's critical path.  Separate,
      // fully vectorised element-wise kernels are cheaper.
      HIMO_RET(conv(HX, 0, 288, 1, w->gru_zr_w, w->gru_zr_b, 384, 1, 1, 2, none, 0, P, stream, w->gru_zr_s, nb.zr));
      HIMO_RET(dec_rh(nb.zr, nb.h32, n_pad, nb.rhx, P, ps, stream));
      HIMO_RET(conv(RHX, 0, 288, 1, w->gru_q_w, w->gru_q_b, 192, 1, 1, 3, none, 0, P, stream, w->gru_q_s, nb.q));
      HIMO_RET(dec_update(nb.zr, nb.q, nb.h32, n_pad, nb.hx, P, ps, stream));
    }
    HIMO_RET(conv(HX, 0, 288, 1, w->dec0_w, w->dec0_b, 64, 1, 1, 1, none, 0, P, stream, w->dec0_s, nb.y));
    HIMO_RET(dec_out(nb.y, 64, g.pt4, n0, w->dec2_w, w->dec2_b, io->flow_all, stream));
    }
    if (io->valid_idx && io->flow_valid && io->n_valid)
      HIMO_RET(dec_compact(g.pt4, n0, nb.pos, io->n_valid, nb.scan_scratch, io->valid_idx, io->flow_all,
                           io->flow_valid, stream));
  } else if (io->n_valid) {
    HIMO_CUDA_RET(cudaMemsetAsync(io->n_valid, 0, sizeof(int), stream));
  }
  mark(3);
  return HIMO_OK;
}

// Device pointers of the named intermediates inside a deflowpp workspace (tests / profiling).
extern "C" int himo_deflowpp_views(int n_max, int planes, void* workspace, himo_deflowpp_view* out) {
  if (!workspace || !out) return HIMO_ERR_ARG;
  const float vs[3] = {0.2f, 0.2f, 6.f}, cr[6] = {-51.2f, -51.2f, -3.f, 51.2f, 51.2f, 3.f};
  NetBuffers nb;
  net_layout(n_max, planes, vs, cr, &nb, workspace);
  out->canvas = nb.B.p; out->Fstar = nb.Fb.p; out->Lstar = nb.Lb.p; out->Rstar = nb.Rb.p;
  out->S = nb.S.p; out->T = nb.Tt.p; out->U = nb.U.p; out->V = nb.V;
  out->embed_ws = nb.embed_ws; out->h32 = nb.h32;
  return HIMO_OK;
}

// pose_flow = (p @ R^T + t) - p, the ego-motion part of the total flow
// (OSF/src/trainer.py:320-323; wrap_batch_pcs basic/__init__.py:50-52).
namespace himo {
__global__ void __launch_bounds__(256)
k_rigid_flow(const float* __restrict__ pts, int n, const float* __restrict__ T12, float* __restrict__ out,
             const float* __restrict__ add, const int32_t* __restrict__ src) {
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = T12[k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    float fx = __fadd_rn(__fmaf_rn(z, T[2], __fmaf_rn(y, T[1], __fmul_rn(x, T[0]))), T[3]) - x;
    float fy = __fadd_rn(__fmaf_rn(z, T[6], __fmaf_rn(y, T[5], __fmul_rn(x, T[4]))), T[7]) - y;
    float fz = __fadd_rn(__fmaf_rn(z, T[10], __fmaf_rn(y, T[9], __fmul_rn(x, T[8]))), T[11]) - z;
    if (add) {
      const int j = src ? src[i] : i;
      if (j >= 0) { fx += add[3 * (size_t)j]; fy += add[3 * (size_t)j + 1]; fz += add[3 * (size_t)j + 2]; }
    }
    out[3 * (size_t)i] = fx; out[3 * (size_t)i + 1] = fy; out[3 * (size_t)i + 2] = fz;
  }
}
}  // namespace himo

extern "C" int himo_rigid_flow(const float* points, int n, const float* T12_dev, const float* add_flow,
                               float* out, void* stream_) {
  if (n < 0) return HIMO_ERR_ARG;
  if (n == 0) return HIMO_OK;
  if (!points || !T12_dev || !out) return HIMO_ERR_ARG;
  k_rigid_flow<<<min(ceil_div(n, 256), kNumSMs * 8), 256, 0, (cudaStream_t)stream_>>>(points, n, T12_dev, out,
                                                                                     add_flow, nullptr);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

extern "C" int himo_final_flow(const float* points_all, int n_all, const float* T12_dev, const float* flow,
                               const int32_t* src_index, float* out, void* stream_) {
  if (n_all < 0) return HIMO_ERR_ARG;
  if (n_all == 0) return HIMO_OK;
  if (!points_all || !T12_dev || !out) return HIMO_ERR_ARG;
  k_rigid_flow<<<min(ceil_div(n_all, 256), kNumSMs * 8), 256, 0, (cudaStream_t)stream_>>>(points_all, n_all, T12_dev,
                                                                                         out, flow, src_index);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
