/*
 * himo_b200.h -- C ABI of libhimo_b200.so, the B200 (sm_100a) engine for the HiMo / OpenSceneFlow
 * per-frame-pair hot path.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - every entry point is asynchronous on `stream` and re-entrant per (device, stream);
 *     no global mutable state, no hidden allocations: scratch memory is a caller-supplied
 *     `workspace` whose size comes from the matching *_workspace_bytes() query.
 *   - return value: 0 ok, <0 argument / workspace / unsupported error, >0 a cudaError_t.
 *     Nothing throws across this boundary; the Python mirror raises RuntimeError, as the
 *     reference's TORCH_CHECK / AT_CUDA_CHECK do.
 *   - "replaces" cites the reference interface each function stands in for; paths are relative
 *     to the reference root, OSF/ = OpenSceneFlow/.
 */
#ifndef HIMO_B200_H
#define HIMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIMO_B200_ABI_VERSION 1
int himo_abi_version(void);
/* Number of kernels this library has launched in this process (diagnostic; bench.py "gpu_launches"). */
unsigned long long himo_launch_count(void);
/* Human-readable text for a status code returned by any entry point (static storage). */
const char* himo_status_string(int status);

/* ------------------------------------------------------------------------------------------
 * H1  dynamic voxelization
 * replaces: mmcv.dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3)
 *           OSF/assets/cuda/mmcv/pybind.cpp:47-49, voxelization.cpp:62-74,
 *           voxelization_cuda.cu:246-286 (kernel voxelization_cuda_kernel.cuh:13-50)
 * points      [num_points, num_features] f32 row-major (features >= 3; xyz first)
 * voxel_size  HOST float[3]; coors_range HOST float[6] (x_min,y_min,z_min,x_max,y_max,z_max)
 * coors       [num_points, 3] int32, written as (z,y,x); out-of-range rows become (-1,0,0),
 *             (-1,-1,0) or (-1,-1,-1) = what the reference leaves in its pre-zeroed buffer.
 */
int himo_dynamic_voxelize_forward(const float* points, int num_points, int num_features,
                                  const float* voxel_size, const float* coors_range,
                                  int32_t* coors, void* stream);

/* ------------------------------------------------------------------------------------------
 * H1  point-to-voxel scatter (sum / mean / max)
 * replaces: mmcv.dynamic_point_to_voxel_forward(feats, coors, reduce_type)
 *           OSF/assets/cuda/mmcv/pybind.cpp:33-35, scatter_points.cpp:36-41,
 *           scatter_points_cuda.cu:9-66 (kernel scatter_points_cuda_kernel.cuh:91-112)
 * coors       [num_points,3] int32 or int64 (coors_is_int64); rows with a negative entry are
 *             invalid (point2voxel = -1)
 * dims        HOST int32[3]: exclusive upper bound of every coors column (the voxel grid
 *             (gz,gy,gx); the Python mirror derives it from coors.max()).  The reference learns
 *             the same thing from its radix sort; we need it to size the occupancy bitmap.
 * reduce_type 0 sum, 1 mean, 2 max
 * outputs are caller-allocated for the worst case M = num_points voxels:
 *   voxel_feats [num_points,num_feats] f32, voxel_coors [num_points,3] (dtype of coors),
 *   point2voxel [num_points] i32, voxel_count [num_points] i32, num_voxels DEVICE int32[1] (= M).
 *   Rows >= M are unspecified.  Voxel order = ascending (c0,c1,c2), the reference's
 *   at::unique_dim order.
 */
size_t himo_dynamic_point_to_voxel_workspace_bytes(int num_points, int num_feats,
                                                   const int32_t* dims);
int himo_dynamic_point_to_voxel_forward(const float* feats, const void* coors, int coors_is_int64,
                                        int num_points, int num_feats, int reduce_type,
                                        const int32_t* dims, float* voxel_feats, void* voxel_coors,
                                        int32_t* point2voxel, int32_t* voxel_count,
                                        int32_t* num_voxels, void* workspace,
                                        size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * H2  exact nearest neighbour both ways (Chamfer correspondence)
 * replaces: chamfer3D.forward(pc0, pc1, dist0, dist1, idx0, idx1) -> 1
 *           OSF/assets/cuda/chamfer3D/chamfer3D_cuda.cpp:18-35, chamfer3D.cu:33-105
 * pc0 [n0,3], pc1 [n1,3] f32 contiguous.  dist* = squared L2 to the nearest point of the other
 * cloud (1e20 when that cloud is empty), idx* = its index (-1 when empty); ties resolve to the
 * lowest index, as the reference's ascending strict-< scan does.  The squared distance is rounded
 * exactly as the compiled reference kernel rounds it: fma(dz,dz, fma(dx,dx, dy*dy)) with d = pc1 - pc0
 * component differences (sequence read off the reference's sm_100a SASS).  cell_size <= 0 selects the
 * default search cell (0.5 m).
 */
size_t himo_chamfer_workspace_bytes(int n0, int n1);
int himo_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                         float* dist1, int32_t* idx0, int32_t* idx1, float cell_size,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Radius-limited form of the same search (the "grid-radius KNN" of the north star): the exact nearest
 * neighbour when it lies within `radius` metres (dist^2 <= radius^2), otherwise (1e20, -1).  Serves the
 * reference's truncated consumers without searching past the truncation: nnChamferDis.truncated_dis
 * zeroes d^2 >= 2 (OSF/assets/cuda/chamfer3D/__init__.py:77-82), SeFlow's TRUNCATED_DIST = 4
 * (OSF/src/lossfuncs/selfsupervise.py:23), the nnd auto-label keeps d < 4.4 m (OSF/process.py:124). */
int himo_chamfer_forward_radius(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                                float* dist1, int32_t* idx0, int32_t* idx1, float radius,
                                void* workspace, size_t workspace_bytes, void* stream);
/* replaces: chamfer3D.backward(pc0, pc1, idx0, idx1, grad_dist0, grad_dist1, grad_pc0, grad_pc1)
 *           OSF/assets/cuda/chamfer3D/chamfer3D.cu:107-154.  grad_pc0/grad_pc1 are accumulated
 *           into (the caller pre-zeroes them, as chamfer3D/__init__.py:44-45 does). */
int himo_chamfer_backward(const float* pc0, int n0, const float* pc1, int n1, const int32_t* idx0,
                          const int32_t* idx1, const float* grad_dist0, const float* grad_dist1,
                          float* grad_pc0, float* grad_pc1, void* stream);

/* A/B knob: 1 selects the warp-cooperative search kernel (one warp per query: coalesced leaf scans, redux.sync / shuffle
 * min-reductions) instead of one thread per query (default 0: identical results, but measured 1.2-3.6x slower on a B200
 * because the search is latency-bound and a warp per query has 32x fewer dependent-load chains in flight). */
int himo_chamfer_set_warp_search(int enable);

/* ------------------------------------------------------------------------------------------
 * (f)3  batched per-instance bidirectional nearest-neighbour distances (instance-level Chamfer of the HiMo metric)
 * replaces: the per-instance scipy cKDTree queries of `cal_chamfer` (eval.py:50-62) inside InstanceMetrics.step_eval
 *           (eval.py:88-96) and of tools/test/score.py:262-275, for ALL instances of a frame in one launch.
 * a, b        [na,3], [nb,3] float64 DEVICE: the points of every instance, instance by instance
 * a_off,b_off [n_seg+1] int32 DEVICE: CSR offsets of the instances in a / b
 * dist_a[i]   = min_j |a_i - b_j| over the points b_j of the SAME instance (dist_b likewise); +inf if it has none
 * max_seg_points = the largest instance size in a or b (sizes the grid).  Exact (brute force), float64 like the reference.
 */
int himo_segmented_nn(const double* a, const int32_t* a_off, const double* b, const int32_t* b_off, int n_seg,
                      int max_seg_points, double* dist_a, double* dist_b, void* stream);

/* ------------------------------------------------------------------------------------------
 * H4  dense convolution of the SeFlow++ backbone (tcgen05 implicit GEMM) and 2x bilinear upsample
 * replaces: the ATen/cuDNN work behind nn.Conv2d (+ BatchNorm2d + GELU) in
 *           ConvWithNorms.forward (OSF/src/models/basic/__init__.py:76-94) and
 *           UpsampleSkip.forward / decoder_step4 (OSF/src/models/basic/unet.py:18-35,130),
 *           i.e. every layer of UNetThreeFrame.forward (unet.py:131-166).
 * Activations are NHWC 16-bit "planes".  in_planes == 2: split fp16, plane 0 = fp16(x),
 * plane 1 = fp16(x - plane0) (fp32-class products from three tensor-core MMAs per k-step);
 * in_planes == 1: plain bf16.  Weights follow the same format; in split mode the host pre-scales
 * them by a power of two and passes its inverse as acc_scale (applied to the accumulator before
 * the bias), so their low plane stays in fp16's normal range.
 *   in      [in_planes][H_in][W_in][Cin_total] bf16; the conv reads channels
 *           [cin_off + g*cin_group_stride, +Cin) for group g (groups = frames sharing weights)
 *   wgt     [in_planes][Cout][ksize*ksize*Cin] bf16, K index = (ky*ksize + kx)*Cin + ci
 *   bias    [Cout] f32 or NULL (BatchNorm folded in by the host)
 *   out     NHWC, channels [cout_off + g*cout_group_stride, +Cout) of Cout_total; bf16 planes
 *           (out_planes 1|2) or fp32 (out_fp32 = 1)
 *   ksize 1|3 (padding ksize/2), stride 1|2, act 0 none | 1 exact-erf GELU | 2 sigmoid | 3 tanh | 4 ReLU | 5,6 fused ConvGRU
 */
typedef struct himo_conv_desc {
  const void* in;
  int in_planes;
  long long in_plane_stride; /* elements between input planes */
  int H_in, W_in, Cin_total, cin_off, Cin;
  const void* wgt;
  const float* bias;
  int Cout, ksize, stride;
  void* out;
  int out_planes;
  long long out_plane_stride;
  int Cout_total, cout_off;
  int out_fp32, act;
  int n_groups, cin_group_stride, cout_group_stride;
  float acc_scale;            /* accumulator multiplier before bias/activation; 0 means 1 */
  /* --- GEMM extensions (all optional, zero = off); used by the FastNSF MLP ------------------ */
  void* out_t;                /* transposed copy of the output: [planes][Cout_total][ld_t] */
  long long out_t_plane_stride;
  int ld_t;
  const void* mask_src;       /* ReLU-backward mask: output zeroed where this tensor (layout of out) is 0 */
  long long mask_plane_stride;
  int mask_planes;
  int b_group_k_stride;       /* split-K: K offset of the weight operand per group */
  long long out_group_pix_stride; /* split-K: output row offset per group */
  long long b_k_total;        /* row length of the weight operand when it is not ksize^2*Cin */
  const int32_t* stop_flag;   /* device flag; non-zero makes the launch a no-op (early stopping) */
  /* fused ConvGRU epilogues (OSF/src/models/basic/decoder.py:185-192): act 5 = [z|r] GEMM -> z to aux_z,
   * split(r*h) to out2; act 6 = q GEMM -> h = (1-z)h + z*tanh(.) to aux_h and split(h) to out2 */
  float* aux_h; float* aux_z; int aux_ld;
  void* out2; long long out2_plane_stride; int out2_ld;
  /* composed (1x1 conv -> zero padding -> 3x3 conv) layers: [3][3][Cout] f32 bias per border class (row class, column
   * class: 0 first, 1 interior, 2 last); replaces `bias` for the pixels on the image border.  NULL = off. */
  const float* border_bias;
} himo_conv_desc;
int himo_conv2d_nhwc(const himo_conv_desc* desc, void* stream);
/* Split-mode accuracy/speed knob: hi*hi MMAs (K = 16 each) accumulated in tensor memory before the partial
 * sum is drained into fp32 registers (default 48).  Process-wide. */
int himo_conv_set_flush_iters(int mmas);
/* A/B knob: 0 disables the haloed-row reuse of the activation tile across the kx taps (default on). */
int himo_conv_set_halo(int enable);
/* A/B knob: 1 stages the activation slices in tensor memory (tcgen05.cp.128x256b + TS-form MMAs: a_hi is read from
 * shared memory once for its two products); default 0 = both operands from shared memory (SS form).  Bit-identical;
 * measured neutral to slower on B200 (profiles/r01_conv_a_tmem_ab.txt). */
int himo_conv_set_a_tmem(int enable);
/* Experiment knob: SMs a persistent convolution launch may occupy (default 148). */
int himo_conv_set_max_sms(int n);
/* Tuning knob: CTA pairs (cta_group::2) are used for tiles with at least this many hi*hi MMAs (default 48). */
int himo_conv_set_pair_min_mmas(int n);
/* A/B knob: 0 disables the 256-channel-wide tiles (one N tile per CTA pair, accumulators fill TMEM) used for the
 * 256-channel encoder layers (default on). */
int himo_conv_set_wide_tiles(int enable);
/* A/B knob: 0 disables the weights-resident variants (whole weight tensor in shared memory) of the 64-channel
 * encoder layers (default on). */
int himo_conv_set_weights_resident(int enable);
/* A/B knob: 1 enables the TMA-store epilogue of the short-K 128-wide tiles (TMEM -> activation -> swizzled shared-memory
 * tile -> one cp.async.bulk.tensor store per plane).  Default 0: measured slower than per-thread stores on a B200
 * (profiles/r02_conv_tma_store_ab.txt). */
int himo_conv_set_tma_store(int enable);
/* A/B knob: 0 disables the two-output-rows tiles (k_conv_rows2: a CTA pair computes 2 rows x 256 px x 96 channels so that
 * activation rows and weight taps are shared; default on) of the 96-channel 3x3 decoder-half layers. */
int himo_conv_set_rows2(int enable);
/* A/B knob: 0 launches the backbone kernels (convolutions, upsample) without programmatic dependent launch, i.e. with
 * full stream serialisation between them (default on: the next kernel's prologue overlaps the previous kernel's tail). */
int himo_conv_set_pdl(int enable);
/* A/B knob: 0 launches one CTA per output tile instead of the persistent tile loop (default on). */
int himo_conv_set_persistent(int enable);
/* A/B knob: 0 disables the CTA-pair path (tcgen05.mma.cta_group::2 over a cluster of 2; default on). */
int himo_conv_set_2cta(int enable);
/* replaces: F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) of
 *           BilinearDecoder.forward (OSF/src/models/basic/unet.py:7-16); in [h][w][c] planes ->
 *           channels [cout_off, +c) of out [2h][2w][Cout_total] planes. */
int himo_upsample2x_nhwc(const void* in, int in_planes, long long in_plane_stride, int h, int w, int c,
                         void* out, int out_planes, long long out_plane_stride, int Cout_total,
                         int cout_off, void* stream);

/* ------------------------------------------------------------------------------------------
 * H1+H4  fused point embedder for the frames of one tuple (SeFlow++: t-1, t, t+1)
 * replaces: DynamicEmbedder.forward (OSF/src/models/basic/encoder.py:618-631) =
 *           DynamicVoxelizer (:567-600) + DynamicPillarFeatureNet (:430-475, two DynamicScatter
 *           calls) + PointPillarsScatter (:126-147), preceded by the rigid warp of
 *           wrap_batch_pcs (OSF/src/models/basic/__init__.py:50,57), for every frame.
 * points[f]    [num_points[f],3] f32 (ground-free; NaN rows = padding, dropped)
 * transform[f] row-major 3x4 (R|t) applied as p @ R^T + t when has_transform[f]
 * pfn_weight   [32,9] f32 and pfn_bias [32] f32: Linear(9,32,bias=False) with BatchNorm1d folded
 * canvas       [canvas_planes][gy][gx][n_frames*32] bf16 (split-bf16 planes); frame f owns
 *              channels [32f, 32f+32).  Cleared here unless skip_canvas_clear.
 * All per-point / per-voxel results stay in `workspace` (see himo_embed_views).
 */
#define HIMO_MAX_FRAMES 4
typedef struct himo_embed_desc {
  int n_frames;
  int n_max;                       /* row capacity of the per-frame workspace arrays */
  const float* points[HIMO_MAX_FRAMES];
  int num_points[HIMO_MAX_FRAMES];
  int has_transform[HIMO_MAX_FRAMES];
  float transform[HIMO_MAX_FRAMES][12];
  float voxel_size[3];
  float coors_range[6];
  double voxel_size_f64[3];        /* the Python-side doubles behind x_offset = vx/2 + x_min */
  double coors_range_f64[6];
  const float* pfn_weight;
  const float* pfn_bias;
  void* canvas;
  int canvas_planes;
  int skip_canvas_clear;
  void* workspace;
  size_t workspace_bytes;
  /* optional: the canvas is the channel slice [canvas_ch_off, +n_frames*32) of a wider NHWC buffer with canvas_ld
   * channels per pixel (0 = a dense canvas of n_frames*32 channels); `canvas` then points at the buffer's channel 0 */
  int canvas_ld;
  int canvas_ch_off;
} himo_embed_desc;
typedef struct himo_embed_view {     /* device pointers into the embed workspace, rows = frames */
  float* pt4;            /* [F][n_max][4] warped xyz + int32 cell key (y*gx+x, -1 = dropped) */
  unsigned* bitmap;      /* [F][n_words] occupancy bits */
  int* word_prefix;      /* [F][n_words] exclusive popcount prefix (voxel rank base) */
  int* num_voxels;       /* [F] */
  int* rank;             /* [F][n_max] point -> voxel (sorted order), -1 = dropped: point2voxel */
  int* voxel_count;      /* [F][n_max+1] */
  int* seg_start;        /* [F][n_max+1] */
  int* sorted_idx;       /* [F][n_max] */
  float* voxel_feats;    /* [F][n_max][32] */
  float* voxel_mean;     /* [F][n_max][3] */
  int* voxel_key;        /* [F][n_max] */
  int n_words;
} himo_embed_view;
size_t himo_embed_workspace_bytes(int n_frames, int n_max, const float* voxel_size,
                                  const float* coors_range);
int himo_embed_frames(const himo_embed_desc* desc, void* stream);
int himo_embed_views(int n_frames, int n_max, const float* voxel_size, const float* coors_range,
                     void* workspace, himo_embed_view* out);

/* ------------------------------------------------------------------------------------------
 * H4  SeFlow++ network forward for one frame triple
 * replaces: DeFlowPP.forward (OSF/src/models/deflow.py:115-158) behind the model-level boundary
 *           `model(batch) -> {"flow", "pose_flow", "pc0_valid_point_idxes", ...}` that
 *           ModelWrapper.test_step drives (OSF/src/trainer.py:290-343).
 * Weights are packed by the host (himo_b200/deflowpp.py): BatchNorm folded, conv weights as
 * [planes][Cout][taps*Cin] bf16 split planes (see himo_conv_desc).
 */
typedef struct himo_deflowpp_weights {
  int planes;                               /* 1 = bf16, 2 = split-bf16 (fp32-class) */
  const float* pfn_w; const float* pfn_b;   /* [32,9], [32] */
  const void* enc_w[16]; const float* enc_b[16];        /* encoder_step_1..3 in order */
  const void* dec_w[3][4]; const float* dec_b[3][4];    /* decoder_step1..3: u1, u3, u4, u5 */
  const void* dec4_w; const float* dec4_b;
  const float* off_w; const float* off_b;               /* head.offset_encoder [96,3],[96] */
  const void* gru_zr_w; const float* gru_zr_b;          /* [planes][384][288] = [convz; convr], [384] */
  const void* gru_q_w; const float* gru_q_b;            /* [planes][192][288], [192] */
  const void* dec0_w; const float* dec0_b;              /* head.decoder.0 padded to 64 rows */
  const float* dec2_w; const float* dec2_b;             /* head.decoder.2 [3,48],[3] */
  /* accumulator scales (inverse of the power-of-two weight pre-scale) per GEMM, 0 = 1 */
  float enc_s[16]; float dec_s[3][4]; float dec4_s; float gru_zr_s; float gru_q_s; float dec0_s;
  /* UpsampleSkip.forward (unet.py:31-35) applies u3 (1x1, on the skip) and u4 (3x3, on the concatenation) with nothing
   * in between, so the host may compose them: dec_w[b][2] then holds u4 with its skip half multiplied by u3
   * (W4[:, latent:] @ W3 per tap), dec_b[b][2] the interior bias b4 + sum_taps W4_tap b3, and dec_bb[b] the
   * [3][3][Cout] border-class biases (taps that fall into the zero padding do not see b3).  The u3 launch and its
   * output tensor disappear; every skip producer writes straight into the concatenation buffer.  NULL = not composed. */
  const float* dec_bb[3];
} himo_deflowpp_weights;
typedef struct himo_deflowpp_io {
  const float* pch1; int n_h1;              /* t-1 cloud, ground-free, sensor frame */
  const float* pc0; int n0;
  const float* pc1; int n1;
  float T_h1[12]; float T_0[12];            /* (R|t) of inv(pose1)@poseh1 and inv(pose1)@pose0, fp32 */
  int n_max;                                /* workspace row capacity (>= every n) */
  int num_iters;                            /* GRU iterations (2 for SeFlow++) */
  float* flow_all;                          /* [n0,3] network flow per pc0 point, 0 where dropped */
  int64_t* valid_idx;                       /* optional [n0]: pc0_valid_point_idxes (first n_valid) */
  float* flow_valid;                        /* optional [n0,3]: flow in the reference's compact form */
  int32_t* n_valid;                         /* optional DEVICE int32[1] */
  void* workspace; size_t workspace_bytes;
  /* optional cudaEvent_t handles recorded on `stream` at the stage boundaries: [0] start, [1] after the
   * embedder, [2] after the backbone (UNetThreeFrame), [3] after the decoder.  NULL entries are skipped. */
  void* stage_events[4];
} himo_deflowpp_io;
/* A/B knob: 0 runs the ConvGRU decoder as separate GEMM + element-wise launches instead of the fused
 * persistent kernel (default 1, split-plane mode only). */
int himo_deflowpp_set_fused_decoder(int enable);
typedef struct himo_deflowpp_view {
  void* canvas; void* Fstar; void* Lstar; void* Rstar; void* S; void* T; void* U; float* V;
  void* embed_ws; float* h32;
} himo_deflowpp_view;
size_t himo_deflowpp_workspace_bytes(int n_max, int planes);
int himo_deflowpp_forward(const himo_deflowpp_weights* w, const himo_deflowpp_io* io, void* stream);
int himo_deflowpp_views(int n_max, int planes, void* workspace, himo_deflowpp_view* out);
/* out[i] = (p_i @ R^T + t) - p_i (+ add_flow[i] if given): pose flow / final-flow assembly of
 * ModelWrapper.test_step (OSF/src/trainer.py:320-335).  T12_dev: DEVICE float[12]. */
int himo_rigid_flow(const float* points, int n, const float* T12_dev, const float* add_flow, float* out,
                    void* stream);
/* Total flow of ALL points of pc0 (ground included), the array ModelWrapper.test_step writes to the
 * .h5 (OSF/src/trainer.py:320-343): out[i] = pose_flow(points_all[i]) + (src[i] >= 0 ? flow[src[i]] : 0),
 * src = index of point i in the ground-free cloud (NULL = identity). */
int himo_final_flow(const float* points_all, int n_all, const float* T12_dev, const float* flow,
                    const int32_t* src_index, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * H3  FastNSF: distance volume + per-frame-pair MLP optimisation
 * replaces: FastNSF.optimize (OSF/src/models/fastnsf.py:105-169) = DT.__init__ (:30-57, which calls the
 *           third-party FastGeodis.generalised_geodesic3d), DT.torch_bilinear_distance (:59-80),
 *           Neural_Prior.forward + autograd (OSF/src/models/basic/nsfp_module.py:7-47), torch.optim.Adam
 *           and EarlyStopping.step (nsfp_module.py:51-97).
 * himo_nsf_volume_geometry: lo = floor(min*gf-1)/gf and dims = ceil((hi-lo)*gf)+2 over both clouds
 *   (fastnsf.py:120-126, :34-36); returns them to the HOST (one small sync per frame pair).
 *   workspace: >= 1 KiB of device scratch.
 * himo_nsf_dt_build: D[dims0][dims1][dims2] f32 = raster Euclidean transform (FastGeodis semantics,
 *   lamb = 0, v = 1e10, 1 iteration) of the occupancy of pc1 at round((p - lo)*gf).
 * himo_nsf_optimize: runs up to max_iters iterations with early stopping on the device and writes the
 *   best flow [n,3].  init_params / final_params: DEVICE float[116483] in the reference's state_dict
 *   order (W0[128,3], b0, W1[128,128], b1, ..., W7, b7, W8[3,128], b8).  Blocking call.
 */
#define HIMO_NSF_NUM_PARAMS 116483
typedef struct himo_nsf_desc {
  const float* pc0; int n;            /* ego-compensated, range-limited source cloud */
  int n_max;                          /* workspace capacity in points */
  const float* D; float lo[3]; int32_t dims[3]; float grid_factor;
  const float* init_params; float* final_params;
  float* exp_avg_out;                 /* optional DEVICE float[116483]: Adam first moment at exit (tests) */
  int planes;                         /* 2 = split fp16 (fp32-class), 1 = bf16 */
  int max_iters; float lr; float min_delta; int patience; int poll_iters;
  float* best_flow;                   /* [n,3] */
  int32_t* iterations_out; float* best_loss_out; float* last_loss_out;   /* HOST, optional */
  void* workspace; size_t workspace_bytes;
} himo_nsf_desc;
size_t himo_nsf_workspace_bytes(int n_max, int planes);
/* A/B knob: 0 runs the seven hidden layers of the prior as GEMM launches (7 forward + 7 backward per iteration) instead of
 * the two fused chain kernels (k_mlp_chain: a 256-point tile stays in shared memory across the layers; default 1). */
int himo_nsf_set_fused(int enable);
/* 1 (default): the FastNSF head (output layer, trilinear lookup, delta_8) runs one warp per point with coalesced rows
 * (k_nsf_head_warp); 0: one thread per point (k_nsf_head).  Same results up to the order of the fp32 sums. */
int himo_nsf_set_head_warp(int enable);
/* 1: himo_nsf_optimize waits for each chunk of iterations on a blocking-sync event (the calling thread sleeps); 0 (default):
 * it spins.  Set by FastNSFEngine.infer_stream, which optimises several pairs from worker threads: spinning pollers would
 * each take a host core for the whole run.  Costs ~0.05 ms per iteration in wake-up latency for a lone optimiser. */
int himo_nsf_set_blocking_poll(int enable);
/* 1 (default): the axis-0 / axis-1 raster passes of himo_nsf_dt_build run as one thread-block-cluster launch each
 * (k_nsf_dt_sweep: halo rows exchanged through distributed shared memory, one cluster barrier per plane); 0: the tiled
 * multi-launch passes.  Bit-identical results. */
int himo_nsf_set_dt_cluster(int enable);
/* Variant of the tiled raster pass on planes of >= 256 k cells (the axis-2 pass): 0 = 16x16 tiles advancing 16 planes per
 * launch, 1 = 32x32 tiles x 16 planes, 2 = 16x16 x 8 planes, 3 (default) = 16x16 x 4 planes.  Bit-identical results;
 * 1.90 / 3.87 / 1.41 / 1.28 ms per direction at 1040 x 1030 x 52. */
int himo_nsf_set_dt_big_tiles(int variant);
/* One raster pass of the distance transform in place on D[dims] (axis 0..2, dir +1 / -1); sweep = 1 runs it as the
 * cluster kernel (axis 0 / 1 only; HIMO_ERR_UNSUPPORTED when the plane does not fit), 0 as the tiled launches.  For tests. */
int himo_nsf_dt_pass(float* D, const int32_t* dims, float grid_factor, int axis, int dir, int sweep, void* stream);
/* Profiling aid: device buffer of 16 x 4 x 4 int64 (or NULL).  When set, k_nsf_dt_sweep adds up the clock64 cycles its steps
 * spend in (cp.async wait, halo mbarrier wait, compute, __syncthreads) for four probe threads of every CTA. */
int himo_nsf_set_dt_debug_buffer(long long* device_buffer);
int himo_nsf_volume_geometry(const float* pc0, int n0, const float* pc1, int n1, float grid_factor,
                             float* lo_host, int32_t* dims_host, void* workspace, void* stream);
int himo_nsf_dt_build(const float* pc1, int n1, const float* lo, const int32_t* dims, float grid_factor,
                      float* D, void* stream);
int himo_nsf_optimize(const himo_nsf_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------
 * H3 / (f)1  the same 3 -> 128 x 8 -> 3 ReLU prior with the loss OUTSIDE the library (NSFP)
 * replaces: Neural_Prior.forward, its autograd backward and torch.optim.Adam.step as NSFP.optimize uses them
 *           (OSF/src/models/nsfp.py:74-131, OSF/src/models/basic/nsfp_module.py:7-47), plus the loop's
 *           best-flow / EarlyStopping bookkeeping (nsfp.py:104-113, nsfp_module.py:60-82) on the device.
 * All state -- parameters (reference state_dict order, HIMO_NSF_NUM_PARAMS floats), Adam moments, activations,
 * control block -- lives in the caller's workspace of himo_nsf_workspace_bytes(n_max, planes) bytes; every call
 * re-derives its views from (workspace, n_max, planes).  `ctl_workspace` (NULL = own) makes a network follow the
 * control block (iteration count, stop flag) of another one: NSFP steps `net` and `net_inv` in lockstep.
 *   himo_mlp_init        parameters <- init_params (DEVICE), moments <- 0, control block reset
 *   himo_mlp_forward     out[n,3] = MLP(x[n,3]); keeps the activations for the backward pass
 *   himo_mlp_backward    d_out[n,3] = d loss / d out  ->  parameter gradients (in the workspace) and, if d_x != NULL,
 *                        d_x[n,3] = d loss / d x
 *   himo_mlp_adam_step   Adam(lr, betas 0.9 / 0.999, eps 1e-8, no weight decay) with those gradients; the step
 *                        number is the control block's iteration count
 *   himo_mlp_control     one loss evaluation: iteration count += 1; if loss <= best: best <- loss and
 *                        best_out[count] <- out[count]; EarlyStopping.step(loss) -> stop flag (kernels of a stopped
 *                        network are no-ops).  loss_dev: DEVICE float.
 *   himo_mlp_read_state  blocking: state_host[4] = {stop, iterations, best loss, last loss}; optional DEVICE copies
 *                        of the parameters and of Adam's first moment
 */
int himo_mlp_init(void* workspace, size_t workspace_bytes, int n_max, int planes, const float* init_params,
                  void* stream);
int himo_mlp_forward(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace,
                     const float* x, int n, float* out, void* stream);
int himo_mlp_backward(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace, int n,
                      const float* d_out, float* d_x, void* stream);
int himo_mlp_adam_step(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace, int n,
                       float lr, void* stream);
int himo_mlp_control(void* workspace, size_t workspace_bytes, int n_max, int planes, const float* loss_dev,
                     float min_delta, int patience, const float* out, float* best_out, long long count, void* stream);
int himo_mlp_read_state(void* workspace, size_t workspace_bytes, int n_max, int planes, float* state_host,
                        float* params_out, float* exp_avg_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIMO_B200_H */
