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
 * lowest index, as the reference's ascending strict-< scan does.  cell_size <= 0 selects the
 * default search cell (0.5 m).
 */
size_t himo_chamfer_workspace_bytes(int n0, int n1);
int himo_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                         float* dist1, int32_t* idx0, int32_t* idx1, float cell_size,
                         void* workspace, size_t workspace_bytes, void* stream);
/* replaces: chamfer3D.backward(pc0, pc1, idx0, idx1, grad_dist0, grad_dist1, grad_pc0, grad_pc1)
 *           OSF/assets/cuda/chamfer3D/chamfer3D.cu:107-154.  grad_pc0/grad_pc1 are accumulated
 *           into (the caller pre-zeroes them, as chamfer3D/__init__.py:44-45 does). */
int himo_chamfer_backward(const float* pc0, int n0, const float* pc1, int n1, const int32_t* idx0,
                          const int32_t* idx1, const float* grad_dist0, const float* grad_dist1,
                          float* grad_pc0, float* grad_pc1, void* stream);

/* ------------------------------------------------------------------------------------------
 * H4  dense convolution of the SeFlow++ backbone (tcgen05 implicit GEMM) and 2x bilinear upsample
 * replaces: the ATen/cuDNN work behind nn.Conv2d (+ BatchNorm2d + GELU) in
 *           ConvWithNorms.forward (OSF/src/models/basic/__init__.py:76-94) and
 *           UpsampleSkip.forward / decoder_step4 (OSF/src/models/basic/unet.py:18-35,130),
 *           i.e. every layer of UNetThreeFrame.forward (unet.py:131-166).
 * Activations are NHWC bf16 "planes": plane 0 = bf16(x), optional plane 1 = bf16(x - plane0)
 * (split-bf16; two planes give fp32-class products on the bf16 tensor cores).
 *   in      [in_planes][H_in][W_in][Cin_total] bf16; the conv reads channels
 *           [cin_off + g*cin_group_stride, +Cin) for group g (groups = frames sharing weights)
 *   wgt     [in_planes][Cout][ksize*ksize*Cin] bf16, K index = (ky*ksize + kx)*Cin + ci
 *   bias    [Cout] f32 or NULL (BatchNorm folded in by the host)
 *   out     NHWC, channels [cout_off + g*cout_group_stride, +Cout) of Cout_total; bf16 planes
 *           (out_planes 1|2) or fp32 (out_fp32 = 1)
 *   ksize 1|3 (padding ksize/2), stride 1|2, act 0 none | 1 exact-erf GELU
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
} himo_conv_desc;
int himo_conv2d_nhwc(const himo_conv_desc* desc, void* stream);
/* replaces: F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) of
 *           BilinearDecoder.forward (OSF/src/models/basic/unet.py:7-16); in [h][w][c] planes ->
 *           channels [cout_off, +c) of out [2h][2w][Cout_total] planes. */
int himo_upsample2x_nhwc(const void* in, int in_planes, long long in_plane_stride, int h, int w, int c,
                         void* out, int out_planes, long long out_plane_stride, int Cout_total,
                         int cout_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIMO_B200_H */
