/*
 * oracle/leaf_ops.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's native leaf operators on the hot path,
 * used as the parity checker for the sm_100a kernels in himo_b200/csrc and as the
 * CPU baseline of bench.py.  Nothing here is copied from the reference; each
 * function states the reference lines whose arithmetic it follows.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/__init__.py).
 * -ffp-contract=off matters: every fused multiply-add below is spelled out with
 * fmaf() so the rounding sequence is the one the reference's device code has.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

/* float -> int conversion with the CUDA semantics of cvt.rzi.s32.f32 (saturating,
 * NaN -> 0).  The reference kernel assigns floorf(...) to an int on the device
 * (OSF/assets/cuda/mmcv/voxelization_cuda_kernel.cuh:26,32,39). */
static inline int f2i_sat(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT_MAX;
  if (f <= -2147483648.0f) return INT_MIN;
  return (int)f;
}

/* dynamic_voxelize_forward
 * follows OSF/assets/cuda/mmcv/voxelization_cuda_kernel.cuh:13-50 (per-point body)
 * and OSF/assets/cuda/mmcv/voxelization_cuda.cu:259-271 (grid = round((max-min)/voxel)).
 * coors must be pre-zeroed by the caller (OSF/assets/cuda/mmcv/voxelize.py:78):
 * the early exits only overwrite the leading entries. */
void oracle_dynamic_voxelize(const float *points, int n, int num_features,
                             const float *voxel_size, const float *coors_range,
                             int32_t *coors) {
  const float vx = voxel_size[0], vy = voxel_size[1], vz = voxel_size[2];
  const float x_min = coors_range[0], y_min = coors_range[1], z_min = coors_range[2];
  const float x_max = coors_range[3], y_max = coors_range[4], z_max = coors_range[5];
  const int grid_x = (int)round((double)((x_max - x_min) / vx));
  const int grid_y = (int)round((double)((y_max - y_min) / vy));
  const int grid_z = (int)round((double)((z_max - z_min) / vz));
  for (int i = 0; i < n; ++i) {
    const float *p = points + (size_t)i * num_features;
    int32_t *c = coors + (size_t)i * 3;
    int c_x = f2i_sat(floorf((p[0] - x_min) / vx));
    if (c_x < 0 || c_x >= grid_x) { c[0] = -1; continue; }
    int c_y = f2i_sat(floorf((p[1] - y_min) / vy));
    if (c_y < 0 || c_y >= grid_y) { c[0] = -1; c[1] = -1; continue; }
    int c_z = f2i_sat(floorf((p[2] - z_min) / vz));
    if (c_z < 0 || c_z >= grid_z) { c[0] = -1; c[1] = -1; c[2] = -1; }
    else { c[0] = c_z; c[1] = c_y; c[2] = c_x; }
  }
}

/* ---- dynamic_point_to_voxel_forward ------------------------------------------------
 * follows OSF/assets/cuda/mmcv/scatter_points_cuda.cu:9-66:
 *   rows with any negative entry -> (-1,-1,-1)                      (:22)
 *   unique rows, lexicographically sorted, inverse map and counts   (:24-27)
 *   a leading negative row is dropped and the map shifted by -1     (:29-34)
 *   sum (or max) of the features per voxel, mean = sum / count      (:46-61)
 * and the per-point body of OSF/assets/cuda/mmcv/scatter_points_cuda_kernel.cuh:91-112.
 * The device kernel accumulates with fp32 atomics in unspecified order; this
 * restatement takes the points in ascending index order (accum_mode 0) or sums in
 * double and rounds once (accum_mode 1, the order-free limit of the same sum).
 * Returns M (number of voxels); outputs are caller-allocated for the worst case M<=n. */
typedef struct { int64_t c0, c1, c2; int idx; } row_t;
static int row_cmp(const void *a, const void *b) {
  const row_t *x = (const row_t *)a, *y = (const row_t *)b;
  if (x->c0 != y->c0) return x->c0 < y->c0 ? -1 : 1;
  if (x->c1 != y->c1) return x->c1 < y->c1 ? -1 : 1;
  if (x->c2 != y->c2) return x->c2 < y->c2 ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int oracle_dynamic_point_to_voxel(const float *feats, const int64_t *coors, int n, int c,
                                  int reduce_type /*0 sum,1 mean,2 max*/, int accum_mode,
                                  float *voxel_feats, int64_t *voxel_coors,
                                  int32_t *point2voxel, int32_t *voxel_count) {
  if (n == 0) return 0;
  row_t *rows = (row_t *)malloc(sizeof(row_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    int64_t a = coors[3 * (size_t)i], b = coors[3 * (size_t)i + 1], d = coors[3 * (size_t)i + 2];
    if (a < 0 || b < 0 || d < 0) a = b = d = -1;
    rows[i].c0 = a; rows[i].c1 = b; rows[i].c2 = d; rows[i].idx = i;
  }
  qsort(rows, (size_t)n, sizeof(row_t), row_cmp);
  int m = -1;  /* unique-row ordinal, counting the (-1,-1,-1) row if present */
  int has_neg = rows[0].c0 < 0;
  int *row_of = (int *)malloc(sizeof(int) * (size_t)n);
  for (int k = 0; k < n; ++k) {
    if (k == 0 || rows[k].c0 != rows[k - 1].c0 || rows[k].c1 != rows[k - 1].c1 ||
        rows[k].c2 != rows[k - 1].c2)
      ++m;
    row_of[rows[k].idx] = m;
  }
  int M = m + 1 - has_neg;
  for (int k = 0; k < n; ++k) {
    int v = row_of[rows[k].idx] - has_neg;
    if (v >= 0) { voxel_coors[3 * (size_t)v] = rows[k].c0; voxel_coors[3 * (size_t)v + 1] = rows[k].c1;
                  voxel_coors[3 * (size_t)v + 2] = rows[k].c2; }
  }
  for (int v = 0; v < M; ++v) voxel_count[v] = 0;
  double *acc = NULL;
  if (accum_mode == 1 && reduce_type != 2) acc = (double *)calloc((size_t)M * c, sizeof(double));
  for (size_t k = 0; k < (size_t)M * c; ++k) voxel_feats[k] = reduce_type == 2 ? -INFINITY : 0.0f;
  for (int i = 0; i < n; ++i) {          /* ascending point index */
    int v = row_of[i] - has_neg;
    point2voxel[i] = v;                   /* -1 for invalid points */
    if (v < 0) continue;
    voxel_count[v] += 1;
    const float *f = feats + (size_t)i * c;
    float *o = voxel_feats + (size_t)v * c;
    if (reduce_type == 2) { for (int k = 0; k < c; ++k) if (f[k] > o[k]) o[k] = f[k]; }
    else if (acc)         { for (int k = 0; k < c; ++k) acc[(size_t)v * c + k] += (double)f[k]; }
    else                  { for (int k = 0; k < c; ++k) o[k] = o[k] + f[k]; }
  }
  if (acc) for (size_t k = 0; k < (size_t)M * c; ++k) voxel_feats[k] = (float)acc[k];
  if (reduce_type == 1)
    for (int v = 0; v < M; ++v)
      for (int k = 0; k < c; ++k) voxel_feats[(size_t)v * c + k] /= (float)voxel_count[v];
  free(acc); free(row_of); free(rows);
  return M;
}

/* ---- chamfer3D.forward: exact 1-NN, one direction ---------------------------------
 * follows OSF/assets/cuda/chamfer3D/chamfer3D.cu:33-83: best = 1e20, best_i = -1,
 * candidates scanned in ascending index, strict '<' (lowest index wins ties), squared
 * L2 in fp32.  nvcc (12.9, sm_100a) contracts the reference's expression
 *   (x1-x0)*(x1-x0) + (y1-y0)*(y1-y0) + (z1-z0)*(z1-z0)
 * into  fma(dz,dz, fma(dx,dx, dy*dy))  -- FMUL on dy, FFMA dx, FFMA dz in every copy of the loop
 * body in the SASS of oracle/_ref/chamfer3D.so (oracle/build_ref.py) -- spelled out here. */
void oracle_nn_bruteforce(const float *q, int nq, const float *r, int nr, float *dist,
                          int32_t *idx) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; ++i) {
    const float x0 = q[3 * (size_t)i], y0 = q[3 * (size_t)i + 1], z0 = q[3 * (size_t)i + 2];
    float best = 1e20f;
    int best_i = -1;
    for (int j = 0; j < nr; ++j) {
      float dx = r[3 * (size_t)j] - x0, dy = r[3 * (size_t)j + 1] - y0, dz = r[3 * (size_t)j + 2] - z0;
      float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      if (d < best) { best = d; best_i = j; }
    }
    dist[i] = best;
    idx[i] = best_i;
  }
}

/* chamfer3D.backward, follows OSF/assets/cuda/chamfer3D/chamfer3D.cu:107-154.
 * grad_pc0 / grad_pc1 are accumulated into (caller pre-zeroes).  Sequential index
 * order stands in for the device's unordered fp32 atomics. */
void oracle_chamfer_backward(const float *pc0, int n0, const float *pc1, int n1,
                             const int32_t *idx0, const int32_t *idx1, const float *g0,
                             const float *g1, float *grad_pc0, float *grad_pc1) {
  for (int dir = 0; dir < 2; ++dir) {
    const float *a = dir ? pc1 : pc0, *b = dir ? pc0 : pc1;
    const int32_t *id = dir ? idx1 : idx0;
    const float *gd = dir ? g1 : g0;
    float *ga = dir ? grad_pc1 : grad_pc0, *gb = dir ? grad_pc0 : grad_pc1;
    int n = dir ? n1 : n0;
    for (int i = 0; i < n; ++i) {
      int j = id[i];
      float g = gd[i] * 2;
      for (int k = 0; k < 3; ++k) {
        float t = g * (a[3 * (size_t)i + k] - b[3 * (size_t)j + k]);
        ga[3 * (size_t)i + k] += t;
        gb[3 * (size_t)j + k] += -t;
      }
    }
  }
}

/* ---- FastGeodis.generalised_geodesic3d, lamb = 0 (pure Euclidean), restated ----------
 * Third-party dependency, NOT under /root/reference: FastGeodis (unpinned `pip install
 * FastGeodis` in OSF/Dockerfile:36; "FastGeodis==1.0.4" in a comment at
 * OSF/src/models/fastnsf.py:17).  Call site: OSF/src/models/fastnsf.py:52-57 with
 * image = 0, softmask = 1 except 0 at occupied cells, spacing = [0.1]*3, v = 1e10,
 * lamb = 0.0, iterations = 1.  PARITY UNPINNED: no reference test or golden vector
 * pins these values; this follows the published FastGeodis raster-scan algorithm
 * (Asad et al., "FastGeodis: Fast Generalised Geodesic Distance Transform", JOSS 2022):
 *   dist = v * softmask; per iteration, for each of the three axes in turn (the volume is
 *   permuted so the swept axis is the leading one), a forward then a backward pass over
 *   the planes of that axis; every voxel takes the minimum of itself and the 9 voxels of
 *   the previous plane (in-plane offsets -1,0,+1 in both other axes) plus the Euclidean
 *   length of the offset, sqrt(sp_a^2 + (dh*sp_h)^2 + (dw*sp_w)^2) (lamb = 0 drops the
 *   image term).
 * Axis order: FastGeodis sweeps depth (dim 2 of the 5-D tensor = our axis 0), then
 * height (axis 1), then width (axis 2).
 * d: [n0, n1, n2] row-major, in/out. */
static void dt_pass(float *d, int n0, int n1, int n2, int axis, const float *sp) {
  /* strides */
  const size_t s[3] = {(size_t)n1 * n2, (size_t)n2, 1};
  const int n[3] = {n0, n1, n2};
  const int a = axis, h = (axis + 1) % 3, w = (axis + 2) % 3;
  float l[3][3];
  for (int dh = -1; dh <= 1; ++dh)
    for (int dw = -1; dw <= 1; ++dw)
      l[dh + 1][dw + 1] = sqrtf(sp[a] * sp[a] + (float)(dh * dh) * sp[h] * sp[h] +
                                (float)(dw * dw) * sp[w] * sp[w]);
  for (int dir = 0; dir < 2; ++dir) {
    for (int step = 1; step < n[a]; ++step) {
      int p = dir == 0 ? step : n[a] - 1 - step;
      int pp = dir == 0 ? p - 1 : p + 1;
#pragma omp parallel for schedule(static)
      for (int ih = 0; ih < n[h]; ++ih)
        for (int iw = 0; iw < n[w]; ++iw) {
          size_t o = (size_t)p * s[a] + (size_t)ih * s[h] + (size_t)iw * s[w];
          float best = d[o];
          for (int dh = -1; dh <= 1; ++dh) {
            int jh = ih + dh;
            if (jh < 0 || jh >= n[h]) continue;
            for (int dw = -1; dw <= 1; ++dw) {
              int jw = iw + dw;
              if (jw < 0 || jw >= n[w]) continue;
              float c = d[(size_t)pp * s[a] + (size_t)jh * s[h] + (size_t)jw * s[w]] +
                        l[dh + 1][dw + 1];
              if (c < best) best = c;
            }
          }
          d[o] = best;
        }
    }
  }
}

void oracle_geodesic3d_euclid(float *d, int n0, int n1, int n2, const float *spacing,
                              int iterations) {
  for (int it = 0; it < iterations; ++it)
    for (int axis = 0; axis < 3; ++axis) dt_pass(d, n0, n1, n2, axis, spacing);
}
