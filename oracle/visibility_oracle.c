/* oracle/visibility_oracle.c -- TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C restatement of the reference's CPU (numba) visibility path, the authoritative variant
 * (README.md:122-123 of the reference tells users to avoid its own GPU path):
 *   oracle_project_equirect   <- camera_projection_cpu visibility.py:478-538 with
 *                                equirectangular_projection_cpu :150-182, norm_cpu :129-137,
 *                                field_of_view_cpu :395-435 (no image mask)
 *   oracle_splat_equirect     <- equirectangular_splat_cpu :630-704
 *   oracle_splat_pinhole      <- pinhole_splat_cpu :761-827
 *   oracle_zbuffer            <- visibility_from_splatting_cpu :1134-1195 (sequential z-buffer,
 *                                strict '<', ascending point order; exact=True re-rasterises centres)
 * numba typing reproduced: float32 arrays combined with Python floats evaluate in float64.
 * Pinned against outputs of the executed reference (tests/golden/zbuffer_*.npz, splat_pinhole.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double PI = 3.141592653589793;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* rotation R[9] row-major = pose_to_rotation_matrix_cpu(opk) computed by the caller in float32 */
void oracle_project_equirect(const float* xyz, const float* img_xyz, const float* R, int64_t n,
                             int W, int H, int crop_top, int crop_bottom, float r_min, float r_max,
                             float* dist, double* x_proj, double* y_proj, uint8_t* keep) {
  for (int64_t i = 0; i < n; ++i) {
    const float dx = xyz[3 * i] - img_xyz[0], dy = xyz[3 * i + 1] - img_xyz[1], dz = xyz[3 * i + 2] - img_xyz[2];
    const float d = sqrtf((dx * dx + dy * dy) + dz * dz);           /* norm_cpu, float32 */
    dist[i] = d;
    /* xyz_to_img.dot(R^T): numba hands the product to BLAS sgemm, whose k-loop is a chain of
     * fused multiply-adds in k order (measured on the executed reference: bit-equal for every
     * row); np.arctan2 / np.arccos on float32 are libm's atan2f / acosf */
    const float v0 = fmaf(dz, R[2], fmaf(dy, R[1], dx * R[0]));
    const float v1 = fmaf(dz, R[5], fmaf(dy, R[4], dx * R[3]));
    const float v2 = fmaf(dz, R[8], fmaf(dy, R[7], dx * R[6]));
    const float t = atan2f(v1, v0);
    const float p = acosf(v2 / d);
    double w = ((double)(W - 1) * (1.0 - (double)t / PI) / 2.0);
    double h = ((double)(H - 1) * (double)p / PI);
    w = w - floor(w / (double)W) * (double)W;                       /* python-style modulo */
    h = h - floor(h / (double)H) * (double)H;
    if (w != w) w = 0.0;
    if (h != h) h = 0.0;
    x_proj[i] = w; y_proj[i] = h;
    const int in_range = (r_min < d) && (d < r_max);
    const int in_fov = (0.0 <= w) && (w < (double)W) && ((double)crop_top <= h) && (h < (double)(H - crop_bottom));
    keep[i] = (uint8_t)(in_range && in_fov);
  }
}

/* Pinhole (camera 1) and fisheye (camera 3) projections, visibility.py:219-339.  The rigid
 * transform is passed as p = A . (xyz - t0) + t1:  scannet  A = R(c2w), t0 = 0, t1 = T(c2w) with
 * c2w = inv(extrinsic) (:233-236);  kitti360  A = R^T, t0 = T, t1 = 0 (:239-242, :305-308).
 * intr = fx, fy, cx, cy  |  xi, k1, k2, gamma1, gamma2, u0, v0. */
void oracle_project_camera(const float* xyz, const float* img_xyz, const float* A, const float* t0,
                           const float* t1, const float* intr, int camera, int64_t n, int W, int H,
                           int crop_top, int crop_bottom, float r_min, float r_max, float* dist,
                           double* x_proj, double* y_proj, uint8_t* keep) {
  for (int64_t i = 0; i < n; ++i) {
    const float dx = xyz[3 * i] - img_xyz[0], dy = xyz[3 * i + 1] - img_xyz[1], dz = xyz[3 * i + 2] - img_xyz[2];
    const float d = sqrtf((dx * dx + dy * dy) + dz * dz);
    dist[i] = d;
    const float q0 = xyz[3 * i] - t0[0], q1 = xyz[3 * i + 1] - t0[1], q2 = xyz[3 * i + 2] - t0[2];
    /* R @ xyz.T (sgemm: fused multiply-add chain in k order, see above), then "+ T" */
    const float p0 = fmaf(A[2], q2, fmaf(A[1], q1, A[0] * q0)) + t1[0];
    const float p1 = fmaf(A[5], q2, fmaf(A[4], q1, A[3] * q0)) + t1[1];
    const float p2 = fmaf(A[8], q2, fmaf(A[7], q1, A[6] * q0)) + t1[2];
    double x, y, z;
    if (camera == 1) {
      x = (double)(p0 * intr[0] / p2 + intr[2]);      /* float32 arithmetic, then astype(float64) */
      y = (double)(p1 * intr[1] / p2 + intr[3]);
      z = (double)p2;
    } else {
      const float nrm = sqrtf((p0 * p0 + p1 * p1) + p2 * p2);
      const double den = (double)nrm + 1e-4;
      double fx = (double)p0 / den, fy = (double)p1 / den;
      const double fz = (double)p2 / den;
      fx /= fz + (double)intr[0];
      fy /= fz + (double)intr[0];
      const double r2 = fx * fx + fy * fy, r4 = r2 * r2;
      x = (double)intr[3] * (1.0 + (double)intr[1] * r2 + (double)intr[2] * r4) * fx + (double)intr[5];
      y = (double)intr[4] * (1.0 + (double)intr[1] * r2 + (double)intr[2] * r4) * fy + (double)intr[6];
      z = (double)(nrm * p2) / fabs((double)p2 + 1e-4);
    }
    x_proj[i] = x; y_proj[i] = y;
    const int in_range = (r_min < d) && (d < r_max);
    const int in_fov = (0.0 <= x) && (x < (double)W) && ((double)crop_top <= y) && (y < (double)(H - crop_bottom)) && (0.0 < z);
    keep[i] = (uint8_t)(in_range && in_fov);
  }
}

static void finish_box(double xp, double yp, double wx, double wy, int W, int H, int crop_top,
                       int crop_bottom, int32_t* out) {
  /* np.round(value, 0, out=float32 array) then astype(int32): round half to even */
  int xa = (int)(float)rint(xp - wx / 2.0);
  int xb = (int)(float)rint(xp + wx / 2.0 + 1.0);
  int ya = (int)(float)rint(yp - wy / 2.0);
  int yb = (int)(float)rint(yp + wy / 2.0 + 1.0);
  const int y_min = crop_top, y_max = H - crop_bottom;
  out[0] = clampi(xa, 0, W - 1);
  out[1] = clampi(xb, 1, W);
  out[2] = clampi(ya, y_min, y_max - 1);
  out[3] = clampi(yb, y_min + 1, y_max);
}

void oracle_splat_equirect(const double* x_proj, const double* y_proj, const float* dist, int64_t m,
                           int W, int H, int crop_top, int crop_bottom, double voxel, double k_swell,
                           double d_swell, int32_t* splat) {
  const double ld = log(d_swell);
  for (int64_t i = 0; i < m; ++i) {
    const double d = (double)dist[i];
    const double aw = (1.0 + k_swell * exp(-d / ld)) * voxel / d;
    const double wy = aw * (double)H / PI;
    const double a = aw * (double)W / (2.0 * PI);
    const double b = PI / (double)H;
    const double wx = a / (sin(b * y_proj[i]) + 0.001);
    finish_box(x_proj[i], y_proj[i], wx, wy, W, H, crop_top, crop_bottom, splat + 4 * i);
  }
}

void oracle_splat_pinhole(const double* x_proj, const double* y_proj, const float* dist, int64_t m,
                          int W, int H, int crop_top, int crop_bottom, double voxel, double k_swell,
                          double d_swell, double fx, double fy, int32_t* splat) {
  const double ld = log(d_swell);
  for (int64_t i = 0; i < m; ++i) {
    const double d = (double)dist[i];
    const double swell = (1.0 + k_swell * exp(-d / ld)) * voxel / d;
    finish_box(x_proj[i], y_proj[i], swell * fx, swell * fy, W, H, crop_top, crop_bottom, splat + 4 * i);
  }
}

/* idx_map [W * Hc] int64, -1 = empty; layout [x][y] like the reference's depth_map[x, y] */
void oracle_zbuffer(const int32_t* splat, const float* dist, const double* x_proj, const double* y_proj,
                    int64_t m, int W, int H, int crop_top, int crop_bottom, int exact, int64_t* idx_map) {
  const int Hc = H - crop_top - crop_bottom;
  const int64_t npix = (int64_t)W * Hc;
  float d_max = 0.f;
  for (int64_t i = 0; i < m; ++i) if (dist[i] > d_max) d_max = dist[i];
  float* depth = (float*)malloc(sizeof(float) * (size_t)npix);
  for (int64_t p = 0; p < npix; ++p) { depth[p] = d_max + 2.f; idx_map[p] = -1; }
  for (int64_t i = 0; i < m; ++i) {
    const int xa = splat[4 * i], xb = splat[4 * i + 1];
    const int ya = splat[4 * i + 2] - crop_top, yb = splat[4 * i + 3] - crop_top;
    for (int x = xa; x < xb; ++x)
      for (int y = ya; y < yb; ++y)
        if (dist[i] < depth[(int64_t)x * Hc + y]) { depth[(int64_t)x * Hc + y] = dist[i]; idx_map[(int64_t)x * Hc + y] = i; }
  }
  if (exact) {
    uint8_t* seen = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
    for (int64_t p = 0; p < npix; ++p) if (idx_map[p] >= 0) seen[idx_map[p]] = 1;
    for (int64_t p = 0; p < npix; ++p) idx_map[p] = -1;
    for (int64_t i = 0; i < m; ++i) {      /* ascending: the highest seen index wins a shared centre */
      if (!seen[i]) continue;
      const int x = (int)x_proj[i], y = (int)y_proj[i] - crop_top;
      idx_map[(int64_t)x * Hc + y] = i;
    }
    free(seen);
  }
  free(depth);
}
