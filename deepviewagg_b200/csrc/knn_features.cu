// Neighbourhood-based mapping features: the density and occlusion viewing
// conditions of NeighborhoodBasedMappingFeatures (core/data_transform/multimodal/image.py:431-612).
//
//   1. exact k-NN of every point among all points (:504-514: KeOps `argKmin` of the squared
//      distances; the FAISS branch is an approximate search and is not reproduced) on a uniform
//      grid: points are counting-sorted by cell (host side: CUB sort through torch), one thread
//      per query walks cubic shells of cells outwards and stops as soon as its k-th best distance
//      is inside the visited cube.  Squared distances are (dx*dx + dy*dy) + dz*dz in fp32 without
//      FMA contraction; ties are ordered by point index, so the result is a deterministic
//      function of the input (the oracle restates exactly this order).
//   2. density (:521-546): (k+1) / (3.1416 d_k^2) / (1/voxel^2) per point, expanded to its views.
//   3. occlusion (:556-588): (1 + #neighbours seen by the same image) / (k+1) per view.  The
//      reference materialises a dense bool [n_points, n_images] table and k fancy-index gathers;
//      here a view scans the (short) image lists of its point's neighbours in the view CSR.
#include "dva_common.cuh"

namespace dva {

constexpr int kKnnMax = 64;
constexpr int kKnnMaxShells = 6;   // 13^3 cells; a query still open after that scans all points

__global__ void __launch_bounds__(256)
knn_cell_ids_kernel(const float* __restrict__ xyz, int64_t* __restrict__ cell, int64_t n, float ox,
                    float oy, float oz, float inv_cs, int gx, int gy, int gz) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cx = min(max((int)floorf((xyz[3 * i + 0] - ox) * inv_cs), 0), gx - 1);
    const int cy = min(max((int)floorf((xyz[3 * i + 1] - oy) * inv_cs), 0), gy - 1);
    const int cz = min(max((int)floorf((xyz[3 * i + 2] - oz) * inv_cs), 0), gz - 1);
    cell[i] = ((int64_t)cz * gy + cy) * gx + cx;
  }
}

// (d2, id) lexicographic order
__device__ __forceinline__ bool knn_less(float d, int64_t i, float d2, int64_t i2) {
  return d < d2 || (d == d2 && i < i2);
}

// xyz_s / cell_s / order: points in cell-sorted order (order[j] = original index of sorted slot j)
__global__ void __launch_bounds__(128)
knn_grid_kernel(const float* __restrict__ xyz_s, const int64_t* __restrict__ cell_s,
                const int64_t* __restrict__ order, const int64_t* __restrict__ cell_ptr, int64_t n,
                int k, float ox, float oy, float oz, float cs, int gx, int gy, int gz,
                int64_t* __restrict__ nbr, float* __restrict__ d2out) {
  float bd[kKnnMax];
  int64_t bi[kKnnMax];
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const float px = xyz_s[3 * q], py = xyz_s[3 * q + 1], pz = xyz_s[3 * q + 2];
    const int64_t c = cell_s[q];
    const int cx = (int)(c % gx), cy = (int)((c / gx) % gy), cz = (int)(c / ((int64_t)gx * gy));
    // distance from the query to the nearest face of its own cell (shell r adds r * cs)
    const float fx = px - (ox + cx * cs), fy = py - (oy + cy * cs), fz = pz - (oz + cz * cs);
    const float inner = fmaxf(fminf(fminf(fminf(fx, cs - fx), fminf(fy, cs - fy)), fminf(fz, cs - fz)), 0.f);
    int cnt = 0;
    auto offer = [&](float d2, int64_t id) {
      if (cnt == k && !knn_less(d2, id, bd[k - 1], bi[k - 1])) return;
      int t = (cnt < k) ? cnt : k - 1;                    // insertion into the sorted prefix
      while (t > 0 && knn_less(d2, id, bd[t - 1], bi[t - 1])) { bd[t] = bd[t - 1]; bi[t] = bi[t - 1]; --t; }
      bd[t] = d2; bi[t] = id;
      if (cnt < k) ++cnt;
    };
    const int rmax = min(max(gx, max(gy, gz)), kKnnMaxShells);
    bool done = false;
    for (int r = 0; r <= rmax; ++r) {
      for (int dz = -r; dz <= r; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= gz) continue;
        for (int dy = -r; dy <= r; ++dy) {
          const int y = cy + dy;
          if (y < 0 || y >= gy) continue;
          const bool face = (dz == -r || dz == r || dy == -r || dy == r);
          // a full x-row of cells on the shell's faces, else only its two end cells
          for (int part = 0; part < (face || r == 0 ? 1 : 2); ++part) {
            int x0, x1;
            if (face || r == 0) { x0 = cx - r; x1 = cx + r; }
            else { x0 = x1 = (part == 0 ? cx - r : cx + r); }
            if (x1 < 0 || x0 >= gx) continue;
            x0 = max(x0, 0); x1 = min(x1, gx - 1);
            const int64_t row = ((int64_t)z * gy + y) * gx;
            const int64_t j0 = cell_ptr[row + x0], j1 = cell_ptr[row + x1 + 1];
            for (int64_t j = j0; j < j1; ++j) {
              const float dx = __fsub_rn(px, xyz_s[3 * j]), dyy = __fsub_rn(py, xyz_s[3 * j + 1]),
                          dzz = __fsub_rn(pz, xyz_s[3 * j + 2]);
              const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dyy, dyy)), __fmul_rn(dzz, dzz));
              offer(d2, order[j]);
            }
          }
        }
      }
      if (cnt == k) {
        // everything outside the visited cube is farther than `reach` (margin for the rounding of
        // the cell assignment)
        const float reach = fmaxf(r * cs + inner - 1e-4f * cs, 0.f);
        if (bd[k - 1] <= reach * reach) { done = true; break; }
      }
    }
    if (!done && rmax < max(gx, max(gy, gz))) {
      // isolated point (outlier, or a cell size far too small here): exhaustive scan
      cnt = 0;
      for (int64_t j = 0; j < n; ++j) {
        const float dx = __fsub_rn(px, xyz_s[3 * j]), dyy = __fsub_rn(py, xyz_s[3 * j + 1]),
                    dzz = __fsub_rn(pz, xyz_s[3 * j + 2]);
        offer(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dyy, dyy)), __fmul_rn(dzz, dzz)), order[j]);
      }
    }
    const int64_t me = order[q];
    for (int t = 0; t < k; ++t) {
      nbr[me * k + t] = (t < cnt) ? bi[t] : -1;
      if (d2out != nullptr) d2out[me * k + t] = (t < cnt) ? bd[t] : INFINITY;
    }
  }
}

// one thread per view: density of its point + occlusion of the view, for every k in klist
__global__ void __launch_bounds__(256)
neighborhood_features_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nbr, int kmax,
                             const int64_t* __restrict__ vptr, const int64_t* __restrict__ images,
                             const int64_t* __restrict__ view_point, const int* __restrict__ klist,
                             int nk, float voxel_density, int do_density, int do_occlusion,
                             float* __restrict__ out, int64_t V) {
  const int width = (do_density ? nk : 0) + (do_occlusion ? nk : 0);
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = view_point[v];
    float* __restrict__ o = out + v * width;
    int col = 0;
    if (do_density) {
      const float px = xyz[3 * p], py = xyz[3 * p + 1], pz = xyz[3 * p + 2];
      for (int a = 0; a < nk; ++a) {
        const int k = klist[a];
        const int64_t q = nbr[p * kmax + (k - 1)];
        // d2_max = ((xyz - xyz[neighbors[:, k-1]])**2).sum(1)                    :527
        const float dx = __fsub_rn(px, xyz[3 * q]), dy = __fsub_rn(py, xyz[3 * q + 1]), dz = __fsub_rn(pz, xyz[3 * q + 2]);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        // density = ((k+1) / (3.1416 d2)) / (1/voxel^2); NaN -> 1                :532-537
        float den = __fdiv_rn(__fdiv_rn((float)(k + 1), __fmul_rn(3.1416f, d2)), voxel_density);
        if (den != den) den = 1.f;
        o[col++] = den;
      }
    }
    if (do_occlusion) {
      const int64_t img = images[v];
      float seen = 1.f;                                   // the point itself                :575
      int a = 0;
      for (int i = 0; i < kmax && a < nk; ++i) {
        const int64_t q = nbr[p * kmax + i];
        bool hit = false;
        for (int64_t w = vptr[q]; w < vptr[q + 1]; ++w) hit |= (images[w] == img);
        seen += hit ? 1.f : 0.f;
        while (a < nk && klist[a] == i + 1) { o[col + a] = __fdiv_rn(seen, (float)(klist[a] + 1)); ++a; }   // :584
      }
    }
  }
}

static inline int k_grid(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace dva

using namespace dva;

extern "C" int dva_knn_cell_ids(const float* xyz, int64_t* cell, int64_t n, float ox, float oy, float oz,
                                float cell_size, int gx, int gy, int gz, void* stream) {
  if (n < 0 || gx < 1 || gy < 1 || gz < 1 || !(cell_size > 0.f)) return fail(DVA_EINVAL, "knn_cell_ids: bad sizes");
  if (n == 0) return DVA_OK;
  if (!xyz || !cell) return fail(DVA_EINVAL, "knn_cell_ids: null pointer");
  knn_cell_ids_kernel<<<k_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(xyz, cell, n, ox, oy, oz, 1.f / cell_size, gx, gy, gz);
  return check_launch("knn_cell_ids");
}

extern "C" int dva_knn_grid(const float* xyz_sorted, const int64_t* cell_sorted, const int64_t* order,
                            const int64_t* cell_ptr, int64_t n, int k, float ox, float oy, float oz,
                            float cell_size, int gx, int gy, int gz, int64_t* neighbors, float* dist2,
                            void* stream) {
  if (n < 0 || gx < 1 || gy < 1 || gz < 1 || !(cell_size > 0.f)) return fail(DVA_EINVAL, "knn_grid: bad sizes");
  if (k < 1 || k > kKnnMax) return fail(DVA_EUNSUPPORTED, "knn_grid: k must be in [1, 64]");
  if (n == 0) return DVA_OK;
  if (!xyz_sorted || !cell_sorted || !order || !cell_ptr || !neighbors) return fail(DVA_EINVAL, "knn_grid: null pointer");
  knn_grid_kernel<<<k_grid(n, 128), 128, 0, (cudaStream_t)stream>>>(xyz_sorted, cell_sorted, order, cell_ptr, n, k, ox, oy,
                                                                     oz, cell_size, gx, gy, gz, neighbors, dist2);
  return check_launch("knn_grid");
}

extern "C" int dva_neighborhood_features(const float* xyz, const int64_t* neighbors, int kmax,
                                         const int64_t* view_ptr, const int64_t* images,
                                         const int64_t* view_point, const int32_t* klist, int nk,
                                         double voxel, int density, int occlusion, float* out,
                                         int64_t N, int64_t V, void* stream) {
  if (N < 0 || V < 0 || kmax < 1 || nk < 1) return fail(DVA_EINVAL, "neighborhood_features: bad sizes");
  if (!density && !occlusion) return fail(DVA_EINVAL, "neighborhood_features: nothing to compute");
  if (V == 0) return DVA_OK;
  if (!xyz || !neighbors || !view_ptr || !images || !view_point || !klist || !out)
    return fail(DVA_EINVAL, "neighborhood_features: null pointer");
  // voxel_density = 1 / voxel**2 evaluated in double, then used as an fp32 scalar            :531
  const float voxel_density = (float)(1.0 / (voxel * voxel));
  neighborhood_features_kernel<<<k_grid(V, 256), 256, 0, (cudaStream_t)stream>>>(
      xyz, neighbors, kmax, view_ptr, images, view_point, klist, nk, voxel_density, density, occlusion, out, V);
  return check_launch("neighborhood_features");
}
