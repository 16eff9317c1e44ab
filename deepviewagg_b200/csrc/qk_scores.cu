// Ragged per-group query.key compatibilities (QKVBimodalCSRPool, pooling.py:499-512).
// The reference expands the per-point queries to views with repeat_interleave (pooling.py:500)
// and reduces a [V,G,D] product; here one thread owns one (point, g*D+d) column, keeps the
// query element in a register and walks the point's views, so Q is read once per point and
// the [V,G*D] expansion never exists.  fp32 throughout (scores feed the softmax statistics).
//   bytes fwd: V*(G*D*4 + G*4) + N*(G*D*4 + 8);  bwd adds V*G*D*4 + N*G*D*4 writes.
#include "dva_common.cuh"

namespace dva {

// fwd: one thread per (point, group); D-loop inside (D is 8 in all shipped configs)
__global__ void __launch_bounds__(256)
qk_scores_fwd_kernel(const float* __restrict__ keys, const float* __restrict__ queries,
                     const int64_t* __restrict__ ptr, float* __restrict__ compat, int64_t N,
                     int G, int D, float scale) {
  const int64_t total = N * G;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / G;
    const int g = (int)(t - i * G);
    const int64_t p0 = ptr[i], p1 = ptr[i + 1];
    const float* q = queries + i * (int64_t)G * D + g * D;
    for (int64_t v = p0; v < p1; ++v) {
      const float* k = keys + v * (int64_t)G * D + g * D;
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc += k[d] * __ldg(q + d);   // reference order: sum over d
      compat[v * G + g] = acc * scale;
    }
  }
}

// bwd: one thread per (point, g*D+d)
__global__ void __launch_bounds__(256)
qk_scores_bwd_kernel(const float* __restrict__ keys, const float* __restrict__ queries,
                     const int64_t* __restrict__ ptr, const float* __restrict__ gcompat,
                     float* __restrict__ gkeys, float* __restrict__ gqueries, int64_t N, int G,
                     int D, float scale) {
  const int GD = G * D;
  const int64_t total = N * GD;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / GD;
    const int j = (int)(t - i * GD);
    const int g = j / D;
    const int64_t p0 = ptr[i], p1 = ptr[i + 1];
    const float q = queries[i * GD + j];
    float gq = 0.f;
    for (int64_t v = p0; v < p1; ++v) {
      const float gc = gcompat[v * G + g] * scale;
      gkeys[v * GD + j] = gc * q;
      gq = fmaf(gc, keys[v * GD + j], gq);
    }
    gqueries[i * GD + j] = gq;
  }
}

static inline int qk_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace dva

using namespace dva;

extern "C" int dva_qk_scores_fwd(const float* keys, const float* queries, const int64_t* ptr,
                                 float* compat, int64_t N, int64_t V, int64_t G, int64_t D,
                                 float scale, void* stream) {
  if (N < 0 || V < 0 || G < 1 || D < 1) return fail(DVA_EINVAL, "qk_scores_fwd: bad sizes");
  if (N == 0 || V == 0) return DVA_OK;
  if (!keys || !queries || !ptr || !compat) return fail(DVA_EINVAL, "qk_scores_fwd: null pointer");
  qk_scores_fwd_kernel<<<qk_grid(N * G), 256, 0, (cudaStream_t)stream>>>(keys, queries, ptr, compat, N, (int)G, (int)D, scale);
  return check_launch("qk_scores_fwd");
}

extern "C" int dva_qk_scores_bwd(const float* keys, const float* queries, const int64_t* ptr,
                                 const float* grad_compat, float* grad_keys, float* grad_queries,
                                 int64_t N, int64_t V, int64_t G, int64_t D, float scale,
                                 void* stream) {
  if (N < 0 || V < 0 || G < 1 || D < 1) return fail(DVA_EINVAL, "qk_scores_bwd: bad sizes");
  if (N == 0) return DVA_OK;
  if (!queries || !ptr || !grad_queries || (V > 0 && (!keys || !grad_compat || !grad_keys)))
    return fail(DVA_EINVAL, "qk_scores_bwd: null pointer");
  qk_scores_bwd_kernel<<<qk_grid(N * G * D), 256, 0, (cudaStream_t)stream>>>(
      keys, queries, ptr, grad_compat, grad_keys, grad_queries, N, (int)G, (int)D, scale);
  return check_launch("qk_scores_bwd");
}
