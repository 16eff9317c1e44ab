// Fused feature-map pixel gather + atomic (pixel -> view) pool.
//   reference: x = self.x[(img_id per pixel, ..., py, px)]  (image.py:1285, 1871-1885) makes a
//   [P,C] copy out of the NCHW map, then BimodalCSRPool reduces it over the atomic CSR
//   (modules.py:497-500 -> pooling.py:63).  Here the [P,C] intermediate never exists: one thread
//   owns one (view, channel) output, walks the view's pixels and reads the map directly.
// With a channels-last map ([B,H,W,C]) the C channels of a pixel are one contiguous run, so a
// warp reads 32 consecutive channels per pixel (coalesced); with the reference's NCHW layout
// every element is H*W apart (32-byte sector per 4-byte element) -- supported for drop-in use,
// channels-last is the fast path.
#include "dva_common.cuh"

namespace dva {

template <typename T, typename PIX, bool CL, int RED>
__global__ void __launch_bounds__(256)
gather_pool_fwd_kernel(const T* __restrict__ fmap, const int64_t* __restrict__ img,
                       const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                       T* __restrict__ out, int64_t* __restrict__ arg, int64_t C, int64_t H,
                       int64_t W, int64_t Vw, int64_t P) {
  const int64_t total = Vw * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = t / C, c = t - w * C;
    const int64_t p0 = aptr[w], p1 = aptr[w + 1];
    const int64_t b = img[w];
    float acc = 0.f;
    int64_t best = P;
    for (int64_t p = p0; p < p1; ++p) {
      const int64_t px = (int64_t)pix[2 * p], py = (int64_t)pix[2 * p + 1];
      const int64_t off = CL ? (((b * H + py) * W + px) * C + c) : (((b * C + c) * H + py) * W + px);
      const float v = Cvt<T>::to_f(fmap[off]);
      if (RED == DVA_SUM || RED == DVA_MEAN) acc += v;
      else if (p == p0 || (RED == DVA_MAX ? v > acc : v < acc)) { acc = v; best = p; }
    }
    if (RED == DVA_MEAN) acc /= (float)((p1 - p0) > 0 ? (p1 - p0) : 1);
    out[t] = Cvt<T>::from_f(acc);
    if ((RED == DVA_MAX || RED == DVA_MIN) && arg != nullptr) arg[t] = best;
  }
}

template <typename T, typename PIX, bool CL, int RED>
__global__ void __launch_bounds__(256)
gather_pool_bwd_kernel(const T* __restrict__ gout, const int64_t* __restrict__ img,
                       const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                       const int64_t* __restrict__ arg, float* __restrict__ gfmap, int64_t C,
                       int64_t H, int64_t W, int64_t Vw) {
  const int64_t total = Vw * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = t / C, c = t - w * C;
    const int64_t p0 = aptr[w], p1 = aptr[w + 1];
    if (p1 <= p0) continue;
    const int64_t b = img[w];
    float g = Cvt<T>::to_f(gout[t]);
    if (RED == DVA_MEAN) g /= (float)(p1 - p0);
    if (RED == DVA_MAX || RED == DVA_MIN) {
      const int64_t p = arg[t];
      const int64_t px = (int64_t)pix[2 * p], py = (int64_t)pix[2 * p + 1];
      const int64_t off = CL ? (((b * H + py) * W + px) * C + c) : (((b * C + c) * H + py) * W + px);
      atomicAdd(gfmap + off, g);
    } else {
      for (int64_t p = p0; p < p1; ++p) {
        const int64_t px = (int64_t)pix[2 * p], py = (int64_t)pix[2 * p + 1];
        const int64_t off = CL ? (((b * H + py) * W + px) * C + c) : (((b * C + c) * H + py) * W + px);
        atomicAdd(gfmap + off, g);
      }
    }
  }
}

static inline int gp_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

template <typename T, typename PIX, bool CL>
static int gp_fwd_red(const void* fmap, const int64_t* img, const void* pix, const int64_t* aptr,
                      void* out, int64_t* arg, int64_t C, int64_t H, int64_t W, int64_t Vw,
                      int64_t P, int reduce, cudaStream_t st) {
  const int grid = gp_grid(Vw * C);
#define GP_F(R) gather_pool_fwd_kernel<T, PIX, CL, R><<<grid, 256, 0, st>>>((const T*)fmap, img, (const PIX*)pix, aptr, (T*)out, arg, C, H, W, Vw, P)
  switch (reduce) {
    case DVA_SUM: GP_F(DVA_SUM); break;
    case DVA_MEAN: GP_F(DVA_MEAN); break;
    case DVA_MAX: GP_F(DVA_MAX); break;
    case DVA_MIN: GP_F(DVA_MIN); break;
    default: return fail(DVA_EINVAL, "gather_pool_fwd: unknown reduce");
  }
#undef GP_F
  return check_launch("gather_pool_fwd");
}

template <typename T, typename PIX, bool CL>
static int gp_bwd_red(const void* gout, const int64_t* img, const void* pix, const int64_t* aptr,
                      const int64_t* arg, float* gfmap, int64_t C, int64_t H, int64_t W,
                      int64_t Vw, int reduce, cudaStream_t st) {
  const int grid = gp_grid(Vw * C);
#define GP_B(R) gather_pool_bwd_kernel<T, PIX, CL, R><<<grid, 256, 0, st>>>((const T*)gout, img, (const PIX*)pix, aptr, arg, gfmap, C, H, W, Vw)
  switch (reduce) {
    case DVA_SUM: GP_B(DVA_SUM); break;
    case DVA_MEAN: GP_B(DVA_MEAN); break;
    case DVA_MAX: GP_B(DVA_MAX); break;
    case DVA_MIN: GP_B(DVA_MIN); break;
    default: return fail(DVA_EINVAL, "gather_pool_bwd: unknown reduce");
  }
#undef GP_B
  return check_launch("gather_pool_bwd");
}

}  // namespace dva

using namespace dva;

#define GP_DISPATCH(FN, ...)                                                              \
  do {                                                                                    \
    if (channels_last) {                                                                  \
      if (pix_is_i16) return FN<T, int16_t, true>(__VA_ARGS__);                           \
      return FN<T, int32_t, true>(__VA_ARGS__);                                           \
    }                                                                                     \
    if (pix_is_i16) return FN<T, int16_t, false>(__VA_ARGS__);                            \
    return FN<T, int32_t, false>(__VA_ARGS__);                                            \
  } while (0)

extern "C" int dva_gather_pool_fwd(const void* fmap, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   void* out, int64_t* arg, int64_t B, int64_t C, int64_t H,
                                   int64_t W, int64_t Vw, int64_t P, int reduce, int dtype,
                                   void* stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || Vw < 0 || P < 0) return fail(DVA_EINVAL, "gather_pool_fwd: negative size");
  if (Vw == 0 || C == 0) return DVA_OK;
  if (!aptr || !out || !img || (P > 0 && (!fmap || !pix))) return fail(DVA_EINVAL, "gather_pool_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: { using T = float; GP_DISPATCH(gp_fwd_red, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, reduce, st); }
    case DVA_BF16: { using T = __nv_bfloat16; GP_DISPATCH(gp_fwd_red, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, reduce, st); }
    case DVA_F16: { using T = __half; GP_DISPATCH(gp_fwd_red, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, reduce, st); }
    default: return fail(DVA_EINVAL, "gather_pool_fwd: unknown dtype");
  }
}

extern "C" int dva_gather_pool_bwd(const void* grad_out, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   const int64_t* arg, float* grad_fmap, int64_t B, int64_t C,
                                   int64_t H, int64_t W, int64_t Vw, int64_t P, int reduce,
                                   int dtype, void* stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || Vw < 0 || P < 0) return fail(DVA_EINVAL, "gather_pool_bwd: negative size");
  if (Vw == 0 || C == 0 || P == 0) return DVA_OK;
  if (!aptr || !grad_out || !img || !pix || !grad_fmap) return fail(DVA_EINVAL, "gather_pool_bwd: null pointer");
  if ((reduce == DVA_MAX || reduce == DVA_MIN) && !arg) return fail(DVA_EINVAL, "gather_pool_bwd: max/min need arg");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: { using T = float; GP_DISPATCH(gp_bwd_red, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, reduce, st); }
    case DVA_BF16: { using T = __nv_bfloat16; GP_DISPATCH(gp_bwd_red, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, reduce, st); }
    case DVA_F16: { using T = __half; GP_DISPATCH(gp_bwd_red, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, reduce, st); }
    default: return fail(DVA_EINVAL, "gather_pool_bwd: unknown dtype");
  }
}
