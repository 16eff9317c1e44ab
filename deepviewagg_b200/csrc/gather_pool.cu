// Fused feature-map pixel gather + atomic (pixel -> view) pool.
//   reference: x = self.x[(img_id per pixel, ..., py, px)]  (image.py:1285, 1871-1885) makes a
//   [P,C] copy out of the NCHW map, then BimodalCSRPool reduces it over the atomic CSR
//   (modules.py:497-500 -> pooling.py:63).  Here the [P,C] intermediate never exists: one thread
//   owns one (view, channel) output, walks the view's pixels and reads the map directly.
// With a channels-last map ([B,H,W,C]) the C channels of a pixel are one contiguous run, so a
// warp reads 32 consecutive channels per pixel (coalesced); with the reference's NCHW layout
// every element is H*W apart (32-byte sector per 4-byte element) -- supported for drop-in use,
// channels-last is the fast path.
//
// INTERP: the `interpolate=True` branch of get_mapped_features (image.py:1278-1283 ->
// sparse_interpolation, image.py:105-170): pixels live at the mapping resolution (map_w, map_h),
// the feature map is smaller, and every pixel reads 4 bilinear corners of the replicate-padded
// map.  The fp32 arithmetic follows the reference operation by operation (no FMA contraction),
// so corner choices, values and therefore max / argmax decisions are identical in fp32.
#include "dva_common.cuh"

namespace dva {

// Bilinear footprint of one mapping pixel (image.py:145-163).
struct Bilin {
  int r0, r1, c0, c1;          // clamped source rows / columns (replicate padding, image.py:133)
  float w00, w01, w10, w11;    // tl, tr, bl, br
};
__device__ __forceinline__ Bilin bilin_setup(int px, int py, int H, int W, float mw1, float mh1) {
  // coords = pixels / (resolution - 1), (x,y) -> (row, col)       image.py:1280-1281
  const float cy = __fdiv_rn((float)py, mh1), cx = __fdiv_rn((float)px, mw1);
  // pixels = coords * (h, w) + 0.5 in the padded frame               image.py:143
  const float p0 = __fadd_rn(__fmul_rn(cy, (float)H), 0.5f);
  const float p1 = __fadd_rn(__fmul_rn(cx, (float)W), 0.5f);
  const float top = floorf(p0), bottom = floorf(__fadd_rn(p0, 1.f));
  const float left = floorf(p1), right = floorf(__fadd_rn(p1, 1.f));
  const float dyt = __fsub_rn(p0, bottom), dyb = __fsub_rn(p0, top);   // weight of top / bottom row
  const float dxl = __fsub_rn(p1, right), dxr = __fsub_rn(p1, left);
  Bilin b;
  b.w00 = fabsf(__fmul_rn(dyt, dxl)); b.w01 = fabsf(__fmul_rn(dyt, dxr));
  b.w10 = fabsf(__fmul_rn(dyb, dxl)); b.w11 = fabsf(__fmul_rn(dyb, dxr));
  b.r0 = min(max((int)top - 1, 0), H - 1); b.r1 = min(max((int)bottom - 1, 0), H - 1);
  b.c0 = min(max((int)left - 1, 0), W - 1); b.c1 = min(max((int)right - 1, 0), W - 1);
  return b;
}

template <bool CL>
__device__ __forceinline__ int64_t fmap_off(int64_t b, int64_t c, int64_t y, int64_t x, int64_t C,
                                            int64_t H, int64_t W) {
  return CL ? (((b * H + y) * W + x) * C + c) : (((b * C + c) * H + y) * W + x);
}

// Memory safety (the reference's x[feature_map_indexing] raises IndexError on a stale or mis-scaled
// mapping; a kernel cannot raise): pixel coordinates and image ids are clamped into the map, so a bad
// mapping can never read or -- in backward -- atomically write outside the feature-map tensor.
// ops.gather_pool(check_indices=True) / DVA_CHECK_INDICES=1 validates them up front and raises.
__device__ __forceinline__ int clamp_px(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }
__device__ __forceinline__ int64_t clamp_img(int64_t b, int64_t B) { return b < 0 ? 0 : (b >= B ? B - 1 : b); }

template <typename T, typename PIX, bool CL, int RED, bool INTERP>
__global__ void __launch_bounds__(256)
gather_pool_fwd_kernel(const T* __restrict__ fmap, const int64_t* __restrict__ img,
                       const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                       T* __restrict__ out, int64_t* __restrict__ arg, int64_t C, int64_t H,
                       int64_t W, int64_t Vw, int64_t P, float mw1, float mh1, int64_t B) {
  const int64_t total = Vw * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = t / C, c = t - w * C;
    const int64_t p0 = aptr[w], p1 = aptr[w + 1];
    const int64_t b = clamp_img(img[w], B);
    float acc = 0.f;
    int64_t best = P;
    for (int64_t p = p0; p < p1; ++p) {
      int64_t px = (int64_t)pix[2 * p], py = (int64_t)pix[2 * p + 1];
      if (!INTERP) { px = clamp_px((int)px, (int)W); py = clamp_px((int)py, (int)H); }
      float v;
      if (INTERP) {
        const Bilin q = bilin_setup((int)px, (int)py, (int)H, (int)W, mw1, mh1);
        const float f00 = Cvt<T>::to_f(fmap[fmap_off<CL>(b, c, q.r0, q.c0, C, H, W)]);
        const float f01 = Cvt<T>::to_f(fmap[fmap_off<CL>(b, c, q.r0, q.c1, C, H, W)]);
        const float f10 = Cvt<T>::to_f(fmap[fmap_off<CL>(b, c, q.r1, q.c0, C, H, W)]);
        const float f11 = Cvt<T>::to_f(fmap[fmap_off<CL>(b, c, q.r1, q.c1, C, H, W)]);
        // image.py:165-168: four products summed left to right
        v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.w00, f00), __fmul_rn(q.w01, f01)),
                                __fmul_rn(q.w10, f10)), __fmul_rn(q.w11, f11));
      } else {
        v = Cvt<T>::to_f(fmap[fmap_off<CL>(b, c, py, px, C, H, W)]);
      }
      if (RED == DVA_SUM || RED == DVA_MEAN) acc += v;
      else if (p == p0 || (RED == DVA_MAX ? v > acc : v < acc)) { acc = v; best = p; }
    }
    if (RED == DVA_MEAN) acc /= (float)((p1 - p0) > 0 ? (p1 - p0) : 1);
    out[t] = Cvt<T>::from_f(acc);
    // the arg table is only needed (and only written) for views with two or more pixels; a
    // one-pixel view (every view under exact splatting) routes its gradient to that pixel
    if ((RED == DVA_MAX || RED == DVA_MIN) && arg != nullptr && p1 - p0 >= 2) arg[t] = best;
  }
}

template <bool CL, bool INTERP, typename PIX>
__device__ __forceinline__ void scatter_pixel(float* __restrict__ gfmap, const PIX* __restrict__ pix,
                                              int64_t p, int64_t b, int64_t c, int64_t C, int64_t H,
                                              int64_t W, float mw1, float mh1, float g) {
  int64_t px = (int64_t)pix[2 * p], py = (int64_t)pix[2 * p + 1];
  if (!INTERP) { px = clamp_px((int)px, (int)W); py = clamp_px((int)py, (int)H); }
  if (INTERP) {
    const Bilin q = bilin_setup((int)px, (int)py, (int)H, (int)W, mw1, mh1);
    atomicAdd(gfmap + fmap_off<CL>(b, c, q.r0, q.c0, C, H, W), q.w00 * g);
    atomicAdd(gfmap + fmap_off<CL>(b, c, q.r0, q.c1, C, H, W), q.w01 * g);
    atomicAdd(gfmap + fmap_off<CL>(b, c, q.r1, q.c0, C, H, W), q.w10 * g);
    atomicAdd(gfmap + fmap_off<CL>(b, c, q.r1, q.c1, C, H, W), q.w11 * g);
  } else {
    atomicAdd(gfmap + fmap_off<CL>(b, c, py, px, C, H, W), g);
  }
}

template <typename T, typename PIX, bool CL, int RED, bool INTERP>
__global__ void __launch_bounds__(256)
gather_pool_bwd_kernel(const T* __restrict__ gout, const int64_t* __restrict__ img,
                       const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                       const int64_t* __restrict__ arg, float* __restrict__ gfmap, int64_t C,
                       int64_t H, int64_t W, int64_t Vw, float mw1, float mh1, int64_t B) {
  const int64_t total = Vw * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = t / C, c = t - w * C;
    const int64_t p0 = aptr[w], p1 = aptr[w + 1];
    if (p1 <= p0) continue;
    const int64_t b = clamp_img(img[w], B);
    float g = Cvt<T>::to_f(gout[t]);
    if (RED == DVA_MEAN) g /= (float)(p1 - p0);
    if (RED == DVA_MAX || RED == DVA_MIN) {
      scatter_pixel<CL, INTERP>(gfmap, pix, (p1 - p0 == 1) ? p0 : arg[t], b, c, C, H, W, mw1, mh1, g);
    } else {
      for (int64_t p = p0; p < p1; ++p) scatter_pixel<CL, INTERP>(gfmap, pix, p, b, c, C, H, W, mw1, mh1, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// channels-last vector path: LPR lanes own the 16-byte chunks of one view's output row, a warp
// works on 32/LPR views per step and kGpUnroll steps at once (their index loads, pixel loads and
// row-chunk loads are issued back to back), so a pixel costs one LDG.128 per lane instead of
// VEC scalar loads plus 64-bit index arithmetic per channel.  Same arithmetic as the scalar kernel
// (the bilinear branch keeps the reference's fp32 operation order).
// ---------------------------------------------------------------------------------------------
constexpr int kGpWarps = 8;

template <typename T, typename PIX, bool INTERP> struct PixLoad {
  static constexpr int VEC = Vec16<T>::N;
  uint4 raw[INTERP ? 4 : 1];
  Bilin q;
  // fb: map of the view's image + this lane's chunk offset (bytes); pixel_bytes = C * sizeof(T)
  __device__ __forceinline__ void issue(const char* __restrict__ fb, const PIX* __restrict__ pix, int64_t p,
                                        int H, int W, uint32_t pixel_bytes, float mw1, float mh1) {
    int px = (int)pix[2 * p], py = (int)pix[2 * p + 1];
    if constexpr (!INTERP) { px = clamp_px(px, W); py = clamp_px(py, H); }
    if constexpr (INTERP) {
      q = bilin_setup(px, py, H, W, mw1, mh1);
      raw[0] = ldg_stream16(fb + ((int64_t)q.r0 * W + q.c0) * pixel_bytes);
      raw[1] = ldg_stream16(fb + ((int64_t)q.r0 * W + q.c1) * pixel_bytes);
      raw[2] = ldg_stream16(fb + ((int64_t)q.r1 * W + q.c0) * pixel_bytes);
      raw[3] = ldg_stream16(fb + ((int64_t)q.r1 * W + q.c1) * pixel_bytes);
    } else {
      raw[0] = ldg_stream16(fb + ((int64_t)py * W + px) * pixel_bytes);
    }
  }
  __device__ __forceinline__ void value(float (&v)[VEC]) const {
    if constexpr (INTERP) {
      float f00[VEC], f01[VEC], f10[VEC], f11[VEC];
      unpack16<T, VEC>(raw[0], f00); unpack16<T, VEC>(raw[1], f01);
      unpack16<T, VEC>(raw[2], f10); unpack16<T, VEC>(raw[3], f11);
#pragma unroll
      for (int j = 0; j < VEC; ++j)   // image.py:165-168: four products summed left to right
        v[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.w00, f00[j]), __fmul_rn(q.w01, f01[j])),
                                   __fmul_rn(q.w10, f10[j])), __fmul_rn(q.w11, f11[j]));
    } else {
      unpack16<T, VEC>(raw[0], v);
    }
  }
};

template <typename T, typename PIX, int LPR, int RED, bool INTERP>
__global__ void __launch_bounds__(kGpWarps * 32)
gather_pool_fwd_cl_kernel(const T* __restrict__ fmap, const int64_t* __restrict__ img,
                          const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                          T* __restrict__ out, int64_t* __restrict__ arg, int C, int H, int W,
                          int64_t Vw, int64_t P, float mw1, float mh1, int64_t B) {
  constexpr int VEC = Vec16<T>::N, RPI = 32 / LPR, U = INTERP ? 2 : 4;
  const int lane = threadIdx.x & 31, sg = lane / LPR, lir = lane % LPR;
  const int cv = C / VEC, tiles = (cv + LPR - 1) / LPR;
  const int64_t items = Vw * tiles;
  const uint32_t pixel_bytes = (uint32_t)C * sizeof(T);
  const int64_t map_bytes = (int64_t)H * W * pixel_bytes;
  const char* __restrict__ fbase = reinterpret_cast<const char*>(fmap);
  const int64_t gwarp = (int64_t)blockIdx.x * kGpWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kGpWarps * RPI * U;
  for (int64_t it0 = gwarp * RPI * U; it0 < items; it0 += stride) {
    int64_t w[U], p0[U]; int n[U], ck[U]; bool act[U];
    const char* fb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t item = it0 + u * RPI + sg;
      act[u] = item < items;
      w[u] = act[u] ? (tiles == 1 ? item : item / tiles) : 0;
      ck[u] = (int)(item - w[u] * tiles) * LPR + lir;
      act[u] = act[u] && ck[u] < cv;
      p0[u] = aptr[w[u]];
      n[u] = (int)(aptr[w[u] + 1] - p0[u]);
      fb[u] = fbase + clamp_img(img[w[u]], B) * map_bytes + (act[u] ? ck[u] * 16 : 0);
    }
    PixLoad<T, PIX, INTERP> pl[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (n[u] > 0) pl[u].issue(fb[u], pix, p0[u], H, W, pixel_bytes, mw1, mh1);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc[VEC]; int best[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) { acc[j] = 0.f; best[j] = 0; }
      if (n[u] > 0) pl[u].value(acc);
      for (int k = 1; k < n[u]; ++k) {              // two or more pixels per view: not under exact splatting
        PixLoad<T, PIX, INTERP> nx;
        nx.issue(fb[u], pix, p0[u] + k, H, W, pixel_bytes, mw1, mh1);
        float v[VEC];
        nx.value(v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if (RED == DVA_SUM || RED == DVA_MEAN) acc[j] += v[j];
          else if (RED == DVA_MAX ? v[j] > acc[j] : v[j] < acc[j]) { acc[j] = v[j]; best[j] = k; }
        }
      }
      if (RED == DVA_MEAN && n[u] > 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] /= (float)n[u];
      }
      if (act[u]) {
        const int64_t e0 = w[u] * C + (int64_t)ck[u] * VEC;
        stg_stream16(reinterpret_cast<char*>(out) + e0 * sizeof(T), pack16<T, VEC>(acc));
        if ((RED == DVA_MAX || RED == DVA_MIN) && arg != nullptr && n[u] >= 2) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) arg[e0 + j] = p0[u] + best[j];
        }
      }
    }
  }
}

// 16-byte vector reduction into global memory (sm_90+): one instruction per four channels
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename PIX, bool INTERP, int VEC>
__device__ __forceinline__ void scatter_chunk(float* __restrict__ gmap_b /* image + chunk offset */,
                                              const PIX* __restrict__ pix, int64_t p, int C, int H, int W,
                                              float mw1, float mh1, const float (&g)[VEC]) {
  int px = (int)pix[2 * p], py = (int)pix[2 * p + 1];
  if constexpr (!INTERP) { px = clamp_px(px, W); py = clamp_px(py, H); }
  if constexpr (INTERP) {
    const Bilin q = bilin_setup(px, py, H, W, mw1, mh1);
    const int64_t o[4] = {((int64_t)q.r0 * W + q.c0) * C, ((int64_t)q.r0 * W + q.c1) * C,
                          ((int64_t)q.r1 * W + q.c0) * C, ((int64_t)q.r1 * W + q.c1) * C};
    const float wq[4] = {q.w00, q.w01, q.w10, q.w11};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < VEC; j += 4)
        red_add_v4(gmap_b + o[k] + j, wq[k] * g[j], wq[k] * g[j + 1], wq[k] * g[j + 2], wq[k] * g[j + 3]);
  } else {
    float* a = gmap_b + ((int64_t)py * W + px) * C;
#pragma unroll
    for (int j = 0; j < VEC; j += 4) red_add_v4(a + j, g[j], g[j + 1], g[j + 2], g[j + 3]);
  }
}

template <typename T, typename PIX, int LPR, int RED, bool INTERP>
__global__ void __launch_bounds__(kGpWarps * 32)
gather_pool_bwd_cl_kernel(const T* __restrict__ gout, const int64_t* __restrict__ img,
                          const PIX* __restrict__ pix, const int64_t* __restrict__ aptr,
                          const int64_t* __restrict__ arg, float* __restrict__ gfmap, int C, int H,
                          int W, int64_t Vw, float mw1, float mh1, int64_t B) {
  constexpr int VEC = Vec16<T>::N, RPI = 32 / LPR, U = 4;
  const int lane = threadIdx.x & 31, sg = lane / LPR, lir = lane % LPR;
  const int cv = C / VEC, tiles = (cv + LPR - 1) / LPR;
  const int64_t items = Vw * tiles;
  const int64_t map_elems = (int64_t)H * W * C;
  const int64_t gwarp = (int64_t)blockIdx.x * kGpWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kGpWarps * RPI * U;
  for (int64_t it0 = gwarp * RPI * U; it0 < items; it0 += stride) {
    int64_t w[U], p0[U], b[U]; int n[U], ck[U]; bool act[U];
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t item = it0 + u * RPI + sg;
      act[u] = item < items;
      w[u] = act[u] ? (tiles == 1 ? item : item / tiles) : 0;
      ck[u] = (int)(item - w[u] * tiles) * LPR + lir;
      act[u] = act[u] && ck[u] < cv;
      p0[u] = aptr[w[u]];
      n[u] = (int)(aptr[w[u] + 1] - p0[u]);
      b[u] = clamp_img(img[w[u]], B);
      act[u] = act[u] && n[u] > 0;
      if (act[u]) raw[u] = ldg_stream16(reinterpret_cast<const char*>(gout) + (w[u] * C + (int64_t)ck[u] * VEC) * sizeof(T));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      float g[VEC];
      unpack16<T, VEC>(raw[u], g);
      if (RED == DVA_MEAN) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] /= (float)n[u];
      }
      const int c0 = ck[u] * VEC;
      float* gm = gfmap + b[u] * map_elems + c0;
      if (RED == DVA_MAX || RED == DVA_MIN) {
        if (n[u] == 1) {
          scatter_chunk<PIX, INTERP, VEC>(gm, pix, p0[u], C, H, W, mw1, mh1, g);
        } else {                                    // per-channel winners: scalar atomics
          const int64_t e0 = w[u] * C + c0;
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            scatter_pixel<true, INTERP>(gfmap, pix, arg[e0 + j], b[u], c0 + j, C, H, W, mw1, mh1, g[j]);
        }
      } else {
        for (int k = 0; k < n[u]; ++k) scatter_chunk<PIX, INTERP, VEC>(gm, pix, p0[u] + k, C, H, W, mw1, mh1, g);
      }
    }
  }
}

template <typename T> static int gp_cl_lpr(int64_t C) {
  const int64_t cv = C / Vec16<T>::N;
  return cv <= 4 ? 4 : (cv <= 8 ? 8 : (cv <= 16 ? 16 : 32));
}
template <typename T> static bool gp_cl_vec_ok(const void* map, const void* rows, int64_t C, int64_t H, int64_t W) {
  return C % Vec16<T>::N == 0 && aligned16(map) && aligned16(rows) && C < (1 << 20) && H * W < (1ll << 31);
}
static inline int gp_cl_grid(int64_t items, int rpi, int unroll) {
  int64_t blocks = (items + (int64_t)kGpWarps * rpi * unroll - 1) / ((int64_t)kGpWarps * rpi * unroll);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

static inline int gp_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

template <typename T, typename PIX, bool CL, bool INTERP>
static int gp_fwd_red(const void* fmap, const int64_t* img, const void* pix, const int64_t* aptr,
                      void* out, int64_t* arg, int64_t C, int64_t H, int64_t W, int64_t Vw,
                      int64_t P, float mw1, float mh1, int reduce, cudaStream_t st, int64_t B) {
  if constexpr (CL) {
    if (gp_cl_vec_ok<T>(fmap, out, C, H, W)) {
      const int lpr = gp_cl_lpr<T>(C);
      const int64_t cvv = C / Vec16<T>::N;
      const int64_t items = Vw * ((cvv + lpr - 1) / lpr);
      const int gridv = gp_cl_grid(items, 32 / lpr, INTERP ? 2 : 4);
#define GP_FV(R, L) gather_pool_fwd_cl_kernel<T, PIX, L, R, INTERP><<<gridv, kGpWarps * 32, 0, st>>>((const T*)fmap, img, (const PIX*)pix, aptr, (T*)out, arg, (int)C, (int)H, (int)W, Vw, P, mw1, mh1, B)
#define GP_FVL(R) do { if (lpr == 4) GP_FV(R, 4); else if (lpr == 8) GP_FV(R, 8); else if (lpr == 16) GP_FV(R, 16); else GP_FV(R, 32); } while (0)
      switch (reduce) {
        case DVA_SUM: GP_FVL(DVA_SUM); break;
        case DVA_MEAN: GP_FVL(DVA_MEAN); break;
        case DVA_MAX: GP_FVL(DVA_MAX); break;
        case DVA_MIN: GP_FVL(DVA_MIN); break;
        default: return fail(DVA_EINVAL, "gather_pool_fwd: unknown reduce");
      }
#undef GP_FVL
#undef GP_FV
      return check_launch("gather_pool_fwd(cl)");
    }
  }
  const int grid = gp_grid(Vw * C);
#define GP_F(R) gather_pool_fwd_kernel<T, PIX, CL, R, INTERP><<<grid, 256, 0, st>>>((const T*)fmap, img, (const PIX*)pix, aptr, (T*)out, arg, C, H, W, Vw, P, mw1, mh1, B)
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

template <typename T, typename PIX, bool CL, bool INTERP>
static int gp_bwd_red(const void* gout, const int64_t* img, const void* pix, const int64_t* aptr,
                      const int64_t* arg, float* gfmap, int64_t C, int64_t H, int64_t W,
                      int64_t Vw, float mw1, float mh1, int reduce, cudaStream_t st, int64_t B) {
  if constexpr (CL) {
    if (gp_cl_vec_ok<T>(gfmap, gout, C, H, W) && C % 4 == 0) {
      const int lpr = gp_cl_lpr<T>(C);
      const int64_t cvv = C / Vec16<T>::N;
      const int64_t items = Vw * ((cvv + lpr - 1) / lpr);
      const int gridv = gp_cl_grid(items, 32 / lpr, 4);
#define GP_BV(R, L) gather_pool_bwd_cl_kernel<T, PIX, L, R, INTERP><<<gridv, kGpWarps * 32, 0, st>>>((const T*)gout, img, (const PIX*)pix, aptr, arg, gfmap, (int)C, (int)H, (int)W, Vw, mw1, mh1, B)
#define GP_BVL(R) do { if (lpr == 4) GP_BV(R, 4); else if (lpr == 8) GP_BV(R, 8); else if (lpr == 16) GP_BV(R, 16); else GP_BV(R, 32); } while (0)
      switch (reduce) {
        case DVA_SUM: GP_BVL(DVA_SUM); break;
        case DVA_MEAN: GP_BVL(DVA_MEAN); break;
        case DVA_MAX: GP_BVL(DVA_MAX); break;
        case DVA_MIN: GP_BVL(DVA_MIN); break;
        default: return fail(DVA_EINVAL, "gather_pool_bwd: unknown reduce");
      }
#undef GP_BVL
#undef GP_BV
      return check_launch("gather_pool_bwd(cl)");
    }
  }
  const int grid = gp_grid(Vw * C);
#define GP_B(R) gather_pool_bwd_kernel<T, PIX, CL, R, INTERP><<<grid, 256, 0, st>>>((const T*)gout, img, (const PIX*)pix, aptr, arg, gfmap, C, H, W, Vw, mw1, mh1, B)
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

#define GP_DISPATCH(FN, INTERP, ...)                                                      \
  do {                                                                                    \
    if (channels_last) {                                                                  \
      if (pix_is_i16) return FN<T, int16_t, true, INTERP>(__VA_ARGS__);                   \
      return FN<T, int32_t, true, INTERP>(__VA_ARGS__);                                   \
    }                                                                                     \
    if (pix_is_i16) return FN<T, int16_t, false, INTERP>(__VA_ARGS__);                    \
    return FN<T, int32_t, false, INTERP>(__VA_ARGS__);                                    \
  } while (0)

template <bool INTERP>
static int gather_pool_fwd_impl(const char* who, const void* fmap, int channels_last, const int64_t* img,
                                const void* pix, int pix_is_i16, const int64_t* aptr, void* out,
                                int64_t* arg, int64_t B, int64_t C, int64_t H, int64_t W,
                                int64_t map_w, int64_t map_h, int64_t Vw, int64_t P, int reduce,
                                int dtype, void* stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || Vw < 0 || P < 0) return failf(DVA_EINVAL, "%s: negative size", who);
  if (Vw == 0 || C == 0) return DVA_OK;
  if (P > 0 && (B < 1 || H < 1 || W < 1)) return failf(DVA_EINVAL, "%s: pixels given but the map is empty", who);
  if (!aptr || !out || !img || (P > 0 && (!fmap || !pix))) return failf(DVA_EINVAL, "%s: null pointer", who);
  if (INTERP && (map_w < 2 || map_h < 2 || H < 1 || W < 1 || H > (1 << 24) || W > (1 << 24)))
    return failf(DVA_EINVAL, "%s: bad map / mapping size", who);
  const float mw1 = (float)(map_w - 1), mh1 = (float)(map_h - 1);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: { using T = float; GP_DISPATCH(gp_fwd_red, INTERP, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, mw1, mh1, reduce, st, B); }
    case DVA_BF16: { using T = __nv_bfloat16; GP_DISPATCH(gp_fwd_red, INTERP, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, mw1, mh1, reduce, st, B); }
    case DVA_F16: { using T = __half; GP_DISPATCH(gp_fwd_red, INTERP, fmap, img, pix, aptr, out, arg, C, H, W, Vw, P, mw1, mh1, reduce, st, B); }
    default: return failf(DVA_EINVAL, "%s: unknown dtype", who);
  }
}

template <bool INTERP>
static int gather_pool_bwd_impl(const char* who, const void* grad_out, int channels_last,
                                const int64_t* img, const void* pix, int pix_is_i16,
                                const int64_t* aptr, const int64_t* arg, float* grad_fmap, int64_t B,
                                int64_t C, int64_t H, int64_t W, int64_t map_w, int64_t map_h,
                                int64_t Vw, int64_t P, int reduce, int dtype, void* stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || Vw < 0 || P < 0) return failf(DVA_EINVAL, "%s: negative size", who);
  if (Vw == 0 || C == 0 || P == 0) return DVA_OK;
  if (B < 1 || H < 1 || W < 1) return failf(DVA_EINVAL, "%s: pixels given but the map is empty", who);
  if (!aptr || !grad_out || !img || !pix || !grad_fmap) return failf(DVA_EINVAL, "%s: null pointer", who);
  if ((reduce == DVA_MAX || reduce == DVA_MIN) && !arg) return failf(DVA_EINVAL, "%s: max/min need arg", who);
  if (INTERP && (map_w < 2 || map_h < 2 || H < 1 || W < 1 || H > (1 << 24) || W > (1 << 24)))
    return failf(DVA_EINVAL, "%s: bad map / mapping size", who);
  const float mw1 = (float)(map_w - 1), mh1 = (float)(map_h - 1);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: { using T = float; GP_DISPATCH(gp_bwd_red, INTERP, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, mw1, mh1, reduce, st, B); }
    case DVA_BF16: { using T = __nv_bfloat16; GP_DISPATCH(gp_bwd_red, INTERP, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, mw1, mh1, reduce, st, B); }
    case DVA_F16: { using T = __half; GP_DISPATCH(gp_bwd_red, INTERP, grad_out, img, pix, aptr, arg, grad_fmap, C, H, W, Vw, mw1, mh1, reduce, st, B); }
    default: return failf(DVA_EINVAL, "%s: unknown dtype", who);
  }
}

extern "C" int dva_gather_pool_fwd(const void* fmap, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   void* out, int64_t* arg, int64_t B, int64_t C, int64_t H,
                                   int64_t W, int64_t Vw, int64_t P, int reduce, int dtype,
                                   void* stream) {
  return gather_pool_fwd_impl<false>("gather_pool_fwd", fmap, channels_last, img, pix, pix_is_i16, aptr,
                                     out, arg, B, C, H, W, 0, 0, Vw, P, reduce, dtype, stream);
}

extern "C" int dva_gather_pool_bwd(const void* grad_out, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   const int64_t* arg, float* grad_fmap, int64_t B, int64_t C,
                                   int64_t H, int64_t W, int64_t Vw, int64_t P, int reduce,
                                   int dtype, void* stream) {
  return gather_pool_bwd_impl<false>("gather_pool_bwd", grad_out, channels_last, img, pix, pix_is_i16,
                                     aptr, arg, grad_fmap, B, C, H, W, 0, 0, Vw, P, reduce, dtype, stream);
}

extern "C" int dva_interp_pool_fwd(const void* fmap, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   void* out, int64_t* arg, int64_t B, int64_t C, int64_t H,
                                   int64_t W, int64_t map_w, int64_t map_h, int64_t Vw, int64_t P,
                                   int reduce, int dtype, void* stream) {
  return gather_pool_fwd_impl<true>("interp_pool_fwd", fmap, channels_last, img, pix, pix_is_i16, aptr,
                                    out, arg, B, C, H, W, map_w, map_h, Vw, P, reduce, dtype, stream);
}

extern "C" int dva_interp_pool_bwd(const void* grad_out, int channels_last, const int64_t* img,
                                   const void* pix, int pix_is_i16, const int64_t* aptr,
                                   const int64_t* arg, float* grad_fmap, int64_t B, int64_t C,
                                   int64_t H, int64_t W, int64_t map_w, int64_t map_h, int64_t Vw,
                                   int64_t P, int reduce, int dtype, void* stream) {
  return gather_pool_bwd_impl<true>("interp_pool_bwd", grad_out, channels_last, img, pix, pix_is_i16,
                                    aptr, arg, grad_fmap, B, C, H, W, map_w, map_h, Vw, P, reduce, dtype, stream);
}

// ---------------------------------------------------------------------------------------------
// [B, R, S] -> [B, S, R] (NCHW <-> NHWC with R = C, S = H*W or the reverse): lets the reference's
// NCHW-contiguous feature maps use the channels-last gather / scatter kernels when a large share of
// the map is gathered.  32 x 32 shared tiles, both sides coalesced.
// ---------------------------------------------------------------------------------------------
namespace dva {
template <typename T>
__global__ void __launch_bounds__(256)
transpose_last2_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t S) {
  __shared__ T tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads
  const T* sb = src + b * R * S;
  T* db = dst + b * R * S;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t r = r0 + ty + j, s = s0 + tx;
    if (r < R && s < S) tile[ty + j][tx] = sb[r * S + s];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t s = s0 + ty + j, r = r0 + tx;
    if (r < R && s < S) db[s * R + r] = tile[tx][ty + j];
  }
}
}  // namespace dva

extern "C" int dva_transpose_last2(const void* src, void* dst, int64_t B, int64_t R, int64_t S, int dtype,
                                   void* stream) {
  if (B < 0 || R < 0 || S < 0) return fail(DVA_EINVAL, "transpose_last2: negative size");
  if (B == 0 || R == 0 || S == 0) return DVA_OK;
  if (!src || !dst) return fail(DVA_EINVAL, "transpose_last2: null pointer");
  if (B > 65535 || (R + 31) / 32 > 65535) return fail(DVA_EUNSUPPORTED, "transpose_last2: batch / row count too large");
  const dim3 grid((unsigned)((S + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: transpose_last2_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, R, S); break;
    case DVA_BF16:
    case DVA_F16: transpose_last2_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, R, S); break;
    default: return fail(DVA_EINVAL, "transpose_last2: unknown dtype");
  }
  return check_launch("transpose_last2");
}
