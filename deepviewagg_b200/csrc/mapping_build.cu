// Native construction of the two-level point -> view -> pixel CSR (ImageMapping) from an unordered list
// of (point, image, pixel) items, and the re-indexing operations built on it:
//   ImageMapping.from_dense               image.py:1728-1795   (MapImages, every dataset build)
//   ImageMapping.select_points('merge')   image.py:2211-2273   (after every strided 3D conv, modules.py:232)
//   ImageData.view_cat_sorting            image.py:1549-1574   (every forward of a multi-setting branch)
// The reference composes these from lexargsort / lexargunique (composite int64 key -> full sort ->
// unique), scatter_mean, repeat_interleave and cumsum, with several host synchronisations (.item(), max).
//
// Here: items are BUCKETED by point -- the point id is a dense integer in [0, num_points), so the top
// level of the CSR is a counting sort (histogram + exclusive scan + scatter), not a comparison sort --
// and each point's handful of items is then ordered by one warp with a rank sort on the key
// (image, [x, y,] source index).  Runs of equal image inside a point are its views; their pixels follow
// in order.  Everything is integer, deterministic (the source index breaks every tie, i.e. the result
// equals a STABLE lexicographic sort) and enqueued on the caller's stream; the only value the host needs
// is the pair (V, P) of output sizes, written to a device word the caller reads once.
//
// HBM-bound integer work: per item ~6 passes of 8..24 bytes; no tensor cores.
#include "dva_common.cuh"

namespace dva {
namespace mb {

constexpr int kScanItems = 2048;          // elements per scan block (256 threads x 8)

// ---- exclusive scan int32 -> int64 (three phases; sizes up to 2^31 blocks of 2048) -------------------
__global__ void __launch_bounds__(256)
scan_block_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ block_sums) {
  __shared__ int64_t red[8];
  const int64_t base = (int64_t)blockIdx.x * kScanItems;
  int64_t s = 0;
  for (int k = 0; k < 8; ++k) {
    const int64_t i = base + k * 256 + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    block_sums[blockIdx.x] = t;
  }
}

// single CTA: exclusive scan of the block sums in place; total -> sums[n_blocks]
__global__ void __launch_bounds__(1024)
scan_of_sums(int64_t* __restrict__ sums, int64_t n_blocks) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_blocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n_blocks ? sums[i] : 0;
    int64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int64_t w = warp_tot[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;                  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int64_t carry = carry_s;
    if (i < n_blocks) sums[i] = carry + warp_tot[threadIdx.x >> 5] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[n_blocks] = carry_s;
}

__global__ void __launch_bounds__(256)
scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ block_offsets,
           int64_t* __restrict__ out /* [n + 1] */) {
  __shared__ int64_t warp_tot[8];
  const int64_t base = (int64_t)blockIdx.x * kScanItems + (int64_t)threadIdx.x * 8;
  int32_t v[8];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
  int64_t inc = s;
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
  __syncthreads();
  int64_t pre = block_offsets[blockIdx.x] + inc - s;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) pre += warp_tot[w];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < n) out[base + k] = pre;
    pre += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_offsets[gridDim.x];
}

static int exclusive_scan(const int32_t* in, int64_t n, int64_t* out, int64_t* block_sums, cudaStream_t st) {
  // out[0..n] = exclusive prefix sums of in[0..n), out[n] = total.  block_sums: ceil(n / 2048) + 1 words
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(out, 0, 8, st);
    return e == cudaSuccess ? DVA_OK : fail((int)e, "scan: memset failed");
  }
  const int64_t nb = (n + kScanItems - 1) / kScanItems;
  scan_block_sums<<<(unsigned)nb, 256, 0, st>>>(in, n, block_sums);
  int rc = check_launch("scan_block_sums");
  if (rc) return rc;
  scan_of_sums<<<1, 1024, 0, st>>>(block_sums, nb);
  if ((rc = check_launch("scan_of_sums"))) return rc;
  scan_apply<<<(unsigned)nb, 256, 0, st>>>(in, n, block_sums, out);
  return check_launch("scan_apply");
}

// ---- bucketing -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_points(const int64_t* __restrict__ point, int64_t n, int64_t num_points, int32_t* __restrict__ cnt,
             int32_t* __restrict__ status) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = point[i];
    if (p < 0 || p >= num_points) { atomicOr(status, 1); continue; }     // reported to the host with the sizes
    atomicAdd(cnt + p, 1);
  }
}

__global__ void __launch_bounds__(256)
scatter_items(const int64_t* __restrict__ point, int64_t n, int64_t num_points, const int64_t* __restrict__ off,
              int32_t* __restrict__ cursor, int64_t* __restrict__ bucket) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = point[i];
    if (p < 0 || p >= num_points) continue;
    bucket[off[p] + atomicAdd(cursor + p, 1)] = i;       // arbitrary order inside the bucket; ordered next
  }
}

template <typename PIX> struct PixIO;
template <> struct PixIO<int16_t> { static __device__ __forceinline__ void ld(const void* p, int64_t i, int& x, int& y) { const short2 v = reinterpret_cast<const short2*>(p)[i]; x = v.x; y = v.y; }
                                    static __device__ __forceinline__ void st(void* p, int64_t i, int x, int y) { reinterpret_cast<short2*>(p)[i] = make_short2((short)x, (short)y); } };
template <> struct PixIO<int32_t> { static __device__ __forceinline__ void ld(const void* p, int64_t i, int& x, int& y) { const int2 v = reinterpret_cast<const int2*>(p)[i]; x = v.x; y = v.y; }
                                    static __device__ __forceinline__ void st(void* p, int64_t i, int x, int y) { reinterpret_cast<int2*>(p)[i] = make_int2(x, y); } };
template <> struct PixIO<int64_t> { static __device__ __forceinline__ void ld(const void* p, int64_t i, int& x, int& y) { const longlong2 v = reinterpret_cast<const longlong2*>(p)[i]; x = (int)v.x; y = (int)v.y; }
                                    static __device__ __forceinline__ void st(void* p, int64_t i, int x, int y) { reinterpret_cast<longlong2*>(p)[i] = make_longlong2(x, y); } };

// key of an item inside its point: (image, [x, y]) then the source index (stability)
struct Key { int64_t a; int64_t src; };
__device__ __forceinline__ bool key_less(const Key& u, const Key& v) { return u.a < v.a || (u.a == v.a && u.src < v.src); }

template <typename PIX>
__device__ __forceinline__ Key make_key(const int64_t* image, const void* pix, int64_t src, bool by_pixel) {
  Key k; k.src = src;
  int64_t a = image[src];
  if (by_pixel) { int x, y; PixIO<PIX>::ld(pix, src, x, y); a = (a << 32) | ((int64_t)(x & 0xffff) << 16) | (int64_t)(y & 0xffff); }
  k.a = a;
  return k;
}

// One warp per point: rank-sort the bucket, then flag view heads / duplicate pixels and count them.
// flags[pos]: bit 0 = first item of a view, bit 1 = kept pixel.  (x, y must fit 16 bits when by_pixel.)
template <typename PIX>
__global__ void __launch_bounds__(256)
order_points(const int64_t* __restrict__ image, const void* __restrict__ pix, const int64_t* __restrict__ off,
             int64_t* __restrict__ bucket, int64_t* __restrict__ sorted, uint8_t* __restrict__ flags,
             int32_t* __restrict__ n_views, int32_t* __restrict__ n_pix, int64_t num_points, int dedupe) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < num_points; p += warps) {
    const int64_t b0 = off[p], L = off[p + 1] - b0;
    if (L == 0) { if (lane == 0) { n_views[p] = 0; n_pix[p] = 0; } continue; }
    // ---- rank sort (stable through the source index) ----
    if (L <= 32) {
      Key mine; mine.a = 0; mine.src = 0;
      if (lane < L) mine = make_key<PIX>(image, pix, bucket[b0 + lane], dedupe != 0);
      int rank = 0;
      for (int j = 0; j < (int)L; ++j) {
        Key o; o.a = __shfl_sync(0xffffffffu, mine.a, j); o.src = __shfl_sync(0xffffffffu, mine.src, j);
        rank += key_less(o, mine) ? 1 : 0;
      }
      if (lane < L) sorted[b0 + rank] = mine.src;
    } else {
      for (int64_t i = lane; i < L; i += 32) {
        const Key mine = make_key<PIX>(image, pix, bucket[b0 + i], dedupe != 0);
        int64_t rank = 0;
        for (int64_t j = 0; j < L; ++j) rank += key_less(make_key<PIX>(image, pix, bucket[b0 + j], dedupe != 0), mine) ? 1 : 0;
        sorted[b0 + rank] = mine.src;
      }
    }
    __syncwarp();
    // ---- flags and counts over the ordered bucket ----
    int nv = 0, np = 0;
    for (int64_t c = 0; c < L; c += 32) {
      const int64_t j = c + lane;
      bool head = false, keep = false;
      if (j < L) {
        const int64_t s = sorted[b0 + j];
        const int64_t m = image[s];
        head = true; keep = true;
        if (j > 0) {
          const int64_t sp = sorted[b0 + j - 1];
          head = image[sp] != m;
          if (dedupe && !head) {
            int x, y, xp, yp;
            PixIO<PIX>::ld(pix, s, x, y); PixIO<PIX>::ld(pix, sp, xp, yp);
            keep = !(x == xp && y == yp);
          }
        }
        flags[b0 + j] = (uint8_t)((head ? 1 : 0) | (keep ? 2 : 0));
      }
      nv += __popc(__ballot_sync(0xffffffffu, head));
      np += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) { n_views[p] = nv; n_pix[p] = np; }
  }
}

// One warp per point: write images / atomic pointers / pixels / per-view mean features.
template <typename PIX>
__global__ void __launch_bounds__(256)
emit_points(const int64_t* __restrict__ image, const void* __restrict__ pix, const float* __restrict__ feat,
            const int64_t* __restrict__ feat_row, const uint8_t* __restrict__ feat_on, int F,
            const int64_t* __restrict__ off, const int64_t* __restrict__ sorted, const uint8_t* __restrict__ flags,
            const int64_t* __restrict__ view_ptr, const int64_t* __restrict__ pix_ptr, int64_t num_points,
            int64_t* __restrict__ images_out, int64_t* __restrict__ atomic_ptr, void* __restrict__ pix_out,
            float* __restrict__ feat_out, int64_t* __restrict__ order_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < num_points; p += warps) {
    const int64_t b0 = off[p], L = off[p + 1] - b0;
    int64_t v_base = view_ptr[p], q_base = pix_ptr[p];
    for (int64_t c = 0; c < L; c += 32) {
      const int64_t j = c + lane;
      const uint8_t f = j < L ? flags[b0 + j] : 0;
      const unsigned heads = __ballot_sync(0xffffffffu, f & 1), keeps = __ballot_sync(0xffffffffu, f & 2);
      const unsigned below = (1u << lane) - 1u;
      const int64_t v = v_base + __popc(heads & below) + ((f & 1) ? 0 : -1);     // view of this item
      const int64_t q = q_base + __popc(keeps & below);                          // kept-pixel slot
      if (j < L) {
        const int64_t s = sorted[b0 + j];
        if (f & 2) {
          int x, y;
          PixIO<PIX>::ld(pix, s, x, y);
          PixIO<PIX>::st(pix_out, q, x, y);
          if (order_out) order_out[q] = s;
        }
        if (f & 1) {
          images_out[v] = image[s];
          atomic_ptr[v] = q;
          if (feat != nullptr) {          // mean over the view's counted items, in order (views are short)
            float acc[16];
            const int Fc = F < 16 ? F : 16;
            for (int k = 0; k < Fc; ++k) acc[k] = 0.f;
            int cnt = 0;
            for (int64_t t = j; t < L; ++t) {
              if (t > j && (flags[b0 + t] & 1)) break;
              const int64_t st = sorted[b0 + t];
              if (feat_on == nullptr || feat_on[st]) {
                const float* row = feat + (feat_row ? feat_row[st] : st) * (int64_t)F;
                for (int k = 0; k < Fc; ++k) acc[k] += row[k];
                ++cnt;
              }
            }
            const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
            for (int k = 0; k < Fc; ++k) feat_out[v * (int64_t)F + k] = acc[k] * inv;
          }
        }
      }
      v_base += __popc(heads);
      q_base += __popc(keeps);
    }
  }
}

__global__ void finish_counts(const int64_t* __restrict__ view_ptr, const int64_t* __restrict__ pix_ptr,
                              int64_t num_points, int64_t* __restrict__ atomic_ptr, int64_t* __restrict__ counts,
                              const int32_t* __restrict__ status) {
  const int64_t V = view_ptr[num_points], P = pix_ptr[num_points];
  atomic_ptr[V] = P;
  counts[0] = V; counts[1] = P; counts[2] = *status;
}

// ---- view_cat_sorting (image.py:1549-1574) in closed form ------------------------------------------------
// Settings s = 0..S-1 each hold a view CSR over the same N points.  The permutation that interleaves the
// concatenated views into point order (a stable argsort of the dense point ids) needs no sort:
//   dest(s, p, j) = sum_s' ptr_s'[p] + sum_{s' < s} (ptr_s'[p+1] - ptr_s'[p]) + (j - ptr_s[p]);  sorting[dest] = base_s + j
__global__ void __launch_bounds__(256)
view_cat_sorting_kernel(const int64_t* const* __restrict__ ptrs, const int64_t* __restrict__ bases, int S,
                        int64_t N, int64_t* __restrict__ sorting, int64_t* __restrict__ csr_cat) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= N; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t start = 0;
    for (int s = 0; s < S; ++s) start += ptrs[s][p];
    csr_cat[p] = start;
    if (p == N) continue;
    int64_t d = start;
    for (int s = 0; s < S; ++s) {
      const int64_t a = ptrs[s][p], b = ptrs[s][p + 1];
      for (int64_t j = a; j < b; ++j) sorting[d++] = bases[s] + j;
    }
  }
}

static inline int grid_for(int64_t total, int per_block = 256) {
  int64_t blocks = (total + per_block - 1) / per_block;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

struct Workspace {
  int32_t *cnt, *cursor, *n_views, *n_pix, *status;
  int64_t *off, *pix_ptr, *bucket, *sorted, *block_sums;
  uint8_t* flags;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t carve(uint8_t* base, int64_t n, int64_t N, Workspace* w) {
  size_t o = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + o : nullptr; o += align256(bytes); return p; };
  const int64_t nb = (N + kScanItems - 1) / kScanItems + 2;
  uint8_t* p;
  p = take((size_t)(N + 1) * 4); if (w) w->cnt = (int32_t*)p;
  p = take((size_t)(N + 1) * 4); if (w) w->cursor = (int32_t*)p;
  p = take((size_t)(N + 1) * 4); if (w) w->n_views = (int32_t*)p;
  p = take((size_t)(N + 1) * 4); if (w) w->n_pix = (int32_t*)p;
  p = take(256); if (w) w->status = (int32_t*)p;
  p = take((size_t)(N + 1) * 8); if (w) w->off = (int64_t*)p;
  p = take((size_t)(N + 1) * 8); if (w) w->pix_ptr = (int64_t*)p;
  p = take((size_t)(n + 1) * 8); if (w) w->bucket = (int64_t*)p;
  p = take((size_t)(n + 1) * 8); if (w) w->sorted = (int64_t*)p;
  p = take((size_t)nb * 8); if (w) w->block_sums = (int64_t*)p;
  p = take((size_t)(n + 1)); if (w) w->flags = p;
  return o;
}

}  // namespace mb
}  // namespace dva

using namespace dva;

extern "C" size_t dva_mapping_build_workspace_bytes(int64_t n_items, int64_t num_points) {
  if (n_items < 0 || num_points < 0) return 0;
  return mb::carve(nullptr, n_items, num_points, nullptr) + 256;
}

// See include/dva_b200.h.  pix_code: 0 = int16, 1 = int32, 2 = int64 pairs (x, y).
extern "C" int dva_mapping_build(const int64_t* point_ids, const int64_t* image_ids, const void* pixels, int pix_code,
                                 const float* feat, const int64_t* feat_row, const uint8_t* feat_on, int64_t F,
                                 int64_t n_items, int64_t num_points, int dedupe_pixels, int64_t* view_ptr,
                                 int64_t* images_out, int64_t* atomic_ptr, void* pixels_out, float* feat_out,
                                 int64_t* order_out, int64_t* counts, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (n_items < 0 || num_points < 0 || F < 0 || F > 16 || pix_code < 0 || pix_code > 2)
    return fail(DVA_EINVAL, "mapping_build: bad sizes (F <= 16, pix_code in 0..2)");
  if (n_items >= (1ll << 31)) return fail(DVA_EUNSUPPORTED, "mapping_build: more than 2^31 items");
  if (!view_ptr || !counts || !atomic_ptr || !workspace) return fail(DVA_EINVAL, "mapping_build: null pointer");
  if (n_items > 0 && (!point_ids || !image_ids || !pixels || !images_out || !pixels_out || (feat && !feat_out)))
    return fail(DVA_EINVAL, "mapping_build: null pointer");
  if (workspace_bytes < dva_mapping_build_workspace_bytes(n_items, num_points)) return fail(DVA_EINVAL, "mapping_build: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  mb::Workspace w;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  mb::carve(base, n_items, num_points, &w);
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, (size_t)((uint8_t*)w.n_views - (uint8_t*)w.cnt), st);   // cnt + cursor
  if (e == cudaSuccess) e = cudaMemsetAsync(w.status, 0, 256, st);
  if (e != cudaSuccess) return fail((int)e, "mapping_build: memset failed");
  int rc;
  if (n_items > 0) {
    mb::count_points<<<mb::grid_for(n_items), 256, 0, st>>>(point_ids, n_items, num_points, w.cnt, w.status);
    if ((rc = check_launch("mb_count_points"))) return rc;
  }
  if ((rc = mb::exclusive_scan(w.cnt, num_points, w.off, w.block_sums, st))) return rc;
  if (n_items > 0) {
    mb::scatter_items<<<mb::grid_for(n_items), 256, 0, st>>>(point_ids, n_items, num_points, w.off, w.cursor, w.bucket);
    if ((rc = check_launch("mb_scatter_items"))) return rc;
  }
  const int pgrid = mb::grid_for(num_points * 32);
#define MB_PIX(CALL) do { if (pix_code == 0) { using PIX = int16_t; CALL; } else if (pix_code == 1) { using PIX = int32_t; CALL; } else { using PIX = int64_t; CALL; } } while (0)
  if (num_points > 0) {
    MB_PIX((mb::order_points<PIX><<<pgrid, 256, 0, st>>>(image_ids, pixels, w.off, w.bucket, w.sorted, w.flags, w.n_views,
                                                          w.n_pix, num_points, dedupe_pixels)));
    if ((rc = check_launch("mb_order_points"))) return rc;
  }
  if ((rc = mb::exclusive_scan(w.n_views, num_points, view_ptr, w.block_sums, st))) return rc;
  if ((rc = mb::exclusive_scan(w.n_pix, num_points, w.pix_ptr, w.block_sums, st))) return rc;
  if (num_points > 0 && n_items > 0) {
    MB_PIX((mb::emit_points<PIX><<<pgrid, 256, 0, st>>>(image_ids, pixels, feat, feat_row, feat_on, (int)F, w.off, w.sorted,
                                                         w.flags, view_ptr, w.pix_ptr, num_points, images_out, atomic_ptr,
                                                         pixels_out, feat_out, order_out)));
    if ((rc = check_launch("mb_emit_points"))) return rc;
  }
#undef MB_PIX
  mb::finish_counts<<<1, 1, 0, st>>>(view_ptr, w.pix_ptr, num_points, atomic_ptr, counts, w.status);
  return check_launch("mb_finish_counts");
}

// sorting [V_total] and csr_cat [N + 1] of S settings whose view pointers (device arrays [N + 1]) are listed in
// ptrs (device array of S device pointers); bases[s] = number of views of the settings before s.
extern "C" int dva_view_cat_sorting(const int64_t* const* ptrs, const int64_t* bases, int64_t S, int64_t N,
                                    int64_t* sorting, int64_t* csr_cat, void* stream) {
  if (S < 1 || N < 0) return fail(DVA_EINVAL, "view_cat_sorting: bad sizes");
  if (!ptrs || !bases || !csr_cat) return fail(DVA_EINVAL, "view_cat_sorting: null pointer");
  mb::view_cat_sorting_kernel<<<mb::grid_for(N + 1), 256, 0, (cudaStream_t)stream>>>(ptrs, bases, (int)S, N, sorting, csr_cat);
  return check_launch("view_cat_sorting");
}
