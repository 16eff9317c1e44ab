// Ring implementation of the fused view-attention pair (same math and C ABI as view_attention.cu;
// reference chain: modules.py:518 row gather -> pooling.py:285-300 / 515-530).
//
// Why a second implementation: the streaming kernels keep the row chunks of ONE point in registers,
// so every point pays the dependent-load chain ptr -> scores -> row ids -> rows before its bytes
// are in flight.  With the short segments and <= 512-byte rows of the shipped configs (S3DIS: ~8
// views x 64 ch, KITTI-360: ~20 views x 128 ch, bf16 storage) that chain, not HBM, sets the pace
// (24-45 % of the measured HBM peak).  Here a warp owns a contiguous RANGE of points, i.e. a
// contiguous range of views, and streams it through a shared-memory ring as fixed-size BATCHES of
// rows that ignore point boundaries:
//
//   producer side (same warp, S-1 batches ahead of the consumer):
//     row ids of batch b+1     coalesced LDG into one register per lane (prefetched a batch early)
//     rows of batch b          one 16-byte cp.async (LDGSTS) per lane and row step, no registers held
//     scores of batch b        cp.async into the batch's score tile, same commit group
//   consumer side: walks the points of the range; a segment is cut into PIECES (its intersection
//     with a batch); forward uses an online softmax over the pieces (running max / denominator,
//     accumulator rescaled when the max moves), so a segment never has to be resident as a whole.
//   backward only: the upstream-gradient rows and the saved softmax statistics of the next group of
//     points are fetched by ONE elected lane with bulk async copies (cp.async.bulk -> UBLKCP,
//     completion on an mbarrier) -- they are contiguous in memory, the natural TMA case.
//
// Bytes in flight per SM = warps x (S-1) x 8 KB regardless of segment length, and nothing in the
// consumer waits on a global load except at range boundaries.
// HBM bytes per launch are those of the streaming kernels (see view_attention.cu).
#include "view_attention.cuh"

namespace dva {

#ifndef DVA_RING_STAGES
#define DVA_RING_STAGES 3
#endif
#ifndef DVA_RING_WARPS
#define DVA_RING_WARPS 4
#endif
#ifndef DVA_RING_BATCH_BYTES
#define DVA_RING_BATCH_BYTES 2048
#endif
#ifndef DVA_RING_FWD_MINB
#define DVA_RING_FWD_MINB 8       // CTAs per SM the register budget is sized for (x kRingWarps warps)
#endif
#ifndef DVA_RING_BWD_MINB
#define DVA_RING_BWD_MINB 4
#endif
#ifndef DVA_RING_RANGES_PER_WARP
#define DVA_RING_RANGES_PER_WARP 1
#endif
constexpr int kRingStages = DVA_RING_STAGES;
constexpr int kRingWarps = DVA_RING_WARPS;     // warps of a CTA never synchronise with each other

template <int LPR> struct RingGeom {
  static constexpr int RPI = 32 / LPR;                       // rows per warp step
  static constexpr int RS = LPR * 16;                        // row stride in the ring (bytes)
  static constexpr int RB = (DVA_RING_BATCH_BYTES / RS) < 32 ? (DVA_RING_BATCH_BYTES / RS) : 32;  // rows per batch
  static constexpr int STEPS = RB / RPI;                     // row steps per batch
  static_assert(RB >= RPI && RB % RPI == 0 && (RB & (RB - 1)) == 0, "batch geometry");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// ---- mbarrier + bulk async copy (backward: contiguous per-point-group tiles)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint32_t load_row_id(const void* idx, int idx64, int64_t v) {
  if (idx == nullptr) return (uint32_t)v;
  return idx64 ? (uint32_t) reinterpret_cast<const int64_t*>(idx)[v]
               : (uint32_t) reinterpret_cast<const int32_t*>(idx)[v];
}


// per-warp shared-memory layout (bytes); fwd: rows | scores | att tile; bwd adds the s' tile, the
// row-id tile, the per-point-group tiles (2 buffers) and two mbarriers
template <int LPR> struct RingSmem {
  using Gm = RingGeom<LPR>;
  size_t rows, comp, att, s_tile, rowid, gout, stats, bars, total;
  __host__ __device__ RingSmem(int G, bool bwd) {
    rows = (size_t)kRingStages * Gm::RB * Gm::RS;
    comp = (size_t)kRingStages * Gm::RB * G * sizeof(float);
    att = (((size_t)G * kTileStride * sizeof(float)) + 15) & ~(size_t)15;
    s_tile = bwd ? att : 0;
    rowid = bwd ? (size_t)kRingStages * Gm::RB * sizeof(uint32_t) : 0;
    gout = bwd ? (size_t)2 * Gm::RB * Gm::RS : 0;
    stats = bwd ? (size_t)2 * 3 * Gm::RB * G * sizeof(float) : 0;   // [buf][max|den|arg][PG*G]
    bars = bwd ? 16 : 0;
    total = rows + comp + att + s_tile + rowid + gout + stats + bars;
  }
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(kRingWarps * 32, DVA_RING_FWD_MINB)
va_ring_fwd_kernel(const VAParams P, const int PR) {
  using Gm = RingGeom<LPR>;
  constexpr int VEC = Vec16<T>::N, RPI = Gm::RPI, RB = Gm::RB, RS = Gm::RS, S = kRingStages;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = P.C, G = P.G;
  const int lg = 31 - __clz(G);                     // G is a power of two
  const RingSmem<LPR> L(G, false);
  unsigned char* base = smem_raw + (size_t)warp * L.total;
  unsigned char* rows_s = base;
  float* comp_s = reinterpret_cast<float*>(base + L.rows);
  float* att_s = reinterpret_cast<float*>(base + L.rows + L.comp);
  const uint32_t rows_u = smem_u32(rows_s), comp_u = smem_u32(comp_s);

  // dead chunks (lanes past the end of a row) are never copied: keep them zero for good
  for (int q = lane; q < (int)(L.rows / 16); q += 32) reinterpret_cast<uint4*>(rows_s)[q] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();

  const int sg = lane / LPR, lir = lane % LPR;
  const bool live = lir * VEC < C;
  const int gk = group_of_channel(live ? lir * VEC : 0, C, G);   // chunks never straddle groups (host)
  const int gl = lane & (G - 1);
  const uint32_t row_bytes = (uint32_t)C * sizeof(T);
  const char* __restrict__ xb = reinterpret_cast<const char*>(P.x) + (live ? lir * 16 : 0);
  char* __restrict__ ob = reinterpret_cast<char*>(P.out) + (live ? lir * 16 : 0);
  const bool gating = P.gate_w != nullptr;
  const float gw = gating ? P.gate_w[gl] : 0.f, gb = gating ? P.gate_b[gl] : 0.f;
  const bool g16 = (G & 3) == 0;                    // score tiles move as 16-byte copies
  const bool save = P.seg_max != nullptr;

  const int64_t n_ranges = (P.N + PR - 1) / PR;
  const int64_t warps_total = (int64_t)gridDim.x * kRingWarps;
  for (int64_t r = (int64_t)blockIdx.x * kRingWarps + warp; r < n_ranges; r += warps_total) {
    const int64_t pa = r * PR;
    const int64_t pb = (pa + PR < P.N) ? pa + PR : P.N;
    const int64_t vb = P.ptr[pa];
    const int nv = (int)(P.ptr[pb] - vb);           // views of this range (host: V < 2^31)
    uint32_t rid = (lane < RB && lane < nv) ? load_row_id(P.idx, P.idx64, vb + lane) : 0u;

    // batch bi of this range -> slot bi % S; then prefetch the row ids of batch bi + 1
    auto issue = [&](int bi) {
      const int slot = bi % S;
      const int v0 = bi * RB;
      const int nrows = (nv - v0 < RB) ? nv - v0 : RB;
      if (nrows > 0) {
        const uint32_t dst0 = rows_u + (uint32_t)(slot * RB) * RS + lir * 16;
#pragma unroll
        for (int st = 0; st < Gm::STEPS; ++st) {
          const int rr = st * RPI + sg;
          const uint32_t srow = __shfl_sync(FULL, rid, rr);
          if (rr < nrows && live) cp_async16(dst0 + rr * RS, xb + (uint64_t)srow * row_bytes);
        }
        const float* csrc = P.compat + (vb + v0) * G;
        const uint32_t cdst = comp_u + (uint32_t)(slot * RB * G) * 4;
        if (g16) {
          for (int q = lane; q < (nrows << lg) >> 2; q += 32) cp_async16(cdst + q * 16, csrc + q * 4);
        } else {
          for (int e = lane; e < (nrows << lg); e += 32) cp_async4(cdst + e * 4, csrc + e);
        }
      }
      cp_async_commit();
      const int v1 = v0 + RB;
      rid = (lane < RB && v1 + lane < nv) ? load_row_id(P.idx, P.idx64, vb + v1 + lane) : 0u;
    };
#pragma unroll
    for (int bi = 0; bi < S - 1; ++bi) issue(bi);
    int cb = -1;                                    // newest batch known to be resident

    // relative pointers of the first group of 32 points: lane k holds point pg + k
    int pl, cnt;
    {
      const int64_t q0 = (pa + lane < pb) ? pa + lane : pb;
      const int64_t q1 = (q0 + 1 < pb) ? q0 + 1 : pb;
      pl = (int)(P.ptr[q0] - vb);
      cnt = (int)(P.ptr[q1] - vb) - pl;
    }
    for (int64_t pg = pa; pg < pb; pg += 32) {
      int pl_n = 0, cnt_n = 0;                      // next group's pointers: loaded a group early
      if (pg + 32 < pb) {
        const int64_t q0 = (pg + 32 + lane < pb) ? pg + 32 + lane : pb;
        const int64_t q1 = (q0 + 1 < pb) ? q0 + 1 : pb;
        pl_n = (int)(P.ptr[q0] - vb);
        cnt_n = (int)(P.ptr[q1] - vb) - pl_n;
      }
      const int kmax = (pb - pg < 32) ? (int)(pb - pg) : 32;
      for (int k = 0; k < kmax; ++k) {
        const int s = __shfl_sync(FULL, pl, k);
        const int n = __shfl_sync(FULL, cnt, k);
        const int64_t i = pg + k;
        if (n == 0) {                               // unseen point: exact zeros
          if (lane < G && save) {
            P.seg_max[i * G + lane] = 0.f; P.seg_den[i * G + lane] = P.eps; P.seg_arg[i * G + lane] = -1;
          }
          if (sg == 0 && live) {
            float z[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) z[j] = 0.f;
            stg_stream16(ob + i * (int64_t)row_bytes, pack16<T, VEC>(z));
          }
          continue;
        }
        const float inv_sq = P.group_scaling ? rsqrtf((float)n) : 1.f;
        float m_run = -INFINITY, den = 0.f;
        int am = -1;                                // first arg-max view (global index), group gl
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const int vend = s + n;
        for (int v = s; v < vend;) {
          const int b = v / RB, lo = v % RB;
          const int np = (RB - lo < vend - v) ? RB - lo : vend - v;
          while (cb < b) {                          // make batch b resident; refill the freed slot
            __syncwarp();
            ++cb;
            issue(cb + S - 1);
            cp_async_wait<S - 1>();
            __syncwarp();
          }
          const int slot = b % S;
          const float* cs = comp_s + ((slot * RB + lo) << lg);
          const int npG = np << lg;
          // piece max + first arg-max (lane owns the elements e = lane, lane+32, ... of group gl)
          float pm = -INFINITY; int pe = 0x7fffffff;
          for (int e = lane; e < npG; e += 32) {
            const float c = cs[e];
            if (c > pm) { pm = c; pe = e; }
          }
          for (int off = 16; off >= G; off >>= 1) {
            const float om = __shfl_xor_sync(FULL, pm, off);
            const int oe = __shfl_xor_sync(FULL, pe, off);
            if (om > pm || (om == pm && oe < pe)) { pm = om; pe = oe; }
          }
          if (pm > m_run) am = (int)(vb + v) + (pe >> lg);
          const float m_new = fmaxf(m_run, pm);
          const float alpha = expf((m_run - m_new) * inv_sq);   // first piece: exp(-inf) = 0
          m_run = m_new;
          den *= alpha;
          __syncwarp();                             // readers of the previous piece's att tile are done
          for (int e = lane; e < npG; e += 32) {
            const float ev = expf((cs[e] - m_new) * inv_sq);
            den += ev;
            att_s[gl * kTileStride + (e >> lg)] = ev;
          }
          __syncwarp();
          if (v != s) {                             // online softmax: rescale what was accumulated
            const float ak = __shfl_sync(FULL, alpha, gk);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] *= ak;
          }
          const unsigned char* rp = rows_s + (size_t)(slot * RB + lo) * RS + lir * 16;
          const float* ap = att_s + gk * kTileStride;
#pragma unroll 4
          for (int v0 = sg; v0 < np + sg; v0 += RPI) {   // warp-uniform trip count; idle sub-groups add 0
            const bool ok = v0 < np;
            const int vv = ok ? v0 : 0;
            const uint4 raw = *reinterpret_cast<const uint4*>(rp + (size_t)vv * RS);
            const float a = ok ? ap[vv] : 0.f;
            float fv[VEC];
            unpack16<T, VEC>(raw, fv);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(a, fv[j], acc[j]);
          }
          v += np;
        }
        den = group_lane_sum(den, G) + P.eps;
        const float t = gating ? tanhf(fmaxf(fmaf(gw, m_run, gb), 0.f)) : 1.f;
        if (lane < G && save) {
          P.seg_max[i * G + lane] = m_run; P.seg_den[i * G + lane] = den; P.seg_arg[i * G + lane] = am;
        }
        const float inv_den = 1.f / den;
        const float sc = __shfl_sync(FULL, t * inv_den, gk);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float a = acc[j];
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
          acc[j] = a * sc;
        }
        if (sg == 0 && live) stg_stream16(ob + i * (int64_t)row_bytes, pack16<T, VEC>(acc));
        if (P.att != nullptr) {                     // normalised attentions (autograd / save_last tap)
          float* __restrict__ ao = P.att + (vb + s) * G;
          const int nG = n << lg;
          if (n <= RB - (s % RB)) {                 // single piece: its e-values are still in the tile
            for (int e = lane; e < nG; e += 32) ao[e] = att_s[gl * kTileStride + (e >> lg)] * inv_den;
          } else {
            const float* __restrict__ cp = P.compat + (vb + s) * G;
            for (int e = lane; e < nG; e += 32) ao[e] = expf((__ldg(cp + e) - m_run) * inv_sq) * inv_den;
          }
        }
      }
      pl = pl_n; cnt = cnt_n;
    }
    cp_async_wait<0>();                             // only empty groups can be left; then reuse the ring
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward (math: see view_attention.cu; regular group layout only)
// ---------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(kRingWarps * 32, DVA_RING_BWD_MINB)
va_ring_bwd_kernel(const VAParams P, const int PR) {
  using Gm = RingGeom<LPR>;
  constexpr int VEC = Vec16<T>::N, RPI = Gm::RPI, RB = Gm::RB, RS = Gm::RS, S = kRingStages;
  constexpr int PG = RB;                            // points per group (one bulk-copied tile)
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float gate_s[kRingWarps][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = P.C, G = P.G;
  const int lg = 31 - __clz(G);
  const RingSmem<LPR> L(G, true);
  unsigned char* base = smem_raw + (size_t)warp * L.total;
  unsigned char* rows_s = base;
  float* comp_s = reinterpret_cast<float*>(base + L.rows);
  float* att_s = reinterpret_cast<float*>(base + L.rows + L.comp);
  float* s_s = reinterpret_cast<float*>(base + L.rows + L.comp + L.att);
  uint32_t* rowid_s = reinterpret_cast<uint32_t*>(base + L.rows + L.comp + L.att + L.s_tile);
  unsigned char* gout_s = base + L.rows + L.comp + L.att + L.s_tile + L.rowid;
  float* stats_s = reinterpret_cast<float*>(gout_s + L.gout);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gout_s + L.gout + L.stats);
  const uint32_t rows_u = smem_u32(rows_s), comp_u = smem_u32(comp_s);
  const uint32_t gout_u = smem_u32(gout_s), stats_u = smem_u32(stats_s);
  const uint32_t bar_u[2] = {smem_u32(bars), smem_u32(bars + 1)};

  for (int q = lane; q < (int)(L.rows / 16); q += 32) reinterpret_cast<uint4*>(rows_s)[q] = make_uint4(0u, 0u, 0u, 0u);
  for (int q = lane; q < (int)(L.gout / 16); q += 32) reinterpret_cast<uint4*>(gout_s)[q] = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) {
    mbar_init(bar_u[0], 1);
    mbar_init(bar_u[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncwarp();

  const int sg = lane / LPR, lir = lane % LPR;
  const bool live = lir * VEC < C;
  const int gk = group_of_channel(live ? lir * VEC : 0, C, G);
  const int gl = lane & (G - 1);
  const uint32_t row_bytes = (uint32_t)C * sizeof(T);
  const char* __restrict__ xb = reinterpret_cast<const char*>(P.x) + (live ? lir * 16 : 0);
  char* __restrict__ gxb = reinterpret_cast<char*>(P.gx) + (live ? lir * 16 : 0);
  const bool gating = P.gate_w != nullptr;
  const float gw = gating ? P.gate_w[gl] : 0.f, gb = gating ? P.gate_b[gl] : 0.f;
  float dw_acc = 0.f, db_acc = 0.f;
  const bool has_idx = P.idx != nullptr;
  const bool scatter = P.scatter && has_idx;
  const bool padded = row_bytes != (uint32_t)RS;    // rows narrower than the ring stride: per-row tile copies
  // lanes of one row step that share a group form aligned blocks of cpe lanes (regular layout)
  const int cpg = (C / G) / VEC;
  const int cpe = cpg < LPR ? cpg : LPR;
  float red_mask[5];
#pragma unroll
  for (int b = 0; b < 5; ++b) red_mask[b] = ((1 << b) < cpe) ? 1.f : 0.f;
  const bool leader = (lir & (cpe - 1)) == 0;
  uint32_t uses0 = 0, uses1 = 0;                    // completed uses of tile buffer 0 / 1 (mbarrier parity)

  const int64_t n_ranges = (P.N + PR - 1) / PR;
  const int64_t warps_total = (int64_t)gridDim.x * kRingWarps;
  for (int64_t r = (int64_t)blockIdx.x * kRingWarps + warp; r < n_ranges; r += warps_total) {
    const int64_t pa = r * PR;
    const int64_t pb = (pa + PR < P.N) ? pa + PR : P.N;
    const int64_t vb = P.ptr[pa];
    const int nv = (int)(P.ptr[pb] - vb);
    uint32_t rid = (lane < RB && lane < nv) ? load_row_id(P.idx, P.idx64, vb + lane) : 0u;

    auto issue = [&](int bi) {
      const int slot = bi % S;
      const int v0 = bi * RB;
      const int nrows = (nv - v0 < RB) ? nv - v0 : RB;
      if (nrows > 0) {
        const uint32_t dst0 = rows_u + (uint32_t)(slot * RB) * RS + lir * 16;
#pragma unroll
        for (int st = 0; st < Gm::STEPS; ++st) {
          const int rr = st * RPI + sg;
          const uint32_t srow = __shfl_sync(FULL, rid, rr);
          if (rr < nrows && live) cp_async16(dst0 + rr * RS, xb + (uint64_t)srow * row_bytes);
        }
        if (lane < nrows) rowid_s[slot * RB + lane] = scatter ? rid : (uint32_t)(vb + v0 + lane);   // dx row
        const float* csrc = P.compat + (vb + v0) * G;
        const uint32_t cdst = comp_u + (uint32_t)(slot * RB * G) * 4;
        for (int q = lane; q < (nrows << lg) >> 2; q += 32) cp_async16(cdst + q * 16, csrc + q * 4);   // G % 4 == 0
      }
      cp_async_commit();
      const int v1 = v0 + RB;
      rid = (lane < RB && v1 + lane < nv) ? load_row_id(P.idx, P.idx64, vb + v1 + lane) : 0u;
    };
    // tile of point group [p0, p0+np): grad_out rows + saved statistics -> buffer j, one mbarrier phase
    auto fetch_group = [&](int64_t p0, int j) {
      const int np = (pb - p0 < PG) ? (int)(pb - p0) : PG;
      const uint32_t sbytes = (uint32_t)(np << lg) * 4;
      if (lane == 0) {
        fence_proxy_async();                        // earlier generic reads of this buffer come first
        mbar_expect_tx(bar_u[j], (uint32_t)np * row_bytes + 3 * sbytes);
        const uint32_t sdst = stats_u + (uint32_t)(j * 3 * PG * G) * 4;
        bulk_g2s(sdst, P.s_max + p0 * G, sbytes, bar_u[j]);
        bulk_g2s(sdst + (uint32_t)(PG * G) * 4, P.s_den + p0 * G, sbytes, bar_u[j]);
        bulk_g2s(sdst + (uint32_t)(2 * PG * G) * 4, P.s_arg + p0 * G, sbytes, bar_u[j]);
        const char* gsrc = reinterpret_cast<const char*>(P.gout) + p0 * (int64_t)row_bytes;
        const uint32_t gdst = gout_u + (uint32_t)(j * PG) * RS;
        if (!padded) {
          bulk_g2s(gdst, gsrc, (uint32_t)np * row_bytes, bar_u[j]);
        } else {
          for (int q = 0; q < np; ++q) bulk_g2s(gdst + q * RS, gsrc + (size_t)q * row_bytes, row_bytes, bar_u[j]);
        }
      }
    };
#pragma unroll
    for (int bi = 0; bi < S - 1; ++bi) issue(bi);
    int cb = -1;
    __syncwarp();
    fetch_group(pa, 0);
    int jbuf = 0;

    // relative pointers of the first group: lane k < PG holds point pg + k
    int pl, cnt;
    {
      const int64_t q0 = (pa + lane < pb) ? pa + lane : pb;
      const int64_t q1 = (q0 + 1 < pb) ? q0 + 1 : pb;
      pl = (int)(P.ptr[q0] - vb);
      cnt = (lane < PG) ? (int)(P.ptr[q1] - vb) - pl : 0;
    }
    for (int64_t pg = pa; pg < pb; pg += PG, jbuf ^= 1) {
      int pl_n = 0, cnt_n = 0;                      // next group's pointers: loaded a group early
      if (pg + PG < pb) {
        const int64_t q0 = (pg + PG + lane < pb) ? pg + PG + lane : pb;
        const int64_t q1 = (q0 + 1 < pb) ? q0 + 1 : pb;
        pl_n = (int)(P.ptr[q0] - vb);
        cnt_n = (lane < PG) ? (int)(P.ptr[q1] - vb) - pl_n : 0;
      }
      __syncwarp();                                 // every lane is done with the other buffer
      if (pg + PG < pb) fetch_group(pg + PG, jbuf ^ 1);
      {
        const uint32_t par = (jbuf ? uses1 : uses0) & 1u;
        mbar_wait(bar_u[jbuf], par);
        if (jbuf) ++uses1; else ++uses0;
      }
      const float* st_max = stats_s + jbuf * 3 * PG * G;
      const float* st_den = st_max + PG * G;
      const int32_t* st_arg = reinterpret_cast<const int32_t*>(st_max + 2 * PG * G);
      const int kmax = (pb - pg < PG) ? (int)(pb - pg) : PG;
      for (int k = 0; k < kmax; ++k) {
        const int s = __shfl_sync(FULL, pl, k);
        const int n = __shfl_sync(FULL, cnt, k);
        if (n == 0) continue;                       // no view: nothing flows back
        const int nG = n << lg;
        const float m = st_max[(k << lg) + gl];
        const float inv_den = 1.f / st_den[(k << lg) + gl];
        const int arg_v = st_arg[(k << lg) + gl];
        const float inv_sq = P.group_scaling ? rsqrtf((float)n) : 1.f;
        const float z = fmaf(gw, m, gb);
        const float t = gating ? tanhf(fmaxf(z, 0.f)) : 1.f;
        // gd = dO * t of this lane's channels (dead lanes: zero tile, zero gd)
        float gd[VEC];
        {
          const uint4 raw = *reinterpret_cast<const uint4*>(gout_s + (size_t)(jbuf * PG + k) * RS + lir * 16);
          unpack16<T, VEC>(raw, gd);
          const float tk = __shfl_sync(FULL, t, gk);
#pragma unroll
          for (int j = 0; j < VEC; ++j) gd[j] = live ? gd[j] * tk : 0.f;
        }
        float* __restrict__ gc = P.gcompat + (vb + s) * G;
        const bool single = n <= RB - (s % RB);
        float Ssum = 0.f;                           // sum_v a_vg s'_vg, group gl (per-lane partial)
        const int vend = s + n;
        for (int v = s; v < vend;) {
          const int b = v / RB, lo = v % RB;
          const int np = (RB - lo < vend - v) ? RB - lo : vend - v;
          while (cb < b) {
            __syncwarp();
            ++cb;
            issue(cb + S - 1);
            cp_async_wait<S - 1>();
            __syncwarp();
          }
          const int slot = b % S;
          const float* cs = comp_s + ((slot * RB + lo) << lg);
          const int npG = np << lg;
          __syncwarp();                             // previous piece's tiles are free
          for (int e = lane; e < npG; e += 32)
            att_s[gl * kTileStride + (e >> lg)] = expf((cs[e] - m) * inv_sq) * inv_den;
          __syncwarp();
          const unsigned char* rp = rows_s + (size_t)(slot * RB + lo) * RS + lir * 16;
          const uint32_t* op = rowid_s + slot * RB + lo;
          const float* ap = att_s + gk * kTileStride;
          float* sp = s_s + gk * kTileStride;
#pragma unroll 2
          for (int v0 = sg; v0 < np + sg; v0 += RPI) {   // warp-uniform trip count
            const bool ok = v0 < np;
            const int vv = ok ? v0 : 0;
            const uint4 raw = *reinterpret_cast<const uint4*>(rp + (size_t)vv * RS);
            const float a = ap[vv];
            const uint32_t orow = op[vv];
            float fv[VEC], dx[VEC];
            unpack16<T, VEC>(raw, fv);
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              dot = fmaf(gd[j], fv[j], dot);
              dx[j] = a * gd[j];
            }
            if (ok && live) stg_stream16(gxb + (uint64_t)orow * row_bytes, pack16<T, VEC>(dx));
            float rsum = dot;
#pragma unroll
            for (int bb = 0; (1 << bb) < LPR; ++bb)
              rsum = fmaf(__shfl_xor_sync(FULL, rsum, 1 << bb), red_mask[bb], rsum);
            if (ok && live && leader) sp[vv] = rsum;
          }
          __syncwarp();
          for (int e = lane; e < npG; e += 32) {
            const int sl = gl * kTileStride + (e >> lg);
            const float sv = s_s[sl];
            Ssum = fmaf(att_s[sl], sv, Ssum);
            if (!single) gc[((v - s) << lg) + e] = sv;   // raw s': finalised below once S is complete
          }
          v += np;
        }
        Ssum = group_lane_sum(Ssum, G);
        const float one_m_t2 = 1.f - t * t;
        const float dLdt = (t != 0.f) ? Ssum / t : 0.f;
        const bool open = gating && z > 0.f;
        const float dq = open ? dLdt * one_m_t2 * gw : 0.f;
        if (open && lane < G) {
          dw_acc += dLdt * one_m_t2 * m;
          db_acc += dLdt * one_m_t2;
        }
        const int first_view = (int)(vb + s);
        if (single) {
          for (int e = lane; e < nG; e += 32) {
            const int sl = gl * kTileStride + (e >> lg);
            float d = att_s[sl] * (s_s[sl] - Ssum) * inv_sq;
            if (first_view + (e >> lg) == arg_v) d += dq;
            gc[e] = d;
          }
        } else {
          __syncwarp();                             // raw s' written by other lanes of this warp
          const float* __restrict__ cp = P.compat + (vb + s) * G;
          for (int e = lane; e < nG; e += 32) {
            const float a = expf((__ldg(cp + e) - m) * inv_sq) * inv_den;
            float d = a * (__ldcg(gc + e) - Ssum) * inv_sq;
            if (first_view + (e >> lg) == arg_v) d += dq;
            gc[e] = d;
          }
        }
      }
      pl = pl_n; cnt = cnt_n;
    }
    cp_async_wait<0>();
    __syncwarp();
  }

  // ---- gate parameter gradients: warp -> block partial (deterministic), block -> workspace
  if (P.gate_partial != nullptr) {
    if (lane < G) { gate_s[warp][0][lane] = dw_acc; gate_s[warp][1][lane] = db_acc; }
    __syncthreads();
    if ((int)threadIdx.x < 2 * G) {
      const int which = threadIdx.x / G, g = threadIdx.x % G;
      float acc = 0.f;
      for (int w = 0; w < kRingWarps; ++w) acc += gate_s[w][which][g];
      P.gate_partial[(int64_t)blockIdx.x * 2 * G + threadIdx.x] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T> static int ring_lpr(const VAParams& P) {
  const int cv = P.C / Vec16<T>::N;                 // 16-byte chunks per row
  if (cv <= 4) return 4;
  if (cv <= 8) return 8;
  if (cv <= 16) return 16;
  return 32;
}

template <typename T>
static bool ring_common_ok(const VAParams& P, const void* o1, const void* o2) {
  constexpr int V16 = Vec16<T>::N;
  const int C = P.C, G = P.G;
  if (C % V16 != 0 || C / V16 > 32) return false;   // rows of at most 512 bytes, whole 16-byte chunks
  if (!aligned16(P.x) || !aligned16(o1) || (o2 != nullptr && !aligned16(o2))) return false;
  if (!aligned16(P.compat)) return false;
  if (P.V >= (1ll << 31) || P.R >= (1ll << 32)) return false;
  for (int c0 = 0; c0 < C; c0 += V16)               // chunks never straddle channel groups
    if (group_of_channel(c0, C, G) != group_of_channel(c0 + V16 - 1, C, G)) return false;
  return true;
}

template <typename T> static bool ring_fwd_ok(const VAParams& P) { return ring_common_ok<T>(P, P.out, nullptr); }
template <typename T> static bool ring_bwd_ok(const VAParams& P) {
  constexpr int V16 = Vec16<T>::N;
  if (!ring_common_ok<T>(P, P.gout, P.gx)) return false;
  if (P.G % 4 != 0) return false;                   // statistic tiles move as 16-byte multiples
  if (P.C % P.G != 0 || (P.C / P.G) % V16 != 0) return false;
  const int cpg = (P.C / P.G) / V16;
  if ((cpg & (cpg - 1)) != 0) return false;         // regular layout (see view_attention.cu)
  if (!aligned16(P.s_max) || !aligned16(P.s_den) || !aligned16(P.s_arg)) return false;
  return true;
}

bool va_ring_fwd_applicable(const VAParams& P, int dtype) {
  switch (dtype) {
    case DVA_F32: return ring_fwd_ok<float>(P);
    case DVA_BF16: return ring_fwd_ok<__nv_bfloat16>(P);
    case DVA_F16: return ring_fwd_ok<__half>(P);
    default: return false;
  }
}
bool va_ring_bwd_applicable(const VAParams& P, int dtype) {
  switch (dtype) {
    case DVA_F32: return ring_bwd_ok<float>(P);
    case DVA_BF16: return ring_bwd_ok<__nv_bfloat16>(P);
    case DVA_F16: return ring_bwd_ok<__half>(P);
    default: return false;
  }
}

// grid = co-resident CTAs (148 SMs x occupancy); PR = points per range
template <typename K>
static int ring_launch_geometry(K kern, size_t smem, int64_t N, int* grid_out, int* pr_out) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return failf((int)e, "view_attention ring: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRingWarps * 32, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > 8) occ = 8;                             // gate-gradient workspace: 148 x 8 partials
  int64_t grid = (int64_t)kNumSMs * occ;
  const int64_t warps = grid * kRingWarps;
  int64_t pr = (N + warps * DVA_RING_RANGES_PER_WARP - 1) / (warps * DVA_RING_RANGES_PER_WARP);
  if (pr < 8) pr = 8;
  const int64_t n_ranges = (N + pr - 1) / pr;
  const int64_t need = (n_ranges + kRingWarps - 1) / kRingWarps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid; *pr_out = (int)pr;
  return DVA_OK;
}

template <typename T, int LPR>
static int ring_fwd_launch(const VAParams& P, cudaStream_t st) {
  const RingSmem<LPR> L(P.G, false);
  const size_t smem = L.total * kRingWarps;
  auto kern = va_ring_fwd_kernel<T, LPR>;
  int grid, pr;
  if (int rc = ring_launch_geometry(kern, smem, P.N, &grid, &pr)) return rc;
  kern<<<grid, kRingWarps * 32, smem, st>>>(P, pr);
  return check_launch("view_attention_fwd(ring)");
}
template <typename T, int LPR>
static int ring_bwd_launch(const VAParams& P, int* grid_out, cudaStream_t st) {
  const RingSmem<LPR> L(P.G, true);
  const size_t smem = L.total * kRingWarps;
  auto kern = va_ring_bwd_kernel<T, LPR>;
  int grid, pr;
  if (int rc = ring_launch_geometry(kern, smem, P.N, &grid, &pr)) return rc;
  *grid_out = grid;
  kern<<<grid, kRingWarps * 32, smem, st>>>(P, pr);
  return check_launch("view_attention_bwd(ring)");
}

template <typename T> static int ring_fwd_typed(const VAParams& P, cudaStream_t st) {
  switch (ring_lpr<T>(P)) {
    case 4: return ring_fwd_launch<T, 4>(P, st);
    case 8: return ring_fwd_launch<T, 8>(P, st);
    case 16: return ring_fwd_launch<T, 16>(P, st);
    default: return ring_fwd_launch<T, 32>(P, st);
  }
}
template <typename T> static int ring_bwd_typed(const VAParams& P, int* grid, cudaStream_t st) {
  switch (ring_lpr<T>(P)) {
    case 4: return ring_bwd_launch<T, 4>(P, grid, st);
    case 8: return ring_bwd_launch<T, 8>(P, grid, st);
    case 16: return ring_bwd_launch<T, 16>(P, grid, st);
    default: return ring_bwd_launch<T, 32>(P, grid, st);
  }
}

int va_ring_fwd(const VAParams& P, int dtype, cudaStream_t st) {
  switch (dtype) {
    case DVA_F32: return ring_fwd_typed<float>(P, st);
    case DVA_BF16: return ring_fwd_typed<__nv_bfloat16>(P, st);
    case DVA_F16: return ring_fwd_typed<__half>(P, st);
    default: return fail(DVA_EINVAL, "view_attention_fwd: unknown dtype");
  }
}
int va_ring_bwd(const VAParams& P, int dtype, int* grid_out, cudaStream_t st) {
  switch (dtype) {
    case DVA_F32: return ring_bwd_typed<float>(P, grid_out, st);
    case DVA_BF16: return ring_bwd_typed<__nv_bfloat16>(P, grid_out, st);
    case DVA_F16: return ring_bwd_typed<__half>(P, grid_out, st);
    default: return fail(DVA_EINVAL, "view_attention_bwd: unknown dtype");
  }
}

}  // namespace dva
