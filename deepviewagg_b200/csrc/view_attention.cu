// Fused CSR-gather + ragged group softmax + attention-weighted sum + gating (fwd and bwd).
//
// Replaces, in one pass over HBM, the reference chain
//   modules.py:518              x_mod = cat(x_mod)[idx_sorting]            (row gather, [V,C] copy)
//   pooling.py:285-286 / 515    a = segment_softmax_csr(compat, csr, scaling)
//   pooling.py:289-291 / 519    y = segment_csr(x_mod * expand_group_feat(a), csr, 'sum')
//   pooling.py:293-300 / 523    out = y * expand_group_feat(Gating(segment_csr(compat,'max')))
// which materialises >= 4 [V,C] temporaries in the reference.
//
// Work decomposition: one warp owns one point (CSR segment) at a time.  A feature row of C
// channels is split into 16-byte chunks; LPR lanes cover one row (LPR*CPL chunks), so a warp
// reads 32/LPR rows per step with every lane issuing one LDG.128 -- for C=128 fp32 a row is
// exactly one 512 B warp-wide load.  Scores live in the flat (view,group) order of `compat`
// so that lane l always owns group l%G; per-group max / sum are xor-shuffle reductions over
// the lanes of equal l%G.  Everything is fp32 in registers; rows are never re-read.
//
// HBM bytes per launch (s = sizeof(T)):
//   fwd: V*(C*s + 4 + 4G) + N*(8 + C*s) (+ N*12G saved statistics when training)
//   bwd: V*(2*C*s + 4 + 8G) + N*(8 + C*s + 12G)
#include "dva_common.cuh"

namespace dva {

struct VAParams {
  const void* x; const void* idx; int idx64;
  const float* compat; const int64_t* ptr;
  const float* gate_w; const float* gate_b;
  // fwd
  void* out; float* att; float* seg_max; float* seg_den; int32_t* seg_arg;
  // bwd
  const void* gout; const float* s_max; const float* s_den; const int32_t* s_arg;
  void* gx; float* gcompat; float* gate_partial; int scatter;
  int64_t N, V, R;
  int C, G, group_scaling;
  float eps;
};

constexpr int kWarps = 8;          // warps per CTA
constexpr int kUnroll = 8;         // row loads in flight per lane (x CPL)

// A row chunk in flight: the raw 16 bytes (or one scalar) -- unpacked to fp32 only at use so
// that kUnroll loads cost 4 registers each whatever the storage type.
template <typename T, int VEC> struct Chunk {
  uint4 raw;
  __device__ __forceinline__ void load(const T* p) { raw = ldg_stream16(p); }
  __device__ __forceinline__ void zero() { raw = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void get(float (&f)[VEC]) const { unpack16<T, VEC>(raw, f); }
};
template <typename T> struct Chunk<T, 1> {
  T raw;
  __device__ __forceinline__ void load(const T* p) { raw = __ldg(p); }
  __device__ __forceinline__ void zero() { raw = Cvt<T>::from_f(0.f); }
  __device__ __forceinline__ void get(float (&f)[1]) const { f[0] = Cvt<T>::to_f(raw); }
};
template <typename T, int VEC>
__device__ __forceinline__ void load_chunk(const T* p, float (&f)[VEC]) {
  Chunk<T, VEC> c; c.load(p); c.get(f);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_chunk(T* p, const float (&f)[VEC]) {
  if constexpr (VEC == 1) {
    *p = Cvt<T>::from_f(f[0]);
  } else {
    stg_stream16(p, pack16<T, VEC>(f));
  }
}

// reduce over the lanes that share (lane % G): offsets 16 .. G
__device__ __forceinline__ float group_lane_sum(float v, int G) {
  for (int off = 16; off >= G; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int LPR, int CPL>
__global__ void __launch_bounds__(kWarps * 32)
view_attention_fwd_kernel(const VAParams P) {
  constexpr int RPI = 32 / LPR;              // rows per warp step
  constexpr int TILE_C = VEC * LPR * CPL;    // channels per pass
  constexpr int U = (kUnroll / CPL) > 0 ? (kUnroll / CPL) : 1;  // row steps in flight
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = P.C, G = P.G;
  float* att_s = reinterpret_cast<float*>(smem_raw) + warp * (32 * G);
  int64_t* row_s = reinterpret_cast<int64_t*>(smem_raw + (size_t)kWarps * 32 * G * sizeof(float)) + warp * 32;
  const int sg = lane / LPR, lir = lane % LPR;
  const T* __restrict__ x = reinterpret_cast<const T*>(P.x);
  T* __restrict__ out = reinterpret_cast<T*>(P.out);
  const int gl = lane % G;                   // group owned by this lane in the flat score order
  const bool gating = P.gate_w != nullptr;
  const float gw = gating ? P.gate_w[gl] : 0.f, gb = gating ? P.gate_b[gl] : 0.f;

  const int64_t warps_total = (int64_t)gridDim.x * kWarps;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < P.N; i += warps_total) {
    const int64_t p0 = P.ptr[i];
    const int n = (int)(P.ptr[i + 1] - p0);
    const int nG = n * G;
    const float* __restrict__ cp = P.compat + p0 * G;

    // ---- per-group max (first arg-max) and softmax denominator
    float m = -INFINITY; int am = 0x7fffffff;
    for (int e = lane; e < nG; e += 32) {
      const float c = __ldg(cp + e);
      if (c > m) { m = c; am = e / G; }
    }
    for (int off = 16; off >= G; off >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, off);
      const int oa = __shfl_xor_sync(0xffffffffu, am, off);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    if (n == 0) { m = 0.f; am = -1; }        // segment_csr(max) of an empty segment is 0
    const float sq = (P.group_scaling && n > 0) ? sqrtf((float)n) : 1.f;
    float den = 0.f;
    for (int e = lane; e < nG; e += 32) den += expf((__ldg(cp + e) - m) / sq);
    den = group_lane_sum(den, G) + P.eps;
    const float t = gating ? tanhf(fmaxf(fmaf(gw, m, gb), 0.f)) : 1.f;
    if (lane < G && P.seg_max != nullptr) {
      P.seg_max[i * G + lane] = m;
      P.seg_den[i * G + lane] = den;
      P.seg_arg[i * G + lane] = (n > 0) ? (int32_t)(p0 + am) : -1;
    }

    for (int ct = 0; ct < C; ct += TILE_C) {
      float acc[CPL][VEC];
      int gk[CPL][VEC];
#pragma unroll
      for (int k = 0; k < CPL; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          acc[k][j] = 0.f;
          const int c = ct + (lir + LPR * k) * VEC + j;
          gk[k][j] = group_of_channel(c < C ? c : C - 1, C, G);
        }

      for (int vs = 0; vs < n; vs += 32) {
        const int nc = min(32, n - vs);
        __syncwarp();
        for (int e = lane; e < nc * G; e += 32) {
          const float a = expf((__ldg(cp + vs * G + e) - m) / sq) / den;
          att_s[e] = a;
          if (ct == 0 && P.att != nullptr) P.att[(p0 + vs) * G + e] = a;
        }
        if (lane < nc) row_s[lane] = load_idx(P.idx, P.idx64, p0 + vs + lane);
        __syncwarp();

        for (int v0 = 0; v0 < nc; v0 += RPI * U) {
          Chunk<T, VEC> f[U][CPL];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int v = v0 + u * RPI + sg;
            ok[u] = v < nc;
            const T* rp = x + row_s[ok[u] ? v : 0] * (int64_t)C + ct;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const int c0 = (lir + LPR * k) * VEC;
              if (ok[u] && ct + c0 < C) f[u][k].load(rp + c0); else f[u][k].zero();
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int v = ok[u] ? v0 + u * RPI + sg : 0;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              float fv[VEC];
              f[u][k].get(fv);
              // a chunk usually lies inside one group: one LDS broadcast per chunk
              const float a0 = att_s[v * G + gk[k][0]];
#pragma unroll
              for (int j = 0; j < VEC; ++j) {
                const float a = (j == 0 || gk[k][j] == gk[k][0]) ? a0 : att_s[v * G + gk[k][j]];
                acc[k][j] = fmaf(a, fv[j], acc[k][j]);
              }
            }
          }
        }
      }

      // combine the RPI row sub-groups, apply gating, store
#pragma unroll
      for (int k = 0; k < CPL; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float a = acc[k][j];
#pragma unroll
          for (int off = LPR; off < 32; off <<= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          acc[k][j] = a * __shfl_sync(0xffffffffu, t, gk[k][j]);
        }
      if (sg == 0) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int c0 = ct + (lir + LPR * k) * VEC;
          if (c0 < C) store_chunk<T, VEC>(out + i * (int64_t)C + c0, acc[k]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
//   s_vg   = sum_{c in g} dO_c x_vc            (one dot product per view and group)
//   S_g    = sum_v a_vg s_vg   (= d/dt_g)
//   dx_vc  = a_vg t_g dO_c
//   dc_vg  = a_vg t_g (s_vg - S_g)/sqrt(n)  +  [v == argmax_g] S_g (1-t_g^2) w_g 1[w q + b > 0]
//   dw_g  += S_g (1-t^2) 1[.] q_g ;  db_g += S_g (1-t^2) 1[.]
// (the reference obtains the same through autograd over pooling.py:285-300.)
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int LPR, int CPL>
__global__ void __launch_bounds__(kWarps * 32)
view_attention_bwd_kernel(const VAParams P) {
  constexpr int RPI = 32 / LPR;
  constexpr int TILE_C = VEC * LPR * CPL;
  constexpr int U = (kUnroll / 2 / CPL) > 0 ? (kUnroll / 2 / CPL) : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = P.C, G = P.G;
  float* att_s = reinterpret_cast<float*>(smem_raw) + warp * (32 * G);
  float* s_s = reinterpret_cast<float*>(smem_raw) + (kWarps + warp) * (32 * G);
  int64_t* row_s = reinterpret_cast<int64_t*>(smem_raw + (size_t)2 * kWarps * 32 * G * sizeof(float)) + warp * 32;
  float* gate_s = reinterpret_cast<float*>(smem_raw + (size_t)2 * kWarps * 32 * G * sizeof(float) +
                                           (size_t)kWarps * 32 * sizeof(int64_t));  // [kWarps][2][G]
  const int sg = lane / LPR, lir = lane % LPR;
  const T* __restrict__ x = reinterpret_cast<const T*>(P.x);
  const T* __restrict__ gout = reinterpret_cast<const T*>(P.gout);
  T* __restrict__ gx = reinterpret_cast<T*>(P.gx);
  const int gl = lane % G;
  const bool gating = P.gate_w != nullptr;
  const float gw = gating ? P.gate_w[gl] : 0.f, gb = gating ? P.gate_b[gl] : 0.f;
  float dw_acc = 0.f, db_acc = 0.f;
  // How the per-(view,group) dot products are reduced across the lanes of a row:
  //   cpg = 16-byte chunks per group when all groups are equally wide and chunk-aligned.
  const int cpg = (C % G == 0 && (C / G) % VEC == 0) ? (C / G) / VEC : 0;
  const bool cpg_pow2 = cpg > 0 && (cpg & (cpg - 1)) == 0;
  const int red_mode = (cpg_pow2 && cpg <= LPR) ? 1 : ((cpg_pow2 && cpg % LPR == 0) ? 2 : 0);

  const int64_t warps_total = (int64_t)gridDim.x * kWarps;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < P.N; i += warps_total) {
    const int64_t p0 = P.ptr[i];
    const int n = (int)(P.ptr[i + 1] - p0);
    if (n == 0) continue;                     // no view: out == 0 and nothing flows back
    const int nG = n * G;
    const float* __restrict__ cp = P.compat + p0 * G;
    float* __restrict__ gc = P.gcompat + p0 * G;
    const float m = P.s_max[i * G + gl], den = P.s_den[i * G + gl];
    const int arg_v = P.s_arg[i * G + gl];
    const float sq = P.group_scaling ? sqrtf((float)n) : 1.f;
    const float z = fmaf(gw, m, gb);
    const float t = gating ? tanhf(fmaxf(z, 0.f)) : 1.f;
    float S = 0.f;                            // sum_v a_vg s_vg for g = lane%G (partial per lane)

    for (int vs = 0; vs < n; vs += 32) {
      const int nc = min(32, n - vs);
      __syncwarp();
      for (int e = lane; e < nc * G; e += 32) {
        att_s[e] = expf((__ldg(cp + vs * G + e) - m) / sq) / den;
        s_s[e] = 0.f;
      }
      if (lane < nc) row_s[lane] = load_idx(P.idx, P.idx64, p0 + vs + lane);
      __syncwarp();

      for (int ct = 0; ct < C; ct += TILE_C) {
        float go[CPL][VEC];                   // dO of this lane's channels
        int gk[CPL];                          // group of each chunk (chunks never straddle groups)
        float tk[CPL];                        // gate value of that group
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int c0 = ct + (lir + LPR * k) * VEC;
          gk[k] = group_of_channel(c0 < C ? c0 : C - 1, C, G);
          tk[k] = __shfl_sync(0xffffffffu, t, gk[k]);
          if (c0 < C) {
            load_chunk<T, VEC>(gout + i * (int64_t)C + c0, go[k]);
          } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) go[k][j] = 0.f;
          }
        }

        for (int v0 = 0; v0 < nc; v0 += RPI * U) {
          Chunk<T, VEC> f[U][CPL];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int v = v0 + u * RPI + sg;
            ok[u] = v < nc;
            const T* rp = x + row_s[ok[u] ? v : 0] * (int64_t)C + ct;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const int c0 = (lir + LPR * k) * VEC;
              if (ok[u] && ct + c0 < C) f[u][k].load(rp + c0); else f[u][k].zero();
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int vv = ok[u] ? v0 + u * RPI + sg : 0;
            const int64_t orow = (P.scatter && P.idx != nullptr) ? row_s[vv] : (p0 + vs + vv);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const int c0 = ct + (lir + LPR * k) * VEC;
              const bool live = ok[u] && c0 < C;
              float fv[VEC], dx[VEC];
              f[u][k].get(fv);
              float dot = 0.f;
              const float a_t = att_s[vv * G + gk[k]] * tk[k];
#pragma unroll
              for (int j = 0; j < VEC; ++j) {
                dot = fmaf(go[k][j], fv[j], dot);
                dx[j] = a_t * go[k][j];
              }
              if (live) store_chunk<T, VEC>(gx + orow * (int64_t)C + c0, dx);
              // ---- s_vg += sum over the lanes of this row whose chunk lies in group g
              if (red_mode == 1) {            // groups = aligned blocks of cpg lanes
                float r = dot;
                for (int off = 1; off < cpg; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
                if (live && (lir & (cpg - 1)) == 0) s_s[vv * G + gk[k]] += r;
              } else if (red_mode == 2) {     // the whole row step lies in one group
                float r = dot;
#pragma unroll
                for (int off = 1; off < LPR; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
                if (live && lir == 0) s_s[vv * G + gk[k]] += r;
              } else {                        // irregular group sizes: one reduction per group
                for (int g = 0; g < G; ++g) {
                  const bool mine = live && (gk[k] == g);
                  if (!__any_sync(0xffffffffu, mine)) continue;
                  float r = mine ? dot : 0.f;
#pragma unroll
                  for (int off = 1; off < LPR; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
                  if (ok[u] && lir == 0) s_s[vv * G + g] += r;
                }
              }
            }
          }
        }
        __syncwarp();
      }

      // S partial and raw s -> grad_compat (finalised below once S is complete)
      __syncwarp();
      for (int e = lane; e < nc * G; e += 32) {
        const float s = s_s[e];
        S = fmaf(att_s[e], s, S);
        gc[vs * G + e] = s;
      }
    }

    S = group_lane_sum(S, G);
    const float one_m_t2 = 1.f - t * t;
    const float dq = (gating && z > 0.f) ? S * one_m_t2 * gw : 0.f;
    if (gating && z > 0.f && lane < G) {
      dw_acc += S * one_m_t2 * m;
      db_acc += S * one_m_t2;
    }
    __syncwarp();
    for (int e = lane; e < nG; e += 32) {
      const float a = expf((__ldg(cp + e) - m) / sq) / den;
      const float s = gc[e];
      float d = a * t * (s - S) / sq;
      if (p0 + e / G == arg_v) d += dq;
      gc[e] = d;
    }
  }

  // ---- gate parameter gradients: warp -> block partial (deterministic), block -> workspace
  if (P.gate_partial != nullptr) {
    if (lane < G) {
      gate_s[(warp * 2 + 0) * G + lane] = dw_acc;
      gate_s[(warp * 2 + 1) * G + lane] = db_acc;
    }
    __syncthreads();
    if (threadIdx.x < 2 * G) {
      float acc = 0.f;
      for (int w = 0; w < kWarps; ++w) acc += gate_s[w * 2 * G + threadIdx.x];
      P.gate_partial[(int64_t)blockIdx.x * 2 * G + threadIdx.x] = acc;
    }
  }
}

__global__ void gate_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                   int blocks, int twoG) {
  const int j = threadIdx.x;
  if (j >= twoG) return;
  float acc = 0.f;
  for (int b = 0; b < blocks; ++b) acc += partial[(int64_t)b * twoG + j];
  out[j] = acc;
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
struct VAConfig { int vec, lpr, cpl; };

template <typename T>
static VAConfig choose_config(const VAParams& P, const void* o1, const void* o2) {
  constexpr int V16 = Vec16<T>::N;
  const int C = P.C, G = P.G;
  bool vec_ok = (C % V16 == 0) && aligned16(P.x) && aligned16(o1) && (o2 == nullptr || aligned16(o2));
  if (vec_ok) {  // every 16-byte chunk must sit inside one channel group
    for (int c0 = 0; c0 < C && vec_ok; c0 += V16)
      if (group_of_channel(c0, C, G) != group_of_channel(c0 + V16 - 1, C, G)) vec_ok = false;
  }
  VAConfig cfg;
  if (!vec_ok) { cfg.vec = 1; cfg.lpr = 32; cfg.cpl = (C > 32) ? 4 : 1; return cfg; }
  const int cv = C / V16;
  cfg.vec = V16;
  if (cv <= 4) { cfg.lpr = 4; cfg.cpl = 1; }
  else if (cv <= 8) { cfg.lpr = 8; cfg.cpl = 1; }
  else if (cv <= 16) { cfg.lpr = 16; cfg.cpl = 1; }
  else if (cv <= 32) { cfg.lpr = 32; cfg.cpl = 1; }
  else if (cv <= 64) { cfg.lpr = 32; cfg.cpl = 2; }
  else { cfg.lpr = 32; cfg.cpl = 4; }
  return cfg;
}

static int va_grid(int64_t N) {
  int64_t blocks = (N + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T, int VEC, int LPR, int CPL>
static int launch_fwd(const VAParams& P, cudaStream_t st) {
  const size_t smem = (size_t)kWarps * 32 * P.G * sizeof(float) + (size_t)kWarps * 32 * sizeof(int64_t);
  auto kern = view_attention_fwd_kernel<T, VEC, LPR, CPL>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<va_grid(P.N), kWarps * 32, smem, st>>>(P);
  return check_launch("view_attention_fwd");
}

template <typename T, int VEC, int LPR, int CPL>
static int launch_bwd(const VAParams& P, int grid, cudaStream_t st) {
  const size_t smem = (size_t)2 * kWarps * 32 * P.G * sizeof(float) + (size_t)kWarps * 32 * sizeof(int64_t) +
                      (size_t)kWarps * 2 * P.G * sizeof(float);
  auto kern = view_attention_bwd_kernel<T, VEC, LPR, CPL>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, kWarps * 32, smem, st>>>(P);
  return check_launch("view_attention_bwd");
}

#define DVA_VA_DISPATCH(FN, T, cfg, ...)                                                   \
  do {                                                                                     \
    constexpr int V16 = Vec16<T>::N;                                                       \
    if (cfg.vec == 1) {                                                                    \
      if (cfg.cpl == 1) return FN<T, 1, 32, 1>(__VA_ARGS__);                               \
      return FN<T, 1, 32, 4>(__VA_ARGS__);                                                 \
    }                                                                                      \
    if (cfg.lpr == 4) return FN<T, V16, 4, 1>(__VA_ARGS__);                                \
    if (cfg.lpr == 8) return FN<T, V16, 8, 1>(__VA_ARGS__);                                \
    if (cfg.lpr == 16) return FN<T, V16, 16, 1>(__VA_ARGS__);                              \
    if (cfg.cpl == 1) return FN<T, V16, 32, 1>(__VA_ARGS__);                               \
    if (cfg.cpl == 2) return FN<T, V16, 32, 2>(__VA_ARGS__);                               \
    return FN<T, V16, 32, 4>(__VA_ARGS__);                                                 \
  } while (0)

template <typename T> static int fwd_typed(const VAParams& P, cudaStream_t st) {
  const VAConfig cfg = choose_config<T>(P, P.out, nullptr);
  DVA_VA_DISPATCH(launch_fwd, T, cfg, P, st);
}
template <typename T> static int bwd_typed(const VAParams& P, int grid, cudaStream_t st) {
  const VAConfig cfg = choose_config<T>(P, P.gout, P.gx);
  DVA_VA_DISPATCH(launch_bwd, T, cfg, P, grid, st);
}

static bool pow2_le32(int64_t g) { return g >= 1 && g <= 32 && (g & (g - 1)) == 0; }

}  // namespace dva

using namespace dva;

extern "C" int dva_view_attention_fwd(const void* x, const void* idx, int idx_is_i64,
                                      const float* compat, const int64_t* ptr,
                                      const float* gate_w, const float* gate_b, void* out,
                                      float* att, float* seg_max, float* seg_den,
                                      int32_t* seg_arg, int64_t N, int64_t V, int64_t R,
                                      int64_t C, int64_t G, int group_scaling, float eps,
                                      int dtype, void* stream) {
  if (N < 0 || V < 0 || R < 0 || C < 1 || G < 1) return fail(DVA_EINVAL, "view_attention_fwd: bad sizes");
  if (G > C) return fail(DVA_EINVAL, "view_attention_fwd: num_groups > channels");
  if (!pow2_le32(G)) return fail(DVA_EUNSUPPORTED, "view_attention_fwd: G must be a power of two <= 32");
  if (C > (1 << 20)) return fail(DVA_EUNSUPPORTED, "view_attention_fwd: C too large");
  if (N == 0) return DVA_OK;
  if (!ptr || !out || (V > 0 && (!x || !compat))) return fail(DVA_EINVAL, "view_attention_fwd: null pointer");
  if ((gate_w == nullptr) != (gate_b == nullptr)) return fail(DVA_EINVAL, "view_attention_fwd: gate_w/gate_b must both be given");
  if ((seg_max != nullptr) && (!seg_den || !seg_arg)) return fail(DVA_EINVAL, "view_attention_fwd: seg_max/seg_den/seg_arg go together");
  if (idx == nullptr && R < V) return fail(DVA_EINVAL, "view_attention_fwd: identity idx needs R >= V");
  VAParams P{};
  P.x = x; P.idx = idx; P.idx64 = idx_is_i64; P.compat = compat; P.ptr = ptr;
  P.gate_w = gate_w; P.gate_b = gate_b; P.out = out; P.att = att;
  P.seg_max = seg_max; P.seg_den = seg_den; P.seg_arg = seg_arg;
  P.N = N; P.V = V; P.R = R; P.C = (int)C; P.G = (int)G; P.group_scaling = group_scaling; P.eps = eps;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DVA_F32: return fwd_typed<float>(P, st);
    case DVA_BF16: return fwd_typed<__nv_bfloat16>(P, st);
    case DVA_F16: return fwd_typed<__half>(P, st);
    default: return fail(DVA_EINVAL, "view_attention_fwd: unknown dtype");
  }
}

extern "C" size_t dva_view_attention_bwd_workspace_bytes(int64_t G) {
  return (size_t)kNumSMs * 8 * 2 * (size_t)(G > 0 ? G : 1) * sizeof(float);
}

extern "C" int dva_view_attention_bwd(const void* x, const void* idx, int idx_is_i64,
                                      const float* compat, const int64_t* ptr,
                                      const float* gate_w, const float* gate_b,
                                      const void* grad_out, const float* seg_max,
                                      const float* seg_den, const int32_t* seg_arg,
                                      void* grad_x_rows, float* grad_compat, float* grad_gate,
                                      int scatter_rows, int64_t N, int64_t V, int64_t R,
                                      int64_t C, int64_t G, int group_scaling, int dtype,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || V < 0 || R < 0 || C < 1 || G < 1) return fail(DVA_EINVAL, "view_attention_bwd: bad sizes");
  if (G > C) return fail(DVA_EINVAL, "view_attention_bwd: num_groups > channels");
  if (!pow2_le32(G)) return fail(DVA_EUNSUPPORTED, "view_attention_bwd: G must be a power of two <= 32");
  const bool gating = gate_w != nullptr;
  if (gating != (gate_b != nullptr)) return fail(DVA_EINVAL, "view_attention_bwd: gate_w/gate_b must both be given");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0 || V == 0) {
    if (gating && grad_gate) cudaMemsetAsync(grad_gate, 0, 2 * G * sizeof(float), st);
    return DVA_OK;
  }
  if (!x || !compat || !ptr || !grad_out || !seg_max || !seg_den || !seg_arg || !grad_x_rows || !grad_compat)
    return fail(DVA_EINVAL, "view_attention_bwd: null pointer");
  if (gating && (!grad_gate || !workspace || workspace_bytes < dva_view_attention_bwd_workspace_bytes(G)))
    return fail(DVA_EINVAL, "view_attention_bwd: gating needs grad_gate and workspace");
  VAParams P{};
  P.x = x; P.idx = idx; P.idx64 = idx_is_i64; P.compat = compat; P.ptr = ptr;
  P.gate_w = gate_w; P.gate_b = gate_b; P.gout = grad_out;
  P.s_max = seg_max; P.s_den = seg_den; P.s_arg = seg_arg;
  P.gx = grad_x_rows; P.gcompat = grad_compat; P.scatter = scatter_rows;
  P.gate_partial = gating ? reinterpret_cast<float*>(workspace) : nullptr;
  P.N = N; P.V = V; P.R = R; P.C = (int)C; P.G = (int)G; P.group_scaling = group_scaling;
  const int grid = va_grid(N);
  int rc;
  switch (dtype) {
    case DVA_F32: rc = bwd_typed<float>(P, grid, st); break;
    case DVA_BF16: rc = bwd_typed<__nv_bfloat16>(P, grid, st); break;
    case DVA_F16: rc = bwd_typed<__half>(P, grid, st); break;
    default: return fail(DVA_EINVAL, "view_attention_bwd: unknown dtype");
  }
  if (rc) return rc;
  if (gating) {
    gate_reduce_kernel<<<1, 64, 0, st>>>(P.gate_partial, grad_gate, grid, 2 * (int)G);
    return check_launch("gate_reduce");
  }
  return DVA_OK;
}
