// Skinny projections of the pool MLPs (base_modules.py:42 -- Linear(bias=False) inside every MLP
// layer of DeepSetFeat / E_mod / E_mix, pooling.py:239-261, 645-656): [rows, K] x [N, K]^T with
// K, N <= 64 and millions of rows (one per view).  2*K*N flops against 4*(K+N) bytes per row is
// ~16 flop/byte at K = N = 32: HBM-bound on the fp32 pipes already, and far too narrow for the
// 128x64 tcgen05 tiles of mlp_gemm.cu (measured there: 0.50 ms per launch at 1.28 M x 32 x 32,
// 10x the HBM time).  Exact fp32 FFMA, weights resident in shared memory, every global access a
// coalesced 16-byte vector through a shared staging tile:
//   layouts 0/1  skinny_rows_kernel : D[M,OUT] = A[M,RED] . Wt[RED,OUT]   (one row per thread)
//   layout  2    skinny_dw_kernel   : D[N,K]   = A[M,N]^T . B[M,K]        (4x4 micro-tiles per
//                thread, rows split over thread slices, per-CTA partials reduced in a fixed order)
#include "dva_common.cuh"

namespace dva {

constexpr int kSkTile = 128;       // rows per CTA tile
constexpr int kSkMax = 64;         // largest K / N served here

// rows [row0, row0 + kSkTile) x cols of a row-major [M, cols] matrix -> dst[r * dst_stride + c];
// rows past M become zero.  vec: 16-byte cp.async (LDGSTS) -- every copy of the tile is in flight
// at once and no register waits on it; the caller commits / waits.  Otherwise plain scalar loads.
__device__ __forceinline__ void sk_load_tile(float* __restrict__ dst, int dst_stride, const float* __restrict__ src,
                                             int64_t row0, int64_t M, int cols, bool vec) {
  const int nthreads = blockDim.x;
  if (vec) {
    const int c4 = cols >> 2;
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    if (nthreads % c4 == 0) {                        // the usual case: a thread keeps its column, rows advance by a constant
      const int c = ((int)threadIdx.x % c4) << 2, rstep = nthreads / c4;
      for (int r = threadIdx.x / c4; r < kSkTile; r += rstep) {
        const bool in = row0 + r < M;
        const float* g = src + (in ? (row0 + r) * cols + c : 0);
        const int bytes = in ? 16 : 0;               // src-size 0: the 16 destination bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + (uint32_t)(r * dst_stride + c) * 4), "l"(g), "r"(bytes) : "memory");
      }
    } else {
      for (int e = threadIdx.x; e < kSkTile * c4; e += nthreads) {
        const int r = e / c4, c = (e - r * c4) << 2;
        const bool in = row0 + r < M;
        const float* g = src + (in ? (row0 + r) * cols + c : 0);
        const int bytes = in ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + (uint32_t)(r * dst_stride + c) * 4), "l"(g), "r"(bytes) : "memory");
      }
    }
  } else {
    for (int e = threadIdx.x; e < kSkTile * cols; e += nthreads) {
      const int r = e / cols, c = e - r * cols;
      dst[r * dst_stride + c] = (row0 + r < M) ? __ldg(src + (row0 + r) * cols + c) : 0.f;
    }
  }
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sk_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// D[M,OUT] = A[M,RED] . Wt[RED,OUT].  W is given as [OUT,RED] (TRANS: layout 0, x . W^T) or as
// [RED,OUT] (layout 1, dz . W).  A CTA tile is 128 rows; thread (rg, cg) owns an 8 x 8 register
// tile: rows rg + 16 i (i < 8; the interleave keeps the A-tile LDS.128 conflict-free) x columns
// 8 cg .. 8 cg + 7.  Per k that is 2 + 2 LDS.128 for 64 FFMA -- with one row per thread the
// weights alone cost 8 LDS.128 per 32 FFMA and the kernel ran at the shared-memory bandwidth.
template <bool TRANS>
__global__ void __launch_bounds__(128)
skinny_rows_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D,
                   int64_t M, int RED, int OUT, int avec, int dvec) {
  extern __shared__ __align__(16) float sk_smem[];
  const int RED4 = (RED + 3) & ~3;
  const int REDP = RED4 + 4;                        // tile strides: multiples of 4, /4 odd -> LDS.128 conflict-free
  const int OUTP = (OUT + 31) & ~31;                // weight tile width (zero padded)
  const int OUTS = OUTP + 4;
  float* wS = sk_smem;                               // [RED4][OUTP]
  const int tile_floats = kSkTile * (REDP > OUTS ? REDP : OUTS);
  float* tbuf = wS + RED4 * OUTP;                    // 2 x (A tile [kSkTile][REDP], later the output tile [kSkTile][OUTS])
  for (int e = threadIdx.x; e < RED4 * OUTP; e += blockDim.x) {
    const int k = e / OUTP, n = e - k * OUTP;
    float v = 0.f;
    if (k < RED && n < OUT) v = TRANS ? __ldg(W + (int64_t)n * RED + k) : __ldg(W + (int64_t)k * OUT + n);
    wS[e] = v;
  }
  const int ncg = OUTP >> 3;                         // column groups of 8: 4 or 8
  const int cg = threadIdx.x % ncg, rg = threadIdx.x / ncg;   // rg < 16
  const int64_t tiles = (M + kSkTile - 1) / kSkTile;
  // tile i of this CTA lives in buffer i & 1; tile i + 1 is copied in while tile i is computed
  if (blockIdx.x < tiles) sk_load_tile(tbuf, REDP, A, (int64_t)blockIdx.x * kSkTile, M, RED, avec != 0);
  sk_commit();
  int cur = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, cur ^= 1) {
    const int64_t row0 = t * kSkTile;
    float* tS = tbuf + cur * tile_floats;
    __syncthreads();                                 // the other buffer's output has been stored
    if (t + gridDim.x < tiles)
      sk_load_tile(tbuf + (cur ^ 1) * tile_floats, REDP, A, (t + gridDim.x) * kSkTile, M, RED, avec != 0);
    sk_commit();
    sk_wait<1>();                                    // tile t has landed
    if (RED4 != RED) {                               // zero the padding columns read by the float4 loop
      for (int e = threadIdx.x; e < kSkTile * (RED4 - RED); e += blockDim.x) {
        const int r = e / (RED4 - RED), c = RED + e % (RED4 - RED);
        tS[r * REDP + c] = 0.f;
      }
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* ar = tS + rg * REDP;
    const float* wc = wS + cg * 8;
    for (int k = 0; k < RED4; k += 4) {
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4*>(ar + (16 * i) * REDP + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(wc + (k + kk) * OUTP);
        const float4 w1 = *reinterpret_cast<const float4*>(wc + (k + kk) * OUTP + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = kk == 0 ? a4[i].x : (kk == 1 ? a4[i].y : (kk == 2 ? a4[i].z : a4[i].w));
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();                                 // every thread is done reading the A tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* orow = tS + (rg + 16 * i) * OUTS + cg * 8;
      *reinterpret_cast<float4*>(orow) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(orow + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();
    if (dvec && (int)blockDim.x % (OUT >> 2) == 0) {
      const int c4 = OUT >> 2, c = ((int)threadIdx.x % c4) << 2, rstep = (int)blockDim.x / c4;
      for (int r = threadIdx.x / c4; r < kSkTile; r += rstep)
        if (row0 + r < M)
          *reinterpret_cast<float4*>(D + (row0 + r) * OUT + c) = *reinterpret_cast<const float4*>(tS + r * OUTS + c);
    } else if (dvec) {
      const int c4 = OUT >> 2;
      for (int e = threadIdx.x; e < kSkTile * c4; e += blockDim.x) {
        const int r = e / c4, c = (e - r * c4) << 2;
        if (row0 + r < M)
          *reinterpret_cast<float4*>(D + (row0 + r) * OUT + c) = *reinterpret_cast<const float4*>(tS + r * OUTS + c);
      }
    } else {
      for (int e = threadIdx.x; e < kSkTile * OUT; e += blockDim.x) {
        const int r = e / OUT, c = e - r * OUT;
        if (row0 + r < M) D[(row0 + r) * OUT + c] = tS[r * OUTS + c];
      }
    }
  }
}

// partial[cta][N*K] = sum over the CTA's rows of A[r][n] * B[r][k].  8 x 8 micro-tiles of the
// [N,K] result per thread (2 + 2 LDS.128 for 64 FFMA per row); the rows of a tile are split over
// 128 / (#micro-tiles) thread slices whose partial sums meet in shared memory at the end.
constexpr int kSkDwThreads = 128;
__global__ void __launch_bounds__(kSkDwThreads)
skinny_dw_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ partial,
                 int64_t M, int N, int K, int avec, int bvec) {
  extern __shared__ __align__(16) float sk_smem[];
  const int N8 = (N + 7) & ~7, K8 = (K + 7) & ~7;
  const int NP = N8 + 4, KP = K8 + 4;
  const int pair_floats = kSkTile * (NP + KP);       // one buffer = A tile [kSkTile][NP] + B tile [kSkTile][KP]
  float* rS = sk_smem;                               // cross-slice reduction [threads][64], after the last tile
  const int ntk = K8 >> 3, nt = (N8 >> 3) * ntk;     // 8x8 micro-tiles: at most 64
  const int slices = kSkDwThreads / nt;
  const int mt = threadIdx.x % nt, slice = threadIdx.x / nt;
  const bool worker = slice < slices;
  const int n0 = (mt / ntk) << 3, k0 = (mt % ntk) << 3;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // padding columns stay zero for the whole kernel
  for (int e = threadIdx.x; e < 2 * pair_floats; e += blockDim.x) sk_smem[e] = 0.f;
  __syncthreads();
  const int64_t tiles = (M + kSkTile - 1) / kSkTile;
  if (blockIdx.x < tiles) {
    sk_load_tile(sk_smem, NP, A, (int64_t)blockIdx.x * kSkTile, M, N, avec != 0);
    sk_load_tile(sk_smem + kSkTile * NP, KP, B, (int64_t)blockIdx.x * kSkTile, M, K, bvec != 0);
  }
  sk_commit();
  int cur = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, cur ^= 1) {
    const float* aS = sk_smem + cur * pair_floats;
    const float* bS = aS + kSkTile * NP;
    __syncthreads();                                 // every thread is done with the other buffer
    if (t + gridDim.x < tiles) {
      float* nx = sk_smem + (cur ^ 1) * pair_floats;
      sk_load_tile(nx, NP, A, (t + gridDim.x) * kSkTile, M, N, avec != 0);
      sk_load_tile(nx + kSkTile * NP, KP, B, (t + gridDim.x) * kSkTile, M, K, bvec != 0);
    }
    sk_commit();
    sk_wait<1>();
    __syncthreads();
    if (worker) {
#pragma unroll 2
      for (int r = slice; r < kSkTile; r += slices) {
        const float4 a0 = *reinterpret_cast<const float4*>(aS + r * NP + n0);
        const float4 a1 = *reinterpret_cast<const float4*>(aS + r * NP + n0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bS + r * KP + k0);
        const float4 b1 = *reinterpret_cast<const float4*>(bS + r * KP + k0 + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(av[i], b0.x, acc[i][0]); acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], b0.z, acc[i][2]); acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
          acc[i][4] = fmaf(av[i], b1.x, acc[i][4]); acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
          acc[i][6] = fmaf(av[i], b1.z, acc[i][6]); acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
        }
      }
    }
  }
  __syncthreads();                                   // tiles are dead: their memory becomes rS
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) rS[threadIdx.x * 64 + i * 8 + j] = worker ? acc[i][j] : 0.f;
  __syncthreads();
  // slices summed in a fixed order; micro-tile (n0,k0) element (i,j) -> partial[n0+i][k0+j]
  float* out = partial + (int64_t)blockIdx.x * N * K;
  for (int e = threadIdx.x; e < nt * 64; e += blockDim.x) {
    const int m = e >> 6, ij = e & 63;
    float sum = 0.f;
    for (int sl = 0; sl < slices; ++sl) sum += rS[(sl * nt + m) * 64 + ij];
    const int n = ((m / ntk) << 3) + (ij >> 3), k = ((m % ntk) << 3) + (ij & 7);
    if (n < N && k < K) out[n * K + k] = sum;
  }
}

// D[e] = sum_cta partial[cta][e]: one warp per element, lanes stride over the CTAs (coalesced
// across the warps of a block), fixed-order butterfly at the end -> deterministic
__global__ void __launch_bounds__(256)
skinny_dw_reduce_kernel(const float* __restrict__ partial, float* __restrict__ D, int ctas, int NK) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= NK) return;
  float s = 0.f;
  for (int c = lane; c < ctas; c += 32) s += partial[(int64_t)c * NK + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) D[e] = s;
}

static int sk_dw_grid(int64_t M) {
  const int64_t tiles = (M + kSkTile - 1) / kSkTile;
  const int64_t cap = (int64_t)kNumSMs * 3;    // 74 KB of shared memory per CTA at N = K = 32
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

}  // namespace dva

using namespace dva;

// Served here: both small dimensions <= 64 (any values, multiples of 4 take the vector loads).
extern "C" int dva_skinny_gemm_supported(int64_t M, int64_t N, int64_t K, int layout) {
  (void)layout;
  return M >= 1 && N >= 1 && K >= 1 && N <= kSkMax && K <= kSkMax;
}

extern "C" size_t dva_skinny_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int layout) {
  if (layout != 2) return 16;
  return (size_t)sk_dw_grid(M) * (size_t)N * (size_t)K * sizeof(float) + 16;
}

// layout 0: D[M,N] = A[M,K] . B[N,K]^T;  1: D[M,N] = A[M,K] . B[K,N];  2: D[N,K] = A[M,N]^T . B[M,K]
extern "C" int dva_skinny_gemm(const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t K,
                               int layout, void* workspace, size_t workspace_bytes, void* stream) {
  if (M == 0) return DVA_OK;
  if (!dva_skinny_gemm_supported(M, N, K, layout)) return fail(DVA_EUNSUPPORTED, "skinny_gemm: N and K must be <= 64");
  if (!A || !B || !D) return fail(DVA_EINVAL, "skinny_gemm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == 0 || layout == 1) {
    const int RED = (int)K, OUT = (int)N;
    const int RED4 = (RED + 3) & ~3, REDP = RED4 + 4, OUTP = (OUT + 31) & ~31, OUTS = OUTP + 4;
    const int tile_floats = kSkTile * (REDP > OUTS ? REDP : OUTS);
    const size_t smem = (size_t)(RED4 * OUTP + 2 * tile_floats) * sizeof(float);
    const int threads = 16 * (OUTP / 8);             // 64 (OUT <= 32) or 128
    const int avec = (RED % 4 == 0) && aligned16(A), dvec = (OUT % 4 == 0) && aligned16(D);
    const int64_t tiles = (M + kSkTile - 1) / kSkTile;
    const int64_t cap = (int64_t)kNumSMs * 5;    // ~41 KB per CTA at K = N = 32
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (layout == 0) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(skinny_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      skinny_rows_kernel<true><<<grid, threads, smem, st>>>(A, B, D, M, RED, OUT, avec, dvec);
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(skinny_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      skinny_rows_kernel<false><<<grid, threads, smem, st>>>(A, B, D, M, RED, OUT, avec, dvec);
    }
    return check_launch("skinny_gemm(rows)");
  }
  if (layout != 2) return fail(DVA_EINVAL, "skinny_gemm: bad layout");
  const int grid = sk_dw_grid(M);
  if (!workspace || workspace_bytes < (size_t)grid * N * K * sizeof(float))
    return fail(DVA_EINVAL, "skinny_gemm: workspace too small");
  const int N8 = ((int)N + 7) & ~7, K8 = ((int)K + 7) & ~7;
  size_t smem = (size_t)2 * (kSkTile * (N8 + 4) + kSkTile * (K8 + 4)) * sizeof(float);
  if (smem < (size_t)kSkDwThreads * 64 * sizeof(float)) smem = (size_t)kSkDwThreads * 64 * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(skinny_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int avec = (N % 4 == 0) && aligned16(A), bvec = (K % 4 == 0) && aligned16(B);
  float* partial = reinterpret_cast<float*>(workspace);
  skinny_dw_kernel<<<grid, kSkDwThreads, smem, st>>>(A, B, partial, M, (int)N, (int)K, avec, bvec);
  if (int rc = check_launch("skinny_gemm(dw)")) return rc;
  const int NK = (int)(N * K);
  skinny_dw_reduce_kernel<<<(NK + 7) / 8, 256, 0, st>>>(partial, D, grid, NK);
  return check_launch("skinny_gemm(dw reduce)");
}
