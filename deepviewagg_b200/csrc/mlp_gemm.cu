// Dense projection GEMMs of the pool MLPs (E_mod / E_map / E_mix / E_main; reference
// core/common_modules/base_modules.py:38-48 `nn.Linear(bias=False)`) on the 5th-generation tensor
// cores: tcgen05.mma (UTCHMMA) fed by TMA (UTMALDG) with TMEM accumulators, assembled from the
// CUTLASS 4.5 sm100 collectives inside this translation unit (header-only templates, no library
// call).  These are the only GEMM-shaped operations on the path (north_star: "tensor cores used
// only on the dense projection GEMMs"); M = views (millions), N, K = channels (<= 512), so the
// kernels are HBM-bound: 4(MK + MN) bytes per launch.
//
//   layout 0 ("TN", forward):   D[M,N] = A[M,K] . W[N,K]^T
//   layout 1 ("NN", backward):  D[M,N] = A[M,K] . B[K,N]          (dX = dZ . W)
//   layout 2 ("TN-reduce"):     D[N,K] = A[M,N]^T . B[M,K]        (dW = dZ^T . X; the contraction runs
//                               over the M rows -> stream-K scheduler splits it across all SMs)
//   precision 0: fast-FP32 = 9 x BF16 split products (fp32-grade accuracy; parity mode)
//   precision 1: TF32 (10-bit mantissa inputs, fp32 accumulate)
// This file is compiled once per (layout, precision) with -DDVA_GEMM_LAYOUT / -DDVA_GEMM_PREC.
#include <cuda_runtime.h>
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/util/packed_stride.hpp"

#ifndef DVA_GEMM_LAYOUT
#define DVA_GEMM_LAYOUT 0
#endif
#ifndef DVA_GEMM_PREC
#define DVA_GEMM_PREC 0
#endif

namespace {
using namespace cute;
using ElementA = float;
#if DVA_GEMM_LAYOUT == 2
using LayoutA = cutlass::layout::ColumnMajor;   // dZ [M,N] row-major == A [N,M] column-major
#else
using LayoutA = cutlass::layout::RowMajor;
#endif
using ElementB = float;
#if DVA_GEMM_LAYOUT == 0
using LayoutB = cutlass::layout::ColumnMajor;   // W [N,K] row-major == B [K,N] column-major
#else
using LayoutB = cutlass::layout::RowMajor;      // B [K,N] row-major
#endif
using ElementC = float; using LayoutC = cutlass::layout::RowMajor;
using ElementAcc = float;
constexpr int kAlign = 4;                       // 16-byte TMA rows
using MmaTileShape = Shape<_128, _64, _32>;
using ClusterShape = Shape<_1, _1, _1>;
#if DVA_GEMM_PREC == 0
using MainSchedule = cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32SmemSm100;
#else
using MainSchedule = cutlass::gemm::KernelTmaWarpSpecialized1SmSm100;
#endif
using EpiSchedule = cutlass::epilogue::TmaWarpSpecialized1Sm;

using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, kAlign,
    ElementC, LayoutC, kAlign, EpiSchedule>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlign, ElementB, LayoutB, kAlign,
    ElementAcc, MmaTileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    MainSchedule>::CollectiveOp;
#if DVA_GEMM_LAYOUT == 2
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue,
                                                        cutlass::gemm::StreamKScheduler>;
#else
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue>;
#endif
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

// (M, N, K) are the dimensions of the ABI call; layout 2 maps them to the GEMM problem (m=N, n=K, k=M)
typename Gemm::Arguments make_args(const float* A, const float* B, float* D, int M_, int N_, int K_) {
#if DVA_GEMM_LAYOUT == 2
  const int M = N_, N = K_, K = M_;
#else
  const int M = M_, N = N_, K = K_;
#endif
  using StrideA = typename Gemm::GemmKernel::StrideA; using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC; using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, cute::make_shape(M, N, 1));
  return typename Gemm::Arguments{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {A, sa, B, sb},
                                  {{1.f, 0.f}, D, sc, D, sd}};
}
}  // namespace

#define DVA_CAT2(a, b, c) a##b##_##c
#define DVA_CAT(a, b, c) DVA_CAT2(a, b, c)
#define DVA_FN(name) DVA_CAT(name, DVA_GEMM_LAYOUT, DVA_GEMM_PREC)

// internal entry points (one pair per compiled variant); the C ABI lives in linear_api.cu
extern "C" size_t DVA_FN(dva_gemm_ws_)(int M, int N, int K) {
  auto args = make_args(nullptr, nullptr, nullptr, M, N, K);
  return Gemm::get_workspace_size(args);
}

extern "C" int DVA_FN(dva_gemm_run_)(const float* A, const float* B, float* D, int M, int N, int K, void* ws,
                                     size_t ws_bytes, cudaStream_t st) {
  auto args = make_args(A, B, D, M, N, K);
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -3;
  if (Gemm::get_workspace_size(args) > ws_bytes) return -1;
  if (gemm.initialize(args, ws, st) != cutlass::Status::kSuccess) return 1000;
  return gemm.run(st) == cutlass::Status::kSuccess ? 0 : 1001;
}
