// C ABI of the tensor-core projection GEMMs (variants compiled from mlp_gemm.cu).
#include "dva_common.cuh"

extern "C" {
size_t dva_gemm_ws_0_0(int, int, int); size_t dva_gemm_ws_0_1(int, int, int);
size_t dva_gemm_ws_1_0(int, int, int); size_t dva_gemm_ws_1_1(int, int, int);
size_t dva_gemm_ws_2_0(int, int, int); size_t dva_gemm_ws_2_1(int, int, int);
int dva_gemm_run_2_0(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int dva_gemm_run_2_1(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int dva_gemm_run_0_0(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int dva_gemm_run_0_1(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int dva_gemm_run_1_0(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int dva_gemm_run_1_1(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
}

extern "C" int dva_skinny_gemm_supported(int64_t M, int64_t N, int64_t K, int layout);
extern "C" size_t dva_skinny_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int layout);
extern "C" int dva_skinny_gemm(const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t K,
                               int layout, void* workspace, size_t workspace_bytes, void* stream);

using namespace dva;

static bool gemm_shape_ok(int64_t M, int64_t N, int64_t K) {
  return M >= 1 && N >= 4 && K >= 4 && N % 4 == 0 && K % 4 == 0 && M < (1ll << 31) && N <= 65536 && K <= 65536;
}

extern "C" size_t dva_linear_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int layout, int precision) {
  // narrow projections (both small dimensions <= 64: every MLP of the map encoders) run on the
  // exact-fp32 skinny kernels of skinny_gemm.cu whatever the precision mode
  if (dva_skinny_gemm_supported(M, N, K, layout)) return dva_skinny_gemm_workspace_bytes(M, N, K, layout);
  if (!gemm_shape_ok(M, N, K)) return 0;
  size_t w = 0;
  if (layout == 0) w = precision == 0 ? dva_gemm_ws_0_0((int)M, (int)N, (int)K) : dva_gemm_ws_0_1((int)M, (int)N, (int)K);
  else if (layout == 1) w = precision == 0 ? dva_gemm_ws_1_0((int)M, (int)N, (int)K) : dva_gemm_ws_1_1((int)M, (int)N, (int)K);
  else w = precision == 0 ? dva_gemm_ws_2_0((int)M, (int)N, (int)K) : dva_gemm_ws_2_1((int)M, (int)N, (int)K);
  return w < 16 ? 16 : w;
}

extern "C" int dva_linear_gemm(const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t K,
                               int layout, int precision, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (M == 0) return DVA_OK;
  if (layout >= 0 && layout <= 2 && dva_skinny_gemm_supported(M, N, K, layout))
    return dva_skinny_gemm(A, B, D, M, N, K, layout, workspace, workspace_bytes, stream);
  if (!gemm_shape_ok(M, N, K)) return fail(DVA_EUNSUPPORTED, "linear_gemm: N and K must be multiples of 4 (16-byte TMA rows)");
  if (layout < 0 || layout > 2 || (precision != 0 && precision != 1)) return fail(DVA_EINVAL, "linear_gemm: bad layout/precision");
  if (!A || !B || !D) return fail(DVA_EINVAL, "linear_gemm: null pointer");
  if (!aligned16(A) || !aligned16(B) || !aligned16(D)) return fail(DVA_EALIGN, "linear_gemm: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (layout == 0) rc = precision == 0 ? dva_gemm_run_0_0(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st)
                                       : dva_gemm_run_0_1(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st);
  else if (layout == 1) rc = precision == 0 ? dva_gemm_run_1_0(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st)
                                            : dva_gemm_run_1_1(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st);
  else rc = precision == 0 ? dva_gemm_run_2_0(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st)
                           : dva_gemm_run_2_1(A, B, D, (int)M, (int)N, (int)K, workspace, workspace_bytes, st);
  if (rc == -3) return fail(DVA_EUNSUPPORTED, "linear_gemm: shape not supported by the tcgen05 kernel");
  if (rc == -1) return fail(DVA_EINVAL, "linear_gemm: workspace too small");
  if (rc != 0) return fail(rc, "linear_gemm: launch failed");
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  return DVA_OK;
}
