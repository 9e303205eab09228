// st_gemm_bf.cu — bf16-operand instantiation of the tcgen05 GEMM (kind::f16, fp32 accumulate).
#include "st_gemm_impl.cuh"

namespace st {
int gemm_bf16(cudaStream_t stream, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
              int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits) {
  return gemm_run<__nv_bfloat16, false>(stream, mode, A, lda, B, ldb, C, ldc, c_lp, M, N, K, ep, k_splits);
}
}  // namespace st
