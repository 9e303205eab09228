// st_gemm.cu — TF32 instantiation of the tcgen05 GEMM (st_gemm_impl.cuh) and the dispatcher over element types.
#include "st_gemm_impl.cuh"

namespace st {

int gemm_f16(cudaStream_t stream, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
             int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits);    // st_gemm_h.cu
int gemm_bf16(cudaStream_t stream, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
              int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits);   // st_gemm_bf.cu

int gemm_tf32(cudaStream_t stream, GemmMode mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
              int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, int k_splits) {
  return gemm_run<float, true>(stream, mode, A, lda, B, ldb, C, ldc, 0, M, N, K, ep, k_splits);
}

int gemm_any(cudaStream_t stream, int dtype, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
             int64_t ldc, int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits) {
  switch (dtype) {
    case ST_DTYPE_F32: return gemm_run<float, true>(stream, mode, A, lda, B, ldb, C, ldc, 0, M, N, K, ep, k_splits);
    case ST_DTYPE_F16: return gemm_f16(stream, mode, A, lda, B, ldb, C, ldc, c_lp, M, N, K, ep, k_splits);
    case ST_DTYPE_BF16: return gemm_bf16(stream, mode, A, lda, B, ldb, C, ldc, c_lp, M, N, K, ep, k_splits);
  }
  set_error("gemm: bad dtype %d", dtype);
  return ST_ERR_INVALID;
}

}  // namespace st
