// st_gemm.cuh — interface of the TF32 tcgen05 GEMM core used by every projection on the path
// (reference: the nn.Linear calls at transformer/Attention.py:74-76,92 and SubLayers.py:25-26).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/st_b200.h"

namespace st {

// Operand layouts (all row-major fp32 in HBM):
//   GEMM_NT  C[M,N] = A[M,K] * B[N,K]^T      forward linear      (A K-major,  B K-major)
//   GEMM_NN  C[M,N] = A[M,K] * B[K,N]        data gradient       (A K-major,  B MN-major)
//   GEMM_TN  C[M,N] = A[K,M]^T * B[K,N]      weight gradient     (A MN-major, B MN-major)
enum GemmMode : int { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

struct GemmEpilogue {
  const float* bias = nullptr;  // [N], added to every row
  const void* aux = nullptr;    // [M, ldaux], fp32 for the TF32 GEMM, the operand type for 16-bit GEMMs
  int64_t ldaux = 0;
  int aux_mode = 0;    // 0: none   1: out += aux (residual)   2: out = aux > 0 ? out * aux_scale : 0 (ReLU backward)
  float aux_scale = 1.f;
  int relu = 0;        // out = max(out, 0) (after bias)
  int round_tf32 = 0;  // round the stored value to TF32 (round-to-nearest) — the value feeds another MMA
  int atomic = 0;      // accumulate into C with red.global.add (split-K partial sums)
  // inverted dropout applied after ReLU (SubLayers.py:25); keep iff hash(seed, row*N+col) >= thresh
  uint32_t drop_thresh = 0;
  float drop_scale = 1.f;
  uint64_t drop_seed = 0;
  // optional [N]: column sums of the stored values are ACCUMULATED here (caller zeroes) — the bias gradient of the
  // layer whose input gradient this GEMM produces, for free instead of a separate pass over the output
  float* colsum = nullptr;
  // optional device scalar (mixed mode, st_common.cuh grad_scale_from_amax): the operands carry a power-of-two gradient
  // scale S derived from *unscale_amax; fp32 results that leave the operator (split-K weight gradients, the fp32 input
  // gradient of the AUX_ADD epilogue, the column sums) are multiplied by 1/S.  16-bit outputs stay scaled.
  const float* unscale_amax = nullptr;
  // optional device scalar (16-bit operands, fp32 output, plain / residual-add epilogue): max|stored value| is accumulated
  // here with atomicMax on the bit pattern (caller clears it) — the next backward operator's unscale_amax, measured for free
  float* amax_out = nullptr;
};

// k_splits > 1 requires ep.atomic and a zero-initialised C.
int gemm_tf32(cudaStream_t stream, GemmMode mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
              int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, int k_splits = 1);

// Any operand type (ST_DTYPE_F32 = the TF32 GEMM above; ST_DTYPE_F16 / ST_DTYPE_BF16 = kind::f16 with fp32 accumulate).
// A, B and ep.aux are of the operand type; C is fp32, or the operand type when c_lp != 0.  Leading dimensions in elements
// (multiples of 16 bytes).  ep.round_tf32 is meaningless for 16-bit operands (the output conversion rounds).
int gemm_any(cudaStream_t stream, int dtype, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
             int64_t ldc, int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits = 1);

}  // namespace st
