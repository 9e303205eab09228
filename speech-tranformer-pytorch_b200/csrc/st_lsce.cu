// st_lsce.cu — one-pass label-smoothed / soft-target cross-entropy, forward + gradient.
//
// Reference: transformer/Loss.py:13-39 (LabelSmoothingLoss builds a dense smoothed target q from
// `one_hot`, the target column and the padding rows) and Loss.py:44-73 (CrossEntropyLoss:
// loss = sum_i sum_c w_c q_ic (-log_softmax(x_i)_c) / Z).  The reference materialises three N x V
// temporaries; here q is implicit, a row never leaves the SM between the log-sum-exp and the
// gradient, and HBM sees one read of the logits and one write of the gradient.
//
//   loss_i   = -sum_c w_c q_ic (x_ic - lse_i)
//   dL/dx_ic = inv_z * ( softmax(x_i)_c * sum_c' w_c' q_ic' - w_c q_ic )
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

constexpr int CE_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < CE_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <bool DENSE>
__global__ void __launch_bounds__(CE_THREADS)
lsce_kernel(const LsceArgs a) {
  __shared__ float red[CE_THREADS / 32];
  const int64_t row = blockIdx.x;
  const float* x = a.logits + row * a.ldl;
  const float* qd = DENSE ? a.q_dense + row * static_cast<int64_t>(a.V) : nullptr;
  int64_t tgt = -1;
  bool zero_row = false;
  if (!DENSE) {
    tgt = a.target[row];
    zero_row = (a.padding_idx >= 0 && tgt == a.padding_idx);
  }
  float* g = a.grad ? a.grad + row * a.ldg : nullptr;
  if (zero_row) {  // q_i == 0: no loss, no gradient (Loss.py:35-37)
    if (g)
      for (int c = threadIdx.x; c < a.V; c += CE_THREADS) g[c] = 0.f;
    if (threadIdx.x == 0) a.row_loss[row] = 0.f;
    return;
  }
  auto q_at = [&](int c) -> float {
    if (DENSE) return qd[c];
    return (c == tgt) ? a.confidence : a.one_hot[c];
  };

  // pass 1: row max
  float m = -INFINITY;
  for (int c = threadIdx.x; c < a.V; c += CE_THREADS) m = fmaxf(m, x[c]);
  m = block_reduce(m, red, true);
  // pass 2 (row now L1/L2 resident): sum exp, sum w q, sum w q x
  float se = 0.f, wq = 0.f, wqx = 0.f;
  for (int c = threadIdx.x; c < a.V; c += CE_THREADS) {
    const float xv = x[c];
    se += __expf(xv - m);
    const float t = a.weight[c] * q_at(c);
    wq += t;
    // q == 0 must contribute exactly 0 even when x is -inf (0 * -inf would be NaN in the reference
    // as well; keep the reference behaviour by multiplying only when t != 0 is NOT done: mirror it).
    wqx += t * xv;
  }
  se = block_reduce(se, red, false);
  wq = block_reduce(wq, red, false);
  wqx = block_reduce(wqx, red, false);
  const float lse = m + logf(se);
  if (threadIdx.x == 0) a.row_loss[row] = wq * lse - wqx;  // -sum w q (x - lse)
  // pass 3: gradient
  if (g) {
    const float inv_se = 1.f / se;
    for (int c = threadIdx.x; c < a.V; c += CE_THREADS) {
      const float p = __expf(x[c] - m) * inv_se;
      g[c] = a.inv_z * (p * wq - a.weight[c] * q_at(c));
    }
  }
}

// deterministic final reduction of the per-row losses (double accumulation, fixed order)
__global__ void __launch_bounds__(CE_THREADS)
lsce_reduce_kernel(const float* __restrict__ row_loss, int64_t n, float inv_z, float* __restrict__ loss) {
  __shared__ double red[CE_THREADS];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += CE_THREADS) s += static_cast<double>(row_loss[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = CE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = static_cast<float>(red[0] * static_cast<double>(inv_z));
}

}  // namespace

int lsce_fwd_bwd(cudaStream_t stream, const LsceArgs& a) {
  ST_REQUIRE(a.N >= 0 && a.V > 0, "lsce: bad shape N=%lld V=%d", (long long)a.N, a.V);
  ST_REQUIRE(a.weight != nullptr, "lsce: class weight vector is required (Loss.py:59)");
  ST_REQUIRE((a.target != nullptr) != (a.q_dense != nullptr), "lsce: exactly one of target / q_dense must be given");
  // algorithmic bytes: read logits once + write the gradient once
  ProfScope prof(stream, PROF_LSCE, (a.grad ? 2.0 : 1.0) * a.N * a.V * 4 + (a.q_dense ? 1.0 * a.N * a.V * 4 : 8.0 * a.N));
  if (a.N > 0) {
    if (a.q_dense)
      lsce_kernel<true><<<static_cast<unsigned>(a.N), CE_THREADS, 0, stream>>>(a);
    else
      lsce_kernel<false><<<static_cast<unsigned>(a.N), CE_THREADS, 0, stream>>>(a);
    ST_CHECK_LAUNCH();
  }
  lsce_reduce_kernel<<<1, CE_THREADS, 0, stream>>>(a.row_loss, a.N, a.inv_z, a.loss);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
