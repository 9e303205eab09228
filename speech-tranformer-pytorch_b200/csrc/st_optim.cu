// st_optim.cu — flat-buffer gradient norm and fused clip + Adam update.
//
// Reference: train.py:45-46 (`clip_grad_norm_` then `optimizer.step`) with the Adam settings of
// transformer/Optim.py:9-14 (betas (0.9, 0.98), eps 1e-9) and the Noam learning rate computed on the
// host (Optim.py:36-45).  All parameters / gradients / moments live in one contiguous buffer each,
// so the whole update is two HBM-bound launches with no host synchronisation: the clip coefficient
// is derived on the device from the squared norm left in `norm_ws` by st_sumsq.
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {
namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = x4[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

struct AdamParams {
  float* p; const float* g; float* m; float* v;
  int64_t n;
  float lr, b1, b2, eps, bc1, bc2_rsqrt, max_norm, gscale;
  const float* sumsq;
  void* p_twin;    // optional: operand-precision copy of the updated parameters (what the next step's GEMMs consume)
  int twin_dt;     // ST_DTYPE_F32: TF32-rounded fp32; ST_DTYPE_F16 / ST_DTYPE_BF16: 16-bit
};

__global__ void __launch_bounds__(256)
adam_kernel(const AdamParams a) {
  // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
  float coef = a.gscale;
  // a non-finite gradient norm (an overflowed fp16 activation gradient under loss scaling) skips the update altogether:
  // parameters, moments and twins stay as they are (the caller reads norm_ws to notice and lowers its loss scale)
  if (a.sumsq && !isfinite(*a.sumsq)) return;
  if (a.sumsq && a.max_norm > 0.f) {
    const float total = sqrtf(*a.sumsq) * a.gscale;
    coef *= fminf(a.max_norm / (total + 1e-6f), 1.f);
  }
  const float step_size = a.lr / a.bc1;
  const int64_t n4 = a.n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = reinterpret_cast<float*>(&p);
    const float* gg = reinterpret_cast<const float*>(&g4);
    float* mm = reinterpret_cast<float*>(&m);
    float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float g = gg[t] * coef;
      mm[t] = a.b1 * mm[t] + (1.f - a.b1) * g;
      vv[t] = a.b2 * vv[t] + (1.f - a.b2) * g * g;
      pp[t] -= step_size * mm[t] / (sqrtf(vv[t]) * a.bc2_rsqrt + a.eps);
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    if (a.p_twin) {
      if (a.twin_dt == ST_DTYPE_F32)
        reinterpret_cast<float4*>(a.p_twin)[i] = make_float4(tf32_rna(pp[0]), tf32_rna(pp[1]), tf32_rna(pp[2]), tf32_rna(pp[3]));
      else if (a.twin_dt == ST_DTYPE_F16)
        reinterpret_cast<uint2*>(a.p_twin)[i] = make_uint2(pack2<__half>(pp[0], pp[1]), pack2<__half>(pp[2], pp[3]));
      else
        reinterpret_cast<uint2*>(a.p_twin)[i] = make_uint2(pack2<__nv_bfloat16>(pp[0], pp[1]), pack2<__nv_bfloat16>(pp[2], pp[3]));
    }
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
}

}  // namespace

int sumsq_add(cudaStream_t s, const float* x, int64_t n, float* out) {
  if (n == 0) return ST_OK;
  ST_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: pointer must be 16-byte aligned");
  const int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(s, PROF_SUMSQ, 4.0 * n);
  sumsq_kernel<<<static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap), 256, 0, s>>>(x, n, out);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
              float eps, int step, float max_norm, float gscale, const float* sumsq, void* p_twin, int twin_dt) {
  if (n == 0) return ST_OK;
  ST_REQUIRE((n & 3) == 0, "adam_step: flat buffer length must be a multiple of 4 (pad it)");
  ST_REQUIRE(step >= 1, "adam_step: step must be >= 1");
  AdamParams a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.n = n;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), step));
  a.bc2_rsqrt = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), step)));
  a.max_norm = max_norm; a.gscale = gscale; a.sumsq = sumsq; a.p_twin = p_twin; a.twin_dt = twin_dt;
  const int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(s, PROF_ADAM, (p_twin ? (twin_dt == ST_DTYPE_F32 ? 8.0 : 7.5) : 7.0) * 4.0 * n);  // read p,g,m,v + write p,m,v (+ TF32 copy)
  adam_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, s>>>(a);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
