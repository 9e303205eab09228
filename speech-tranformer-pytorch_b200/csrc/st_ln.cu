// st_ln.cu — HBM-bound row kernels: residual + LayerNorm forward/backward
// (reference: `self.layernorm(output + v)` transformer/Attention.py:94 and
// `self.dropout2(self.layernorm(inputs + ffn_output))` transformer/SubLayers.py:27),
// TF32 rounding copies and column sums (bias gradients).
//
// One warp owns one row; every lane keeps its slice of the row in registers (float4 x VPL), so each
// element is read once and written once.  Grids are persistent (a multiple of the SM count) and
// stride over rows; per-column reductions stay in registers until the end of the kernel.
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------- forward
template <int VPL>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ out, float* __restrict__ z_out,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int d, float eps,
                  int round_out, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed,
                  const float* __restrict__ post, int64_t post_rows) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);

  float4 g[VPL], bt[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) { g[i] = ld4(gamma + c); bt[i] = ld4(beta + c); }
  }

  // The next row's loads are issued before the current row is reduced: one row (2 KB at d = 512) in flight per warp
  // leaves the kernel latency-bound at ~60 % of the HBM rate with the 16 warps per SM its registers allow.
  const int64_t stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  float4 xn[VPL];
  if (row_first < rows) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) xn[i] = ld4(a + row_first * d + c);
    }
  }
  for (int64_t row = row_first; row < rows; row += stride) {
    const float* br = b ? b + row * d : nullptr;
    float4 x[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) x[i] = xn[i];
    if (row + stride < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) xn[i] = ld4(a + (row + stride) * d + c);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        if (br) {
          const float4 y = ld4(br + c);
          x[i].x += y.x; x[i].y += y.y; x[i].z += y.z; x[i].w += y.w;
        }
        s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float d0 = x[i].x - mean, d1 = x[i].y - mean, d2 = x[i].z - mean, d3 = x[i].w - mean;
        v += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    const float rstd = rsqrtf(warp_sum(v) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        if (z_out) st4(z_out + row * d + c, x[i]);
        float o[4] = {(x[i].x - mean) * rstd * g[i].x + bt[i].x, (x[i].y - mean) * rstd * g[i].y + bt[i].y,
                      (x[i].z - mean) * rstd * g[i].z + bt[i].z, (x[i].w - mean) * rstd * g[i].w + bt[i].w};
        if (drop_thresh) {
          const uint32_t key = dropout_row_key(drop_seed, static_cast<uint64_t>(row));
#pragma unroll
          for (int t = 0; t < 4; t += 2) {
            const uint32_t bits = dropout_pair(key, c + t);
            o[t] = dropout_keep(bits, 0, drop_thresh) ? o[t] * drop_scale : 0.f;
            o[t + 1] = dropout_keep(bits, 1, drop_thresh) ? o[t + 1] * drop_scale : 0.f;
          }
        }
        if (post) {   // + positional encoding row (row mod post_rows), Models.py:42-44
          const float4 pv = ld4(post + (row % post_rows) * d + c);
          o[0] += pv.x; o[1] += pv.y; o[2] += pv.z; o[3] += pv.w;
        }
        if (round_out) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tf32_rna(o[t]);
        }
        st4(out + row * d + c, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }
}

// ---------------------------------------------------------------- backward
// dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat;  dbeta += dy;
// dzsum += dz (the bias gradient of the linear layer that produced z).
template <int VPL>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean_in,
                  const float* __restrict__ rstd_in, const float* __restrict__ gamma, float* __restrict__ dz,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dzsum, int64_t rows,
                  int d, int round_out, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed,
                  const float* __restrict__ gate, float gate_scale) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  __shared__ float red[LN_WARPS][32 * 4 + 4];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);

  float4 g[VPL];
  float4 acc_g[VPL], acc_b[VPL], acc_z[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    g[i] = (c < d) ? ld4(gamma + c) : make_float4(0, 0, 0, 0);
    acc_g[i] = acc_b[i] = acc_z[i] = make_float4(0, 0, 0, 0);
  }

  // next row's dy / z / statistics are in flight while the current row is reduced (see the forward kernel)
  const int64_t stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  float4 dyn[VPL], zn[VPL];
  float mean_n = 0.f, rstd_n = 0.f;
  if (row_first < rows) {
    mean_n = mean_in[row_first]; rstd_n = rstd_in[row_first];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) { dyn[i] = ld4(dy + row_first * d + c); zn[i] = ld4(z + row_first * d + c); }
    }
  }
  for (int64_t row = row_first; row < rows; row += stride) {
    const float mean = mean_n, rstd = rstd_n;
    float4 dyc[VPL], zc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) { dyc[i] = dyn[i]; zc[i] = zn[i]; }
    if (row + stride < rows) {
      mean_n = mean_in[row + stride]; rstd_n = rstd_in[row + stride];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { dyn[i] = ld4(dy + (row + stride) * d + c); zn[i] = ld4(z + (row + stride) * d + c); }
      }
    }
    float4 xh[VPL], gy[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float4 dyv = dyc[i];
        if (drop_thresh) {
          float* e = reinterpret_cast<float*>(&dyv);
          const uint32_t key = dropout_row_key(drop_seed, static_cast<uint64_t>(row));
#pragma unroll
          for (int t = 0; t < 4; t += 2) {
            const uint32_t bits = dropout_pair(key, c + t);
            e[t] = dropout_keep(bits, 0, drop_thresh) ? e[t] * drop_scale : 0.f;
            e[t + 1] = dropout_keep(bits, 1, drop_thresh) ? e[t + 1] * drop_scale : 0.f;
          }
        }
        const float4 zv = zc[i];
        xh[i] = make_float4((zv.x - mean) * rstd, (zv.y - mean) * rstd, (zv.z - mean) * rstd, (zv.w - mean) * rstd);
        gy[i] = make_float4(dyv.x * g[i].x, dyv.y * g[i].y, dyv.z * g[i].z, dyv.w * g[i].w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
        acc_g[i].x += dyv.x * xh[i].x; acc_g[i].y += dyv.y * xh[i].y;
        acc_g[i].z += dyv.z * xh[i].z; acc_g[i].w += dyv.w * xh[i].w;
        acc_b[i].x += dyv.x; acc_b[i].y += dyv.y; acc_b[i].z += dyv.z; acc_b[i].w += dyv.w;
      }
    }
    const float c1 = warp_sum(s1) * inv_d;
    const float c2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float o[4] = {rstd * (gy[i].x - c1 - xh[i].x * c2), rstd * (gy[i].y - c1 - xh[i].y * c2),
                      rstd * (gy[i].z - c1 - xh[i].z * c2), rstd * (gy[i].w - c1 - xh[i].w * c2)};
        if (gate) {   // gradient through dropout(relu(.)) whose output is `gate`: pass (scaled) where it was kept and positive
          const float4 gv = ld4(gate + row * d + c);
          o[0] = gv.x > 0.f ? o[0] * gate_scale : 0.f; o[1] = gv.y > 0.f ? o[1] * gate_scale : 0.f;
          o[2] = gv.z > 0.f ? o[2] * gate_scale : 0.f; o[3] = gv.w > 0.f ? o[3] * gate_scale : 0.f;
        }
        acc_z[i].x += o[0]; acc_z[i].y += o[1]; acc_z[i].z += o[2]; acc_z[i].w += o[3];
        if (round_out) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tf32_rna(o[t]);
        }
        st4(dz + row * d + c, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }

  // block-level column reduction, then one atomic per column per block
  auto reduce_store = [&](float4 (&acc)[VPL], float* dst) {
    if (!dst) return;  // uniform across the block
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      __syncthreads();
      red[warp][lane * 4 + 0] = acc[i].x; red[warp][lane * 4 + 1] = acc[i].y;
      red[warp][lane * 4 + 2] = acc[i].z; red[warp][lane * 4 + 3] = acc[i].w;
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[w][threadIdx.x];
        const int c = i * 128 + threadIdx.x;
        if (c < d) atomicAdd(dst + c, s);
      }
    }
  };
  reduce_store(acc_g, dgamma);
  reduce_store(acc_b, dbeta);
  reduce_store(acc_z, dzsum);
}

// ---------------------------------------------------------------- TF32 rounding copy (2-D, strided)
__global__ void __launch_bounds__(256)
round_tf32_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows,
                  int cols4) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  const int64_t total = rows * cols4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    float4 v = ld4(src + r * lds + c);
    v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
    st4(dst + r * ldd + c, v);
  }
}

// scalar variant for ragged widths / unaligned leading dimensions (vocabulary-sized logits gradients, V = 4337)
__global__ void __launch_bounds__(256)
round_tf32_scalar_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows,
                         int cols) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    dst[r * ldd + c] = tf32_rna(src[r * lds + c]);
  }
}

// ---------------------------------------------------------------- column sums: out[c] += sum_r X[r,c]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int cols, float* __restrict__ out) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  __shared__ float red[8][132];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  float4 acc = make_float4(0, 0, 0, 0);
  if (c < cols) {
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + warp; r < rows; r += static_cast<int64_t>(gridDim.y) * 8) {
      const float4 v = ld4(x + r * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[warp][lane * 4 + 0] = acc.x; red[warp][lane * 4 + 1] = acc.y;
  red[warp][lane * 4 + 2] = acc.z; red[warp][lane * 4 + 3] = acc.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < cols) atomicAdd(out + cc, s);
  }
}

int persistent_grid(int64_t work_blocks, int per_sm) {
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  return static_cast<int>(work_blocks < cap ? (work_blocks > 0 ? work_blocks : 1) : cap);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int add_ln_fwd(cudaStream_t stream, const float* a, const float* b, const float* gamma, const float* beta, float* out,
               float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
               const DropoutCfg& drop, const float* post, int64_t post_rows) {
  if (rows == 0) return ST_OK;
  ST_REQUIRE(!post || (post_rows > 0 && aligned16(post)), "add_ln_fwd: post operand needs post_rows > 0 and 16-byte alignment");
  ST_REQUIRE(d > 0 && (d & 3) == 0 && d <= 1024, "add_ln_fwd: d=%d must be a multiple of 4 and <= 1024", d);
  ST_REQUIRE(aligned16(a) && (!b || aligned16(b)) && aligned16(gamma) && aligned16(beta) && aligned16(out) &&
                 (!z_out || aligned16(z_out)),
             "add_ln_fwd: pointers must be 16-byte aligned");
  const int grid = persistent_grid((rows + LN_WARPS - 1) / LN_WARPS, 8);
  ProfScope prof(stream, PROF_LN_FWD, (b ? 3.0 : 2.0) * rows * d * 4 + (z_out ? 1.0 * rows * d * 4 : 0.0));
#define ST_LAUNCH(VPL)                                                                                            \
  ST_CHECK_CUDA(launch_pdl(add_ln_fwd_kernel<VPL>, dim3(grid), dim3(LN_THREADS), 0, stream, a, b, gamma, beta, out, z_out, \
                           mean_out, rstd_out, rows, d, eps, round_out, drop.thresh, drop.scale, drop.seed, post, post_rows))
  if (d <= 128) ST_LAUNCH(1);
  else if (d <= 256) ST_LAUNCH(2);
  else if (d <= 512) ST_LAUNCH(4);
  else ST_LAUNCH(8);
#undef ST_LAUNCH
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int add_ln_bwd(cudaStream_t stream, const float* dy, const float* z, const float* mean, const float* rstd,
               const float* gamma, float* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d,
               int round_out, const DropoutCfg& drop, const float* gate, float gate_scale) {
  if (rows == 0) return ST_OK;
  ST_REQUIRE(!gate || aligned16(gate), "add_ln_bwd: gate must be 16-byte aligned");
  ST_REQUIRE(d > 0 && (d & 3) == 0 && d <= 1024, "add_ln_bwd: d=%d must be a multiple of 4 and <= 1024", d);
  ST_REQUIRE(aligned16(dy) && aligned16(z) && aligned16(gamma) && aligned16(dz), "add_ln_bwd: pointers must be 16-byte aligned");
  const int grid = persistent_grid((rows + LN_WARPS - 1) / LN_WARPS, 4);
  ProfScope prof(stream, PROF_LN_BWD, 3.0 * rows * d * 4);
#define ST_LAUNCH(VPL)                                                                                         \
  ST_CHECK_CUDA(launch_pdl(add_ln_bwd_kernel<VPL>, dim3(grid), dim3(LN_THREADS), 0, stream, dy, z, mean, rstd, gamma, dz, \
                           dgamma, dbeta, dzsum, rows, d, round_out, drop.thresh, drop.scale, drop.seed, gate, gate_scale))
  if (d <= 128) ST_LAUNCH(1);
  else if (d <= 256) ST_LAUNCH(2);
  else if (d <= 512) ST_LAUNCH(4);
  else ST_LAUNCH(8);
#undef ST_LAUNCH
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int round_tf32_2d(cudaStream_t stream, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols) {
  if (rows == 0 || cols == 0) return ST_OK;
  ST_REQUIRE(lds >= cols && ldd >= cols, "round_tf32: leading dimensions (%lld, %lld) smaller than cols=%d",
             (long long)lds, (long long)ldd, cols);
  if (!((cols & 3) == 0 && (lds & 3) == 0 && (ldd & 3) == 0 && aligned16(src) && aligned16(dst))) {
    const int64_t n = rows * cols;
    ProfScope prof(stream, PROF_ROUND, 2.0 * rows * cols * 4);
    ST_CHECK_CUDA(launch_pdl(round_tf32_scalar_kernel, dim3(persistent_grid((n + 255) / 256, 16)), dim3(256), 0, stream, src,
                             lds, dst, ldd, rows, cols));
    ST_CHECK_LAUNCH();
    return ST_OK;
  }
  const int64_t total = rows * (cols / 4);
  const int grid = persistent_grid((total + 255) / 256, 16);
  ProfScope prof(stream, PROF_ROUND, 2.0 * rows * cols * 4);
  ST_CHECK_CUDA(launch_pdl(round_tf32_kernel, dim3(grid), dim3(256), 0, stream, src, lds, dst, ldd, rows, cols / 4));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int colsum_add(cudaStream_t stream, const float* x, int64_t ld, int64_t rows, int cols, float* out) {
  if (rows == 0 || cols == 0) return ST_OK;
  // float4 loads: a ragged width is fine as long as the (padded) row is long enough to read the last group
  ST_REQUIRE((ld & 3) == 0 && ld >= ((cols + 3) & ~3) && aligned16(x),
             "colsum: ld must be a multiple of 4 and >= cols rounded up to 4 (cols=%d ld=%lld)", cols, (long long)ld);
  dim3 grid((cols + 127) / 128, 1);
  int64_t ychunks = (rows + 63) / 64;
  const int64_t cap = (static_cast<int64_t>(num_sms()) * 8 + grid.x - 1) / grid.x;
  grid.y = static_cast<unsigned>(ychunks < cap ? ychunks : cap);
  ProfScope prof(stream, PROF_COLSUM, 1.0 * rows * cols * 4);
  ST_CHECK_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, x, ld, rows, cols, out));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
