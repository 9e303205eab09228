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

template <typename T> __device__ __forceinline__ float4 ld4(const T* p) { return ldv4(p); }
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v) { stv4(p, v); }

// Column owned by register group i of lane `lane`.  Narrow: lane-interleaved groups of 4 ((i*32 + lane)*4).  Wide (16-bit
// tensors, d a multiple of 256): groups 2j and 2j+1 are ADJACENT — 8 consecutive columns per lane — so a 16-bit tensor is
// accessed with one 16-byte transaction per lane and pair (8-byte accesses left these kernels latency-bound:
// LayerNorm backward 2.4 TB/s in bf16 against 3.1 TB/s in fp32).
template <bool WIDE> __device__ __forceinline__ int ln_col(int i, int lane) {
  return WIDE ? ((i >> 1) * 32 + lane) * 8 + (i & 1) * 4 : (i * 32 + lane) * 4;
}
// load / store register groups i (even) and i+1 of a row: one 16-byte access for 16-bit types in the wide mapping
template <bool WIDE, typename T>
__device__ __forceinline__ void ld_pair(const T* row, int c, float4& a, float4& b) {
  if constexpr (WIDE && sizeof(T) == 2) {
    const uint4 w = *reinterpret_cast<const uint4*>(row + c);
    const float2 p0 = unpack2<T>(w.x), p1 = unpack2<T>(w.y), p2 = unpack2<T>(w.z), p3 = unpack2<T>(w.w);
    a = make_float4(p0.x, p0.y, p1.x, p1.y);
    b = make_float4(p2.x, p2.y, p3.x, p3.y);
  } else {
    a = ld4(row + c);
    b = ld4(row + c + 4);
  }
}
template <bool WIDE, typename T>
__device__ __forceinline__ void st_pair(T* row, int c, const float4& a, const float4& b) {
  if constexpr (WIDE && sizeof(T) == 2) {
    *reinterpret_cast<uint4*>(row + c) = make_uint4(pack2<T>(a.x, a.y), pack2<T>(a.z, a.w), pack2<T>(b.x, b.y), pack2<T>(b.z, b.w));
  } else {
    st4(row + c, a);
    st4(row + c + 4, b);
  }
}

// ---------------------------------------------------------------- forward
// InT: element type of a and b (fp32, or the 16-bit activation type); OutT: element type of `out`.  z_out, the
// statistics, gamma / beta and `post` are always fp32.
template <int VPL, typename InT, typename OutT, bool WIDE>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_fwd_kernel(const InT* __restrict__ a, const InT* __restrict__ b, const float* __restrict__ gamma,
                  const float* __restrict__ beta, OutT* __restrict__ out, float* __restrict__ z_out,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int d, float eps,
                  int round_out, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed,
                  const float* __restrict__ post, int64_t post_rows, __half* __restrict__ out_h16) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);

  float4 g[VPL], bt[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = ln_col<WIDE>(i, lane);
    if (c < d) { g[i] = ld4(gamma + c); bt[i] = ld4(beta + c); }
  }

  // The next row's loads are issued before the current row is reduced: one row (2 KB at d = 512) in flight per warp
  // leaves the kernel latency-bound at ~60 % of the HBM rate with the 16 warps per SM its registers allow.
  const int64_t stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  float4 xn[VPL];
  auto load_row = [&](int64_t r, float4 (&dst)[VPL]) {
    if constexpr (WIDE) {
#pragma unroll
      for (int i = 0; i < VPL; i += 2) {
        const int c = ln_col<WIDE>(i, lane);
        if (c < d) ld_pair<WIDE>(a + r * d, c, dst[i], dst[i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = ln_col<WIDE>(i, lane);
        if (c < d) dst[i] = ld4(a + r * d + c);
      }
    }
  };
  if (row_first < rows) load_row(row_first, xn);
  for (int64_t row = row_first; row < rows; row += stride) {
    const InT* br = b ? b + row * d : nullptr;
    float4 x[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) x[i] = xn[i];
    if (row + stride < rows) load_row(row + stride, xn);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = ln_col<WIDE>(i, lane);
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
      const int c = ln_col<WIDE>(i, lane);
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
      const int c = ln_col<WIDE>(i, lane);
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
        if (sizeof(OutT) == 4 && round_out) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tf32_rna(o[t]);
        }
        if constexpr (WIDE) {
          x[i] = make_float4(o[0], o[1], o[2], o[3]);      // x[i] is dead from here on: park the result for the paired store
          if (i & 1) st_pair<WIDE>(out + row * d, c - 4, x[i - 1], x[i]);
        } else {
          st4(out + row * d + c, make_float4(o[0], o[1], o[2], o[3]));
          if (sizeof(OutT) == 4 && out_h16)     // the fp16 copy the next operator consumes (ST_DTYPE_F32_H16)
            *reinterpret_cast<uint2*>(out_h16 + row * d + c) = make_uint2(pack2<__half>(o[0], o[1]), pack2<__half>(o[2], o[3]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------- backward
// dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat;  dbeta += dy;
// dzsum += dz (the bias gradient of the linear layer that produced z).
// DyT / DzT: element types of dy and dz (fp32, or a 16-bit activation type); z, statistics, gamma, gate are fp32.
// amax: optional device scalar max|dy| — the stored dz is multiplied by grad_scale_from_amax(*amax) (mixed mode: fp32 dy,
// fp16 dz); dgamma / dbeta / dzsum are accumulated from the UNSCALED values.
template <int VPL, typename DyT, typename DzT, bool WIDE>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_bwd_kernel(const DyT* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean_in,
                  const float* __restrict__ rstd_in, const float* __restrict__ gamma, DzT* __restrict__ dz,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dzsum, int64_t rows,
                  int d, int round_out, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed,
                  const float* __restrict__ gate, float gate_scale, const float* __restrict__ amax, float* __restrict__ clear_scalar) {
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  // a scalar a LATER kernel of the same operator accumulates into with atomicMax (st_*_bwd_args.dq_amax / dx_amax): cleared
  // here, in stream order before that kernel's first global access, instead of by a memset node between the launches
  if (clear_scalar && blockIdx.x == 0 && threadIdx.x == 0) *clear_scalar = 0.f;
  __shared__ float red[LN_WARPS][32 * 4 + 4];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  const float out_scale = amax ? grad_scale_from_amax(*amax) : 1.f;

  float4 g[VPL];
  float4 acc_g[VPL], acc_b[VPL], acc_z[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = ln_col<WIDE>(i, lane);
    g[i] = (c < d) ? ld4(gamma + c) : make_float4(0, 0, 0, 0);
    acc_g[i] = acc_b[i] = acc_z[i] = make_float4(0, 0, 0, 0);
  }

  // next row's dy / z / statistics are in flight while the current row is reduced (see the forward kernel)
  const int64_t stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  float4 dyn[VPL], zn[VPL];
  float mean_n = 0.f, rstd_n = 0.f;
  auto load_row = [&](int64_t r) {
    mean_n = mean_in[r]; rstd_n = rstd_in[r];
    if constexpr (WIDE) {
#pragma unroll
      for (int i = 0; i < VPL; i += 2) {
        const int c = ln_col<WIDE>(i, lane);
        if (c < d) { ld_pair<WIDE>(dy + r * d, c, dyn[i], dyn[i + 1]); ld_pair<WIDE>(z + r * d, c, zn[i], zn[i + 1]); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = ln_col<WIDE>(i, lane);
        if (c < d) { dyn[i] = ld4(dy + r * d + c); zn[i] = ld4(z + r * d + c); }
      }
    }
  };
  if (row_first < rows) load_row(row_first);
  for (int64_t row = row_first; row < rows; row += stride) {
    const float mean = mean_n, rstd = rstd_n;
    float4 dyc[VPL], zc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) { dyc[i] = dyn[i]; zc[i] = zn[i]; }
    if (row + stride < rows) load_row(row + stride);
    float4 xh[VPL], gy[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = ln_col<WIDE>(i, lane);
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
      const int c = ln_col<WIDE>(i, lane);
      if (c < d) {
        float o[4] = {rstd * (gy[i].x - c1 - xh[i].x * c2), rstd * (gy[i].y - c1 - xh[i].y * c2),
                      rstd * (gy[i].z - c1 - xh[i].z * c2), rstd * (gy[i].w - c1 - xh[i].w * c2)};
        if (gate) {   // gradient through dropout(relu(.)) whose output is `gate`: pass (scaled) where it was kept and positive
          const float4 gv = ld4(gate + row * d + c);
          o[0] = gv.x > 0.f ? o[0] * gate_scale : 0.f; o[1] = gv.y > 0.f ? o[1] * gate_scale : 0.f;
          o[2] = gv.z > 0.f ? o[2] * gate_scale : 0.f; o[3] = gv.w > 0.f ? o[3] * gate_scale : 0.f;
        }
        acc_z[i].x += o[0]; acc_z[i].y += o[1]; acc_z[i].z += o[2]; acc_z[i].w += o[3];
        if (amax) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] *= out_scale;
        }
        if (sizeof(DzT) == 4 && round_out) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tf32_rna(o[t]);
        }
        if constexpr (WIDE) {
          xh[i] = make_float4(o[0], o[1], o[2], o[3]);     // xh[i] is dead from here on: park the result for the paired store
          if (i & 1) st_pair<WIDE>(dz + row * d, c - 4, xh[i - 1], xh[i]);
        } else {
          st4(dz + row * d + c, make_float4(o[0], o[1], o[2], o[3]));
        }
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
        const int c = ln_col<WIDE>(i, threadIdx.x >> 2) + (threadIdx.x & 3);
        if (c < d) atomicAdd(dst + c, s);
      }
    }
  };
  reduce_store(acc_g, dgamma);
  reduce_store(acc_b, dbeta);
  reduce_store(acc_z, dzsum);
}

// ---------------------------------------------------------------- backward, fp32 dy: rows staged through shared memory
// The register-prefetch kernel above keeps ONE row per warp in flight and needs ~195 registers at d = 512 (8 warps per SM):
// 32 KB in flight per SM, 2.4 TB/s.  Here every warp owns a ring of LNB_STAGES row slots in shared memory filled with
// cp.async (16 bytes per lane and request, no registers held while the data is in flight), so LNB_STAGES rows of dy and z
// per warp are outstanding while one is reduced.  Every lane reads back exactly the bytes its own requests wrote — no
// cross-lane hazard, no barrier; the statistics (8 bytes per row) stay on the one-row register prefetch.
constexpr int LNB_STAGES = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VPL, typename DzT>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_bwd_staged_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean_in,
                         const float* __restrict__ rstd_in, const float* __restrict__ gamma, DzT* __restrict__ dz,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dzsum, int64_t rows,
                         int d, int round_out, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed,
                         const float* __restrict__ amax, float* __restrict__ clear_scalar) {
  constexpr int ROWF = VPL * 128;    // floats per row slot (d <= ROWF)
  extern __shared__ __align__(16) float ln_ring[];
  __shared__ float red[LN_WARPS][32 * 4 + 4];
  pdl_wait();      // programmatic dependent launch (st_host.h): before the first global access
  pdl_trigger();
  if (clear_scalar && blockIdx.x == 0 && threadIdx.x == 0) *clear_scalar = 0.f;   // see add_ln_bwd_kernel
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  const float out_scale = amax ? grad_scale_from_amax(*amax) : 1.f;
  float* ring = ln_ring + warp * (LNB_STAGES * 2 * ROWF);

  float4 g[VPL];
  float4 acc_g[VPL], acc_b[VPL], acc_z[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    g[i] = (c < d) ? ld4(gamma + c) : make_float4(0, 0, 0, 0);
    acc_g[i] = acc_b[i] = acc_z[i] = make_float4(0, 0, 0, 0);
  }

  const int64_t stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  auto issue = [&](int64_t r, int slot) {     // one commit group per call, empty past the last row: the group count stays uniform
    if (r < rows) {
      float* sdy = ring + (slot * 2) * ROWF;
      float* sz = sdy + ROWF;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { cp_async16(sdy + c, dy + r * d + c); cp_async16(sz + c, z + r * d + c); }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < LNB_STAGES; ++s) issue(row_first + s * stride, s);
  float mean_n = 0.f, rstd_n = 0.f;
  if (row_first < rows) { mean_n = mean_in[row_first]; rstd_n = rstd_in[row_first]; }
  int slot = 0;
  for (int64_t row = row_first; row < rows; row += stride) {
    cp_async_wait<LNB_STAGES - 1>();          // the oldest group — this row — has landed
    const float mean = mean_n, rstd = rstd_n;
    float4 dyc[VPL], zc[VPL];
    {
      const float* sdy = ring + (slot * 2) * ROWF;
      const float* sz = sdy + ROWF;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { dyc[i] = *reinterpret_cast<const float4*>(sdy + c); zc[i] = *reinterpret_cast<const float4*>(sz + c); }
      }
    }
    issue(row + LNB_STAGES * stride, slot);   // refill the slot just read (same lane, same bytes: program order suffices)
    if (row + stride < rows) { mean_n = mean_in[row + stride]; rstd_n = rstd_in[row + stride]; }
    slot = (slot + 1 == LNB_STAGES) ? 0 : slot + 1;

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
        acc_z[i].x += o[0]; acc_z[i].y += o[1]; acc_z[i].z += o[2]; acc_z[i].w += o[3];
        if (amax) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] *= out_scale;
        }
        if (sizeof(DzT) == 4 && round_out) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tf32_rna(o[t]);
        }
        st4(dz + row * d + c, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }
  cp_async_wait<0>();

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
        const int c = (i * 32 + (threadIdx.x >> 2)) * 4 + (threadIdx.x & 3);
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
// out <- max(out, max|x|) as an atomicMax on the bit pattern (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, int64_t n4, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  float m = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = x4[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int cols, float* __restrict__ out, const float* __restrict__ amax) {
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
    if (amax) s *= 1.f / grad_scale_from_amax(*amax);      // mixed mode: x carries the operator's gradient scale
    if (cc < cols) atomicAdd(out + cc, s);
  }
}

int persistent_grid(int64_t work_blocks, int per_sm) {
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  return static_cast<int>(work_blocks < cap ? (work_blocks > 0 ? work_blocks : 1) : cap);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

}  // namespace

namespace {

template <typename InT, typename OutT>
int add_ln_fwd_t(cudaStream_t stream, const InT* a, const InT* b, const float* gamma, const float* beta, OutT* out,
                 float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
                 const DropoutCfg& drop, const float* post, int64_t post_rows, void* out_h16 = nullptr) {
  if (rows == 0) return ST_OK;
  ST_REQUIRE(!post || (post_rows > 0 && aligned16(post)), "add_ln_fwd: post operand needs post_rows > 0 and 16-byte alignment");
  ST_REQUIRE(!out_h16 || (sizeof(OutT) == 4 && sizeof(InT) == 4 && aligned8(out_h16)), "add_ln_fwd: the fp16 copy needs fp32 in / out");
  ST_REQUIRE(d > 0 && (d & 3) == 0 && d <= 1024, "add_ln_fwd: d=%d must be a multiple of 4 and <= 1024", d);
  ST_REQUIRE(aligned8(a) && (!b || aligned8(b)) && aligned16(gamma) && aligned16(beta) && aligned8(out) &&
                 (!z_out || aligned16(z_out)) && (sizeof(InT) == 2 || (aligned16(a) && (!b || aligned16(b)))) &&
                 (sizeof(OutT) == 2 || aligned16(out)),
             "add_ln_fwd: pointers must be 16-byte aligned");
  const int grid = persistent_grid((rows + LN_WARPS - 1) / LN_WARPS, 8);
  ProfScope prof(stream, PROF_LN_FWD, (b ? 2.0 : 1.0) * rows * d * sizeof(InT) + 1.0 * rows * d * sizeof(OutT) + (z_out ? 1.0 * rows * d * 4 : 0.0) +
                                          (out_h16 ? 2.0 * rows * d : 0.0));
#define ST_LAUNCH(VPL, W)                                                                                         \
  ST_CHECK_CUDA(launch_pdl(add_ln_fwd_kernel<VPL, InT, OutT, W>, dim3(grid), dim3(LN_THREADS), 0, stream, a, b, gamma, beta, out, z_out, \
                           mean_out, rstd_out, rows, d, eps, round_out, drop.thresh, drop.scale, drop.seed, post, post_rows,         \
                           static_cast<__half*>(out_h16)))
  constexpr bool kAny16 = sizeof(InT) == 2 || sizeof(OutT) == 2;
  const bool wide = kAny16 && (d % 256) == 0 && aligned16(a) && (!b || aligned16(b)) && aligned16(out);
  if (wide) {
    if constexpr (kAny16) {
      if (d <= 256) ST_LAUNCH(2, true);
      else if (d <= 512) ST_LAUNCH(4, true);
      else ST_LAUNCH(8, true);
    }
  } else if (d <= 128) ST_LAUNCH(1, false);
  else if (d <= 256) ST_LAUNCH(2, false);
  else if (d <= 512) ST_LAUNCH(4, false);
  else ST_LAUNCH(8, false);
#undef ST_LAUNCH
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace

int add_ln_fwd(cudaStream_t stream, const float* a, const float* b, const float* gamma, const float* beta, float* out,
               float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
               const DropoutCfg& drop, const float* post, int64_t post_rows) {
  return add_ln_fwd_t<float, float>(stream, a, b, gamma, beta, out, z_out, mean_out, rstd_out, rows, d, eps, round_out, drop, post,
                                    post_rows);
}

// in_dt: element type of a / b (ST_DTYPE_F32 or out_dt); out_dt: element type of out
int add_ln_fwd_any(cudaStream_t stream, int in_dt, int out_dt, const void* a, const void* b, const float* gamma, const float* beta,
                   void* out, float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
                   const DropoutCfg& drop, const float* post, int64_t post_rows, void* out_h16) {
#define ST_CALL(InT, OutT)                                                                                                   \
  return add_ln_fwd_t<InT, OutT>(stream, static_cast<const InT*>(a), static_cast<const InT*>(b), gamma, beta, static_cast<OutT*>(out), \
                                 z_out, mean_out, rstd_out, rows, d, eps, round_out, drop, post, post_rows, out_h16)
  if (in_dt == ST_DTYPE_F32 && out_dt == ST_DTYPE_F32) ST_CALL(float, float);
  if (in_dt == ST_DTYPE_F32 && out_dt == ST_DTYPE_F16) ST_CALL(float, __half);
  if (in_dt == ST_DTYPE_F32 && out_dt == ST_DTYPE_BF16) ST_CALL(float, __nv_bfloat16);
  if (in_dt == ST_DTYPE_F16 && out_dt == ST_DTYPE_F16) ST_CALL(__half, __half);
  if (in_dt == ST_DTYPE_BF16 && out_dt == ST_DTYPE_BF16) ST_CALL(__nv_bfloat16, __nv_bfloat16);
#undef ST_CALL
  set_error("add_ln_fwd: unsupported dtype combination (in %d, out %d)", in_dt, out_dt);
  return ST_ERR_INVALID;
}

namespace {

template <typename DyT, typename DzT>
int add_ln_bwd_t(cudaStream_t stream, const DyT* dy, const float* z, const float* mean, const float* rstd,
                 const float* gamma, DzT* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d,
                 int round_out, const DropoutCfg& drop, const float* gate, float gate_scale, const float* amax = nullptr,
                 float* clear_scalar = nullptr) {
  if (rows == 0) return ST_OK;
  ST_REQUIRE(!gate || aligned16(gate), "add_ln_bwd: gate must be 16-byte aligned");
  ST_REQUIRE(d > 0 && (d & 3) == 0 && d <= 1024, "add_ln_bwd: d=%d must be a multiple of 4 and <= 1024", d);
  ST_REQUIRE(aligned8(dy) && aligned16(z) && aligned16(gamma) && aligned8(dz) && (sizeof(DyT) == 2 || aligned16(dy)) &&
                 (sizeof(DzT) == 2 || aligned16(dz)),
             "add_ln_bwd: pointers must be 16-byte aligned");
  ProfScope prof(stream, PROF_LN_BWD, 1.0 * rows * d * 4 + 1.0 * rows * d * (sizeof(DyT) + sizeof(DzT)));
  if constexpr (sizeof(DyT) == 4) {
    // fp32 dy (TF32 path, fp16-operand engine): rows staged through shared memory with cp.async
    if (!gate && d <= 512 && aligned16(dy) && aligned16(z) && (sizeof(DzT) == 4 ? aligned16(dz) : aligned8(dz)) && !get_option("ln_bwd_registers")) {
      const int vpl = d <= 128 ? 1 : (d <= 256 ? 2 : 4);
      const size_t smem = static_cast<size_t>(LN_WARPS) * LNB_STAGES * 2 * vpl * 128 * sizeof(float);
      const int per_sm = vpl == 4 ? 1 : (vpl == 2 ? 2 : 4);
      const int sgrid = persistent_grid((rows + LN_WARPS - 1) / LN_WARPS, per_sm);
#define ST_LAUNCH_S(VPL)                                                                                                \
  do {                                                                                                                  \
    auto kern = add_ln_bwd_staged_kernel<VPL, DzT>;                                                                     \
    static bool attr = false;                                                                                           \
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); attr = true; } \
    ST_CHECK_CUDA(launch_pdl(kern, dim3(sgrid), dim3(LN_THREADS), smem, stream, dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dzsum,  \
                             rows, d, round_out, drop.thresh, drop.scale, drop.seed, amax, clear_scalar));               \
  } while (0)
      if (vpl == 1) ST_LAUNCH_S(1);
      else if (vpl == 2) ST_LAUNCH_S(2);
      else ST_LAUNCH_S(4);
#undef ST_LAUNCH_S
      ST_CHECK_LAUNCH();
      return ST_OK;
    }
  }
  const int grid = persistent_grid((rows + LN_WARPS - 1) / LN_WARPS, 4);
#define ST_LAUNCH(VPL, W)                                                                                      \
  ST_CHECK_CUDA(launch_pdl(add_ln_bwd_kernel<VPL, DyT, DzT, W>, dim3(grid), dim3(LN_THREADS), 0, stream, dy, z, mean, rstd, gamma, dz, \
                           dgamma, dbeta, dzsum, rows, d, round_out, drop.thresh, drop.scale, drop.seed, gate, gate_scale, amax, clear_scalar))
  constexpr bool k16 = sizeof(DyT) == 2 || sizeof(DzT) == 2;
  const bool wide = k16 && (d % 256) == 0 && aligned16(dy) && aligned16(dz) && !gate;
  if (wide) {
    if constexpr (k16) {
      if (d <= 256) ST_LAUNCH(2, true);
      else if (d <= 512) ST_LAUNCH(4, true);
      else ST_LAUNCH(8, true);
    }
  } else if (d <= 128) ST_LAUNCH(1, false);
  else if (d <= 256) ST_LAUNCH(2, false);
  else if (d <= 512) ST_LAUNCH(4, false);
  else ST_LAUNCH(8, false);
#undef ST_LAUNCH
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace

int add_ln_bwd(cudaStream_t stream, const float* dy, const float* z, const float* mean, const float* rstd,
               const float* gamma, float* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d,
               int round_out, const DropoutCfg& drop, const float* gate, float gate_scale) {
  return add_ln_bwd_t<float, float>(stream, dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dzsum, rows, d, round_out, drop, gate, gate_scale);
}

// mixed mode: fp32 dy -> fp16 dz scaled by grad_scale_from_amax(*amax)
int add_ln_bwd_mixed(cudaStream_t stream, const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                     void* dz16, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, const DropoutCfg& drop,
                     const float* amax, float* clear_scalar) {
  return add_ln_bwd_t<float, __half>(stream, dy, z, mean, rstd, gamma, static_cast<__half*>(dz16), dgamma, dbeta, dzsum, rows, d, 0, drop,
                                     nullptr, 1.f, amax, clear_scalar);
}

// dt: element type of dy and dz
int add_ln_bwd_any(cudaStream_t stream, int dt, const void* dy, const float* z, const float* mean, const float* rstd,
                   const float* gamma, void* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, int round_out,
                   const DropoutCfg& drop, const float* gate, float gate_scale) {
  switch (dt) {
    case ST_DTYPE_F32:
      return add_ln_bwd_t<float, float>(stream, static_cast<const float*>(dy), z, mean, rstd, gamma, static_cast<float*>(dz), dgamma,
                                        dbeta, dzsum, rows, d, round_out, drop, gate, gate_scale);
    case ST_DTYPE_F16:
      return add_ln_bwd_t<__half, __half>(stream, static_cast<const __half*>(dy), z, mean, rstd, gamma, static_cast<__half*>(dz), dgamma,
                                          dbeta, dzsum, rows, d, round_out, drop, gate, gate_scale);
    case ST_DTYPE_BF16:
      return add_ln_bwd_t<__nv_bfloat16, __nv_bfloat16>(stream, static_cast<const __nv_bfloat16*>(dy), z, mean, rstd, gamma,
                                                        static_cast<__nv_bfloat16*>(dz), dgamma, dbeta, dzsum, rows, d, round_out, drop,
                                                        gate, gate_scale);
  }
  set_error("add_ln_bwd: bad dtype %d", dt);
  return ST_ERR_INVALID;
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

namespace {
template <typename T>
int colsum_add_t(cudaStream_t stream, const T* x, int64_t ld, int64_t rows, int cols, float* out, const float* amax = nullptr) {
  if (rows == 0 || cols == 0) return ST_OK;
  // 4-element vector loads: a ragged width is fine as long as the (padded) row is long enough to read the last group
  ST_REQUIRE((ld & 3) == 0 && ld >= ((cols + 3) & ~3) && (sizeof(T) == 2 ? aligned8(x) : aligned16(x)),
             "colsum: ld must be a multiple of 4 and >= cols rounded up to 4 (cols=%d ld=%lld)", cols, (long long)ld);
  dim3 grid((cols + 127) / 128, 1);
  int64_t ychunks = (rows + 63) / 64;
  const int64_t cap = (static_cast<int64_t>(num_sms()) * 8 + grid.x - 1) / grid.x;
  grid.y = static_cast<unsigned>(ychunks < cap ? ychunks : cap);
  ProfScope prof(stream, PROF_COLSUM, 1.0 * rows * cols * sizeof(T));
  ST_CHECK_CUDA(launch_pdl(colsum_kernel<T>, grid, dim3(256), 0, stream, x, ld, rows, cols, out, amax));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// element-type conversion copy (2-D, strided; cols and leading dimensions multiples of 4)
template <typename S, typename D>
__global__ void __launch_bounds__(256)
cast_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst, int64_t ldd, int64_t rows, int cols4, float scale) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = rows * cols4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    float4 v = ldv4(src + r * lds + c);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    stv4(dst + r * ldd + c, v);
  }
}
template <typename S, typename D>
__global__ void __launch_bounds__(256)
cast_scalar_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst, int64_t ldd, int64_t rows, int cols, float scale) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    dst[r * ldd + c] = from_f32<D>(to_f32(src[r * lds + c]) * scale);
  }
}
template <typename S, typename D>
int cast_t(cudaStream_t stream, const S* src, int64_t lds, D* dst, int64_t ldd, int64_t rows, int cols, float scale) {
  ProfScope prof(stream, PROF_ROUND, 1.0 * rows * cols * (sizeof(S) + sizeof(D)));
  const bool vec = (cols & 3) == 0 && (lds & 3) == 0 && (ldd & 3) == 0 && (sizeof(S) == 2 ? aligned8(src) : aligned16(src)) &&
                   (sizeof(D) == 2 ? aligned8(dst) : aligned16(dst));
  if (vec) {
    const int64_t total = rows * (cols / 4);
    ST_CHECK_CUDA(launch_pdl(cast_kernel<S, D>, dim3(persistent_grid((total + 255) / 256, 16)), dim3(256), 0, stream, src, lds, dst, ldd,
                             rows, cols / 4, scale));
  } else {
    const int64_t total = rows * cols;
    ST_CHECK_CUDA(launch_pdl(cast_scalar_kernel<S, D>, dim3(persistent_grid((total + 255) / 256, 16)), dim3(256), 0, stream, src, lds,
                             dst, ldd, rows, cols, scale));
  }
  ST_CHECK_LAUNCH();
  return ST_OK;
}
}  // namespace

int colsum_add(cudaStream_t stream, const float* x, int64_t ld, int64_t rows, int cols, float* out) {
  return colsum_add_t<float>(stream, x, ld, rows, cols, out);
}
int colsum_add_any(cudaStream_t stream, int dt, const void* x, int64_t ld, int64_t rows, int cols, float* out, const float* amax) {
  switch (dt) {
    case ST_DTYPE_F32: return colsum_add_t<float>(stream, static_cast<const float*>(x), ld, rows, cols, out, amax);
    case ST_DTYPE_F16: return colsum_add_t<__half>(stream, static_cast<const __half*>(x), ld, rows, cols, out, amax);
    case ST_DTYPE_BF16: return colsum_add_t<__nv_bfloat16>(stream, static_cast<const __nv_bfloat16*>(x), ld, rows, cols, out, amax);
  }
  set_error("colsum: bad dtype %d", dt);
  return ST_ERR_INVALID;
}

// *out = max|x| over n fp32 values (n a multiple of 4, x 16-byte aligned); `out` is cleared first
int amax_abs(cudaStream_t stream, const float* x, int64_t n, float* out) {
  ST_REQUIRE((n & 3) == 0 && aligned16(x), "amax: n=%lld must be a multiple of 4 and x 16-byte aligned", (long long)n);
  ST_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  if (n == 0) return ST_OK;
  ProfScope prof(stream, PROF_ROUND, 4.0 * n);
  ST_CHECK_CUDA(launch_pdl(amax_kernel, dim3(persistent_grid((n / 4 + 255) / 256, 8)), dim3(256), 0, stream, x, n / 4, out));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// dst = (dst type)(src * scale); one of the two types is fp32
int cast_2d(cudaStream_t stream, const void* src, int src_dt, int64_t lds, void* dst, int dst_dt, int64_t ldd, int64_t rows, int cols,
            float scale) {
  if (rows == 0 || cols == 0) return ST_OK;
  ST_REQUIRE(lds >= cols && ldd >= cols, "cast: leading dimensions (%lld, %lld) smaller than cols=%d", (long long)lds, (long long)ldd, cols);
#define ST_CAST(S, D) return cast_t<S, D>(stream, static_cast<const S*>(src), lds, static_cast<D*>(dst), ldd, rows, cols, scale)
  if (src_dt == ST_DTYPE_F32 && dst_dt == ST_DTYPE_F16) ST_CAST(float, __half);
  if (src_dt == ST_DTYPE_F32 && dst_dt == ST_DTYPE_BF16) ST_CAST(float, __nv_bfloat16);
  if (src_dt == ST_DTYPE_F16 && dst_dt == ST_DTYPE_F32) ST_CAST(__half, float);
  if (src_dt == ST_DTYPE_BF16 && dst_dt == ST_DTYPE_F32) ST_CAST(__nv_bfloat16, float);
  if (src_dt == ST_DTYPE_F32 && dst_dt == ST_DTYPE_F32) ST_CAST(float, float);
#undef ST_CAST
  set_error("cast: unsupported conversion %d -> %d", src_dt, dst_dt);
  return ST_ERR_INVALID;
}

}  // namespace st
