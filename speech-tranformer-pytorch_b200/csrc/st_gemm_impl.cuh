// st_gemm_impl.cuh — persistent, warp-specialised tcgen05 GEMM, templated on the operand element type
// (float = TF32 operands in fp32 storage, __half / __nv_bfloat16 = kind::f16 operands); included by one translation
// unit per element type (st_gemm.cu, st_gemm_h.cu, st_gemm_bf.cu).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..9  : epilogue (tcgen05.ld -> smem transpose -> bias / ReLU / dropout / residual / rounding
//                 -> coalesced global stores); two warps per TMEM lane quarter
//
// Output tile 128 x BN (BN in {64,128,256}); K is consumed in blocks of one 128-byte swizzle row (32 fp32 or 64
// 16-bit elements); accumulators are double-buffered in TMEM (2*BN columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.  Operands may be K-major or MN-major (see
// st_common.cuh; MN-major TF32 tiles use the 32-byte-atom swizzle, 16-bit ones the plain 128-byte swizzle), which
// covers forward (NT), data-gradient (NN) and weight-gradient (TN) GEMMs without any transposed copies in HBM.
//
// Three kernels share the epilogue:
//   gemm_kernel<T,BN,..,CL=1>  independent CTAs (default for small grids / narrow tiles)
//   gemm_kernel<T,256,..,CL=2> 2-CTA cluster, B tile fetched half by each CTA and TMA-multicast to both
//                              (correct, but measured slower than independent CTAs: opt-in, TF32 only)
//   gemm_2sm_kernel<T,..>      CTA pair with tcgen05.mma.cta_group::2: one 256x256 MMA tile per pair, each
//                              CTA stages its own 128 A rows and HALF of the B tile, which cuts the bytes
//                              streamed into each SM per FLOP by 1.5x — the limiter of the K = 512 GEMMs
//
// Epilogue: a TMEM lane is an output row, so tcgen05.ld hands each thread 32 consecutive columns
// of ITS row; storing that directly would touch 32 different 128-byte lines per instruction.  Each
// epilogue warp therefore transposes 32x32 blocks through a swizzled smem tile and stores full row
// segments per instruction (bias and residual loads are coalesced the same way).  With 16-bit operands the
// output is either fp32 (pre-LayerNorm sums, logits, split-K weight gradients) or the operand type.
#pragma once
#include "st_common.cuh"
#include "st_gemm.cuh"
#include "st_kernels.h"
#include "st_host.h"

namespace st {

namespace {

constexpr int BM = 128;
constexpr int A_STAGE_BYTES = BM * 128;      // 128 rows x one 128-byte swizzle row = 16 KB, any element type
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quarter, interleaved over the 32-column chunks
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;   // one XOR-swizzled 32 x 32 fp32 transpose tile per epilogue warp

// float offset of (row, 4-float group g) in a warp's transpose tile: 128-byte rows, 16-byte groups XOR-swizzled by
// the row so that both the row-per-lane writes and the 8-lanes-per-row reads are bank-conflict free without padding
__device__ __forceinline__ int epi_swz(int row, int g) { return row * 32 + ((g ^ (row & 7)) << 2); }

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  void* C;          // fp32, or the operand type when c_lp
  int64_t ldc;      // elements
  int c_lp;
  int M, N, K;
  int m_tiles, n_tiles, k_splits;
  int kblocks_total, kblocks_per_split;
  GemmEpilogue ep;
  int flavour;   // EpiFlavour
  int rows16;    // 0 = transposing epilogue; 1 / 2 = row-layout epilogue, direct / packed-staged stores (epilogue_rows16)
  int clc;       // CTA-pair kernel: 1 = one cluster per tile in the grid, resident clusters take over pending ones' tiles
                 // (cluster launch control) instead of walking a static stride
};

// ---- epilogue ------------------------------------------------------------------------------------------------
// The epilogue runs on 8 warps and, for the K = 512 GEMMs of this path, is the longer leg of the
// (main loop || epilogue) pipeline: it has to be instruction-lean.  The combinations the composite ops use are
// therefore compile-time specialisations ("flavours", chosen on the host); anything else — unaligned C / aux /
// bias, ragged N % 4, exotic flag mixes from the raw st_gemm entry point — takes the generic routine.
enum EpiFlavour : int {
  EPI_PLAIN = 0,            // C = acc (+ bias)
  EPI_ROUND = 1,            // C = tf32(acc (+ bias))                               QKV projections, dctx
  EPI_AUX_ADD = 2,          // C = acc (+ bias) + aux                               out-proj / fc2 + residual, dgrad + residual
  EPI_RELU_DROP_ROUND = 3,  // C = tf32(dropout(relu(acc + bias)))                  fc1
  EPI_AUX_MASK_ROUND = 4,   // C = tf32(aux > 0 ? acc * aux_scale : 0)              dgrad through relu + dropout
  EPI_ATOMIC = 5,           // C += acc (red.global.add)                            split-K wgrad
  EPI_GENERIC = 6,
  EPI_RELU_DROP = 7         // C = dropout(relu(acc + bias)), not rounded               front-end linear (feeds a LayerNorm)
};

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// one 16-byte reduction instead of four scalar ones: split-K wgrad epilogues are bound by L2 atomic operations
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// four consecutive elements <-> float4
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p) {
  const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack2<T>(w.x), b = unpack2<T>(w.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2<T>(a, b), pack2<T>(c, d));
}

// Pull this warp's 32 x BN slice of the aux operand towards L2 while the tile's main loop is still running
// (lane l fetches the 128-byte lines of row l).
template <typename T, int BN>
__device__ __forceinline__ void prefetch_aux_tile(const GemmParams& p, int m_blk, int n_blk, int quarter, int half, int lane) {
  const int row = m_blk * BM + quarter * 32 + lane;
  if (row >= p.M) return;
  const T* a = static_cast<const T*>(p.ep.aux) + static_cast<int64_t>(row) * p.ep.ldaux + n_blk * BN;
  const int cols = min(BN, p.N - n_blk * BN);
  constexpr int LINE = 128 / static_cast<int>(sizeof(T));   // elements per 128-byte line
#pragma unroll
  for (int c = 0; c < BN; c += LINE * (EPI_WARPS / 4))
    if (c + half * LINE < cols) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + c + half * LINE));
}

// Drain this warp's 32 rows of one 128 x BN accumulator.  tmem_acc: TMEM address of (warp's first lane, first
// accumulator column); stg_s: shared-space address of the warp's swizzled 32 x 32 transpose tile.
// Requires the host-checked "vector" conditions: C / aux aligned to 4 elements, bias 16-byte aligned, ldc, ldaux, N
// multiples of 4.  T: operand (and aux) element type; OutT: float or T.
template <typename T, typename OutT, int BN, int F>
__device__ __forceinline__ void epilogue_fast(const GemmParams& p, uint32_t stg_s, uint32_t tmem_acc, int m_blk, int n_blk,
                                              int quarter, int half, int lane) {
  constexpr bool RELU = (F == EPI_RELU_DROP_ROUND || F == EPI_RELU_DROP);
  constexpr bool DROP = (F == EPI_RELU_DROP_ROUND || F == EPI_RELU_DROP);
  constexpr int AUX = (F == EPI_AUX_ADD) ? 1 : (F == EPI_AUX_MASK_ROUND ? 2 : 0);
  constexpr bool ROUND = !Elem<T>::k16 && (F == EPI_ROUND || F == EPI_RELU_DROP_ROUND || F == EPI_AUX_MASK_ROUND);
  constexpr bool ATOMIC = (F == EPI_ATOMIC);
  const GemmEpilogue& ep = p.ep;
  const int lcol = (lane & 7) * 4;   // this lane's 4 columns inside a 32-column chunk
  const int lrow = lane >> 3;        // and its row inside each group of 4 rows
  const int row0 = m_blk * BM + quarter * 32 + lrow;
  const int col0 = n_blk * BN + lcol;
  const int rows_left = p.M - row0;  // row i*4 of this lane is in range iff i*4 < rows_left
  OutT* cp = static_cast<OutT*>(p.C) + static_cast<int64_t>(row0) * p.ldc + col0;
  const int64_t cstep = 4 * p.ldc;
  const T* ap = AUX ? static_cast<const T*>(ep.aux) + static_cast<int64_t>(row0) * ep.ldaux + col0 : nullptr;
  const int64_t astep = 4 * ep.ldaux;
  const bool has_bias = ep.bias != nullptr;
  uint32_t keys[8];
  if (DROP) {
#pragma unroll
    for (int i = 0; i < 8; ++i) keys[i] = dropout_row_key(ep.drop_seed, static_cast<uint64_t>(row0 + 4 * i));
  }
  const int n_chunks = min(BN / 32, (p.N - n_blk * BN + 31) >> 5);
  const bool do_colsum = ep.colsum != nullptr;
  // mixed mode: fp32 results that leave the operator shed the gradient scale its 16-bit operands carry
  constexpr bool OSC = Elem<T>::k16 && (ATOMIC || (sizeof(OutT) == 4 && (AUX == 1 || F == EPI_PLAIN)) || AUX == 2);
  const bool has_osc = OSC && ep.unscale_amax != nullptr;
  const float osc = has_osc ? 1.f / grad_scale_from_amax(__ldg(ep.unscale_amax)) : 1.f;
  constexpr bool AMX = Elem<T>::k16 && !ATOMIC && sizeof(OutT) == 4 && (AUX == 1 || F == EPI_PLAIN);
  const bool do_amx = AMX && ep.amax_out != nullptr;
  float am = 0.f;
#pragma unroll 1
  for (int c = half; c < n_chunks; c += EPI_WARPS / 4) {
    const int col = col0 + c * 32;
    const bool col_ok = col < p.N;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    // residual / ReLU-mask operand and bias: issue the loads first, their latency hides behind the TMEM read and
    // the smem transpose
    float4 aux4[8];
    if (AUX) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        aux4[i] = (col_ok && i * 4 < rows_left) ? ld4(ap + i * astep + c * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ATOMIC && has_bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    {
      uint32_t r[32];
      tmem_ld32(tmem_acc + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(stg_s + epi_swz(lane, j) * 4, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
               __uint_as_float(r[4 * j + 3]));
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (col_ok && i * 4 < rows_left) {
        const float4 s4 = lds128(stg_s + epi_swz(i * 4 + lrow, lane & 7) * 4);
        float v[4] = {s4.x + b4.x, s4.y + b4.y, s4.z + b4.z, s4.w + b4.w};
        if (RELU) {
#pragma unroll
          for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], 0.f);
        }
        if (DROP) {
#pragma unroll
          for (int t = 0; t < 4; t += 2) {  // col is a multiple of 4: one hash per column pair
            const uint32_t bits = dropout_pair(keys[i], col + t);
            v[t] = ((bits & 0xFFFFu) >= ep.drop_thresh) ? v[t] * ep.drop_scale : 0.f;
            v[t + 1] = ((bits >> 16) >= ep.drop_thresh) ? v[t + 1] * ep.drop_scale : 0.f;
          }
        }
        if (AUX == 1) { v[0] += aux4[i].x; v[1] += aux4[i].y; v[2] += aux4[i].z; v[3] += aux4[i].w; }
        if (AUX == 2) {
          v[0] = aux4[i].x > 0.f ? v[0] * ep.aux_scale : 0.f;
          v[1] = aux4[i].y > 0.f ? v[1] * ep.aux_scale : 0.f;
          v[2] = aux4[i].z > 0.f ? v[2] * ep.aux_scale : 0.f;
          v[3] = aux4[i].w > 0.f ? v[3] * ep.aux_scale : 0.f;
        }
        if (OSC && AUX != 2 && has_osc) { v[0] *= osc; v[1] *= osc; v[2] *= osc; v[3] *= osc; }
        if (AMX && do_amx) am = fmaxf(fmaxf(am, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
        cs[0] += v[0]; cs[1] += v[1]; cs[2] += v[2]; cs[3] += v[3];
        if (ROUND) {
#pragma unroll
          for (int t = 0; t < 4; ++t) v[t] = tf32_rna(v[t]);
        }
        OutT* dst = cp + i * cstep + c * 32;
        if constexpr (ATOMIC) {
          red_add_v4(reinterpret_cast<float*>(dst), v[0], v[1], v[2], v[3]);
        } else {
          st4(dst, v[0], v[1], v[2], v[3]);
        }
      }
    }
    if (do_colsum) {   // warp-uniform: the 4 lanes that share these columns (lane >> 3 = 0..3) combine, then one vector reduction
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        cs[t] += __shfl_xor_sync(0xffffffffu, cs[t], 8);
        cs[t] += __shfl_xor_sync(0xffffffffu, cs[t], 16);
      }
      if (lane < 8 && col_ok) red_add_v4(ep.colsum + col, cs[0] * (AUX == 2 ? osc : 1.f), cs[1] * (AUX == 2 ? osc : 1.f),
                                         cs[2] * (AUX == 2 ? osc : 1.f), cs[3] * (AUX == 2 ? osc : 1.f));
    }
    __syncwarp();
  }
  if (AMX && do_amx) {
    am = warp_max(am);
    if (lane == 0 && am > 0.f) atomicMax(reinterpret_cast<unsigned int*>(ep.amax_out), __float_as_uint(am));
  }
}

// Any flag combination, any alignment (scalar loads / stores where needed).
template <typename T, int BN>
__device__ __noinline__ void epilogue_generic(const GemmParams& p, float* stg, uint32_t tmem_acc, int m_blk, int n_blk,
                                              int quarter, int half, int lane) {
  const GemmEpilogue& ep = p.ep;
  const bool c_lp = Elem<T>::k16 && p.c_lp;
  const float osc = (Elem<T>::k16 && !c_lp && ep.unscale_amax != nullptr) ? 1.f / grad_scale_from_amax(__ldg(ep.unscale_amax)) : 1.f;
  const bool vec_ok = !Elem<T>::k16 && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0);
  const int lcol = (lane & 7) * 4;
  const int lrow = lane >> 3;
  const int row_base = m_blk * BM + quarter * 32;
  const bool do_amx = Elem<T>::k16 && !c_lp && !ep.atomic && ep.amax_out != nullptr;
  float am = 0.f;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += EPI_WARPS / 4) {
    const int col = n_blk * BN + c * 32 + lcol;
    if (n_blk * BN + c * 32 >= p.N) break;  // warp-uniform: nothing left in this tile row-block
    {
      uint32_t r[32];
      tmem_ld32(tmem_acc + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + epi_swz(lane, j)) =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                        __uint_as_float(r[4 * j + 3]));
    }
    __syncwarp();
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.bias) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (col + t < p.N) b4[t] = ep.bias[col + t];
    }
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + lrow;
      const int row = row_base + rr;
      if (row >= p.M || col >= p.N) continue;
      const float4 s4 = *reinterpret_cast<const float4*>(stg + epi_swz(rr, lane & 7));
      float v[4] = {s4.x + b4[0], s4.y + b4[1], s4.z + b4[2], s4.w + b4[3]};
      if (ep.relu) {
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], 0.f);
      }
      if (ep.drop_thresh) {
        const uint32_t key = dropout_row_key(ep.drop_seed, static_cast<uint64_t>(row));
#pragma unroll
        for (int t = 0; t < 4; t += 2) {
          const uint32_t bits = dropout_pair(key, col + t);
          v[t] = ((bits & 0xFFFFu) >= ep.drop_thresh) ? v[t] * ep.drop_scale : 0.f;
          v[t + 1] = ((bits >> 16) >= ep.drop_thresh) ? v[t + 1] * ep.drop_scale : 0.f;
        }
      }
      const int64_t coff = static_cast<int64_t>(row) * p.ldc + col;
      const T* ap = ep.aux ? static_cast<const T*>(ep.aux) + static_cast<int64_t>(row) * ep.ldaux + col : nullptr;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (col + t >= p.N) break;
        float x = v[t];
        if (ep.aux_mode == 1) x += to_f32(ap[t]);
        if (ep.aux_mode == 2) x = to_f32(ap[t]) > 0.f ? x * ep.aux_scale : 0.f;
        if (!Elem<T>::k16 && ep.round_tf32) x = tf32_rna(x);
        v[t] = x * osc;
        if (do_amx) am = fmaxf(am, fabsf(v[t]));
      }
      if (c_lp) {
        T* cp = static_cast<T*>(p.C) + coff;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (col + t >= p.N) break;
          cp[t] = from_f32<T>(v[t]);
        }
      } else {
        float* cp = static_cast<float*>(p.C) + coff;
        if (vec_ok && !ep.atomic) {
          *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (col + t >= p.N) break;
            if (ep.atomic) atomicAdd(cp + t, v[t]); else cp[t] = v[t];
          }
        }
      }
    }
    __syncwarp();
  }
  if (do_amx) {
    am = warp_max(am);
    if (lane == 0 && am > 0.f) atomicMax(reinterpret_cast<unsigned int*>(ep.amax_out), __float_as_uint(am));
  }
}

// Row-layout epilogue for 16-bit outputs without aux operand or column sums (QKV / KV projections, dctx, fc1): every
// thread keeps the 32 columns of ITS row that tcgen05.ld delivers, applies bias / ReLU / dropout there and packs them to
// 64 bytes.  The fp32 transpose of epilogue_fast moves 2 x 128 KB per 128 x 256 tile through shared memory — as much as
// the MMAs of a K = 512 main loop read from it — so the two legs of the (main loop || epilogue) pipeline contend for
// shared-memory bandwidth.  Here either nothing goes through shared memory (STAGED = false: each thread stores its own
// 64-byte row segment) or only the packed result does (STAGED = true: 2 x 64 KB per tile, stores coalesced to 64 bytes
// per row and 8 rows per instruction).  Requires N % 8 == 0, ldc % 8 == 0 and a 16-byte aligned C (host-checked).
template <typename T, int BN, bool RELU_DROP, bool STAGED>
__device__ __forceinline__ void epilogue_rows16(const GemmParams& p, uint32_t stg_s, uint32_t tmem_acc, int m_blk, int n_blk,
                                                int quarter, int half, int lane) {
  const GemmEpilogue& ep = p.ep;
  const int row = m_blk * BM + quarter * 32 + lane;
  const bool row_ok = row < p.M;
  const int n_chunks = min(BN / 32, (p.N - n_blk * BN + 31) >> 5);
  const bool has_bias = ep.bias != nullptr;
  const uint32_t key = RELU_DROP ? dropout_row_key(ep.drop_seed, static_cast<uint64_t>(row)) : 0u;
  T* crow = static_cast<T*>(p.C) + static_cast<int64_t>(row) * p.ldc;
#pragma unroll 1
  for (int c = half; c < n_chunks; c += EPI_WARPS / 4) {
    const int col = n_blk * BN + c * 32;
    const int ncols = min(32, p.N - col);          // multiple of 8
    uint32_t r[32];
    tmem_ld32(tmem_acc + c * 32, r);
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)                    // the same address in every lane: one broadcast transaction each
      b4[j] = (has_bias && j * 4 < ncols) ? __ldg(reinterpret_cast<const float4*>(ep.bias + col) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    tmem_ld_wait();
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[4] = {__uint_as_float(r[4 * j]) + b4[j].x, __uint_as_float(r[4 * j + 1]) + b4[j].y,
                    __uint_as_float(r[4 * j + 2]) + b4[j].z, __uint_as_float(r[4 * j + 3]) + b4[j].w};
      if (RELU_DROP) {
#pragma unroll
        for (int t = 0; t < 4; t += 2) {
          const uint32_t bits = dropout_pair(key, col + 4 * j + t);
          v[t] = ((bits & 0xFFFFu) >= ep.drop_thresh) ? fmaxf(v[t], 0.f) * ep.drop_scale : 0.f;
          v[t + 1] = ((bits >> 16) >= ep.drop_thresh) ? fmaxf(v[t + 1], 0.f) * ep.drop_scale : 0.f;
        }
      }
      w[2 * j] = pack2<T>(v[0], v[1]);
      w[2 * j + 1] = pack2<T>(v[2], v[3]);
    }
    if constexpr (!STAGED) {
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q * 8 < ncols) *reinterpret_cast<uint4*>(crow + col + q * 8) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      }
    } else {
      // 32 rows x 64 bytes; 16-byte slot q of row r lives at r * 64 + ((q ^ (r >> 1)) & 3) * 16: conflict-free for the
      // row-per-lane writes and for the (8 rows x 4 slots)-per-instruction reads
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = stg_s + lane * 64 + (((q ^ (lane >> 1)) & 3) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
      }
      __syncwarp();
      const int q = lane & 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + (lane >> 2);
        uint4 v;
        const uint32_t a = stg_s + rr * 64 + (((q ^ (rr >> 1)) & 3) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
        const int grow = m_blk * BM + quarter * 32 + rr;
        if (grow < p.M && q * 8 < ncols)
          *reinterpret_cast<uint4*>(static_cast<T*>(p.C) + static_cast<int64_t>(grow) * p.ldc + col + q * 8) = v;
      }
      __syncwarp();
    }
  }
}

template <typename T, int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, float* stg, uint32_t tmem_acc, int m_blk, int n_blk,
                                              int quarter, int half, int lane) {
  const uint32_t stg_s = smem_u32(stg);
#define ST_EPI(F, OUT) epilogue_fast<T, OUT, BN, F>(p, stg_s, tmem_acc, m_blk, n_blk, quarter, half, lane)
  if constexpr (!Elem<T>::k16) {
    switch (p.flavour) {   // warp-uniform
      case EPI_PLAIN:           ST_EPI(EPI_PLAIN, float); break;
      case EPI_ROUND:           ST_EPI(EPI_ROUND, float); break;
      case EPI_AUX_ADD:         ST_EPI(EPI_AUX_ADD, float); break;
      case EPI_RELU_DROP_ROUND: ST_EPI(EPI_RELU_DROP_ROUND, float); break;
      case EPI_RELU_DROP:       ST_EPI(EPI_RELU_DROP, float); break;
      case EPI_AUX_MASK_ROUND:  ST_EPI(EPI_AUX_MASK_ROUND, float); break;
      case EPI_ATOMIC:          ST_EPI(EPI_ATOMIC, float); break;
      default:                  epilogue_generic<T, BN>(p, stg, tmem_acc, m_blk, n_blk, quarter, half, lane); break;
    }
  } else {   // 16-bit operands: the host maps the *_ROUND flavours onto these and sends the other mixes to the generic routine
    if (p.rows16) {            // warp-uniform; host: 16-bit C, 16-byte rows, no aux / column sums / amax
      if (p.rows16 == 1) {
        if (p.flavour == EPI_PLAIN) epilogue_rows16<T, BN, false, false>(p, stg_s, tmem_acc, m_blk, n_blk, quarter, half, lane);
        else epilogue_rows16<T, BN, true, false>(p, stg_s, tmem_acc, m_blk, n_blk, quarter, half, lane);
      } else {
        if (p.flavour == EPI_PLAIN) epilogue_rows16<T, BN, false, true>(p, stg_s, tmem_acc, m_blk, n_blk, quarter, half, lane);
        else epilogue_rows16<T, BN, true, true>(p, stg_s, tmem_acc, m_blk, n_blk, quarter, half, lane);
      }
      return;
    }
    switch (p.flavour) {
      case EPI_PLAIN:          if (p.c_lp) ST_EPI(EPI_PLAIN, T); else ST_EPI(EPI_PLAIN, float); break;
      case EPI_AUX_ADD:        if (p.c_lp) ST_EPI(EPI_AUX_ADD, T); else ST_EPI(EPI_AUX_ADD, float); break;
      case EPI_RELU_DROP:      ST_EPI(EPI_RELU_DROP, T); break;
      case EPI_AUX_MASK_ROUND: ST_EPI(EPI_AUX_MASK_ROUND, T); break;
      case EPI_ATOMIC:         ST_EPI(EPI_ATOMIC, float); break;
      default:                 epilogue_generic<T, BN>(p, stg, tmem_acc, m_blk, n_blk, quarter, half, lane); break;
    }
  }
#undef ST_EPI
}

// TMA loads of one 128-byte-wide k-block of an operand tile with `rows` MN rows starting at mn0
//   K-major : ONE box {ROW elements of K, rows}
//   MN-major: rows / ROW boxes {ROW elements of MN, MN_BOX_ROWS rows of K}, each MN_BOX_ROWS * 128 bytes
template <typename T, bool MN, int ROWS, typename LoadFn>
__device__ __forceinline__ void load_operand(uint8_t* dst, int kb, int mn0, LoadFn&& ld) {
  using E = Elem<T>;
  if (!MN) {
    ld(dst, kb * E::ROW, mn0);
  } else {
#pragma unroll
    for (int i = 0; i < ROWS / E::ROW; ++i) ld(dst + i * (E::MN_BOX_ROWS * 128), mn0 + i * E::ROW, kb * E::ROW);
  }
}

// ================================================================================ 1-CTA MMA (optionally clustered)
// CL = CTAs per cluster (1 or 2).  With CL == 2 the two CTAs of a cluster own vertically adjacent output tiles
// (same n_blk), each TMA-loads half of the shared B tile and multicasts it to both; a smem stage is then released
// by BOTH consumers (multicast tcgen05.commit).
template <typename T, int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  using E = Elem<T>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int MN_BOX = E::MN_BOX_ROWS * 128;   // bytes of one MN-major TMA box

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the prologue above touched only shared memory, TMEM and kernel parameters
  pdl_trigger();

  // work units: CL vertically adjacent tiles; CTA `crank` of the cluster takes tile m = unit_m * CL + crank
  const int crank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_units = (p.m_tiles + CL - 1) / CL;
  const int tiles_mn = m_units * p.n_tiles;
  const int num_tiles = tiles_mn * p.k_splits;
  const int unit0 = blockIdx.x / CL, unit_stride = gridDim.x / CL;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Converged warp; one lane chosen by elect.sync issues (under a `lane == 0` branch the compiler wraps every
    // uniform-datapath instruction — UTMALDG, UTCHMMA — in an ELECT / BRA.U.ANY loop, 60-120 cycles each).
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
        const int n_blk = tile % p.n_tiles;
        const int m_blk = ((tile / p.n_tiles) % m_units) * CL + crank;
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            uint64_t* fb = &full_bar[stage];
            mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
            load_operand<T, A_MN, BM>(sa, kb, m_blk * BM, [&](uint8_t* d, int c0, int c1) { tma_load_2d(d, &tmap_a, fb, c0, c1); });
            if (CL == 1) {
              load_operand<T, B_MN, BN>(sb, kb, n_blk * BN, [&](uint8_t* d, int c0, int c1) { tma_load_2d(d, &tmap_b, fb, c0, c1); });
            } else {  // this CTA fetches its half of the B tile for the whole cluster
              constexpr uint16_t kAll = (1u << CL) - 1;
              uint8_t* half_dst = sb + crank * (B_MN ? (BN / CL / E::ROW) * MN_BOX : (BN / CL) * 128);
              load_operand<T, B_MN, BN / CL>(half_dst, kb, n_blk * BN + crank * (BN / CL),
                                             [&](uint8_t* d, int c0, int c1) { tma_load_2d_mc(d, &tmap_b, fb, c0, c1, kAll); });
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp + elect.sync) =====================
    {
      constexpr uint32_t idesc = umma_idesc<T>(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            // K-major: +32 B per MMA inside the 128 B swizzle row.  MN-major: the next K slice of the tile (+MN_K_ADV);
            // MN groups are one TMA box apart.  Descriptors = base + constant (no re-encoding).
            const uint64_t a0 = A_MN ? umma_desc_mn<T>(sa, MN_BOX) : umma_desc_kmajor(sa);
            const uint64_t b0 = B_MN ? umma_desc_mn<T>(sb, MN_BOX) : umma_desc_kmajor(sb);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<T>(d_tmem, a0 + static_cast<uint64_t>((A_MN ? k * E::MN_K_ADV : k * 32) >> 4),
                         b0 + static_cast<uint64_t>((B_MN ? k * E::MN_K_ADV : k * 32) >> 4), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            // frees the smem stage once these MMAs retire (in every CTA whose TMA writes into it)
            if (CL > 1) umma_commit_mc(&empty_bar[stage], (1u << CL) - 1); else umma_commit(&empty_bar[stage]);
            if (kb + 1 == kb1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    float* stg = epi_smem + (warp - 2) * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int n_blk = tile % p.n_tiles;
      const int m_blk = ((tile / p.n_tiles) % m_units) * CL + crank;
      if (p.flavour == EPI_AUX_ADD || p.flavour == EPI_AUX_MASK_ROUND) prefetch_aux_tile<T, BN>(p, m_blk, n_blk, quarter, half, lane);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<T, BN>(p, stg, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN, m_blk, n_blk, quarter,
                           half, lane);
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA may exit while a peer can still write into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ================================================================================ CTA-pair MMA (cta_group::2)
// One 256(M) x 256(N) output tile per CTA pair.  CTA r stages A rows [r*128, r*128+128) and B rows (= output
// columns) [r*128, r*128+128) of every k-block; the leader (rank 0) issues tcgen05.mma.cta_group::2 with M = 256,
// which reads both CTAs' smem and accumulates each CTA's 128 rows x 256 columns into that CTA's TMEM.
// Per k-block each SM ingests 32 KB instead of 48 KB for the same MACs.
constexpr int S2_STAGES = 6;
constexpr int S2_STAGE_BYTES = 2 * A_STAGE_BYTES;   // A 16 KB + half of B 16 KB
constexpr int S2_SMEM_BYTES = S2_STAGES * S2_STAGE_BYTES + EPI_BYTES + 1024 + 512;

// ---- cluster launch control (sm_100): dynamic tile scheduling for the CTA-pair kernel.  The grid holds one cluster per
// tile; a resident cluster, instead of exiting after its tile, CANCELS a cluster that has not been launched yet and
// computes that cluster's tile.  Tiles are therefore dealt out as SMs become free: a pair that starts late — behind a
// concurrent kernel of another stream (side streams, NCCL's all-reduce CTAs) — takes fewer tiles instead of holding up the
// launch with a fixed share.  The 16-byte responses go through a 4-deep ring in shared memory, written into BOTH CTAs of
// the pair by one multicast request of the leader's producer warp and consumed by every role that walks the tile sequence
// (producer warps and epilogue warps of both CTAs, the leader's MMA warp: 19 arrivals per slot on the leader's barrier).
struct ClcRing {
  uint4 resp[4];
  uint64_t full[4];    // response k has landed in this CTA (transaction barrier, armed by this CTA's producer warp)
  uint64_t empty[4];   // leader's copy: every consumer of both CTAs has read the slot's previous response
};
constexpr int CLC_CONSUMERS = (2 + EPI_WARPS) + (1 + EPI_WARPS);

__device__ __forceinline__ void clc_try_cancel_pair(void* resp, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
               ::"r"(smem_u32(resp)), "r"(smem_u32(bar)) : "memory");
}
// the tile fetched as number k (k = 1, 2, ...; tile 0 is the cluster's own), or -1 when no cluster was left to cancel
__device__ __forceinline__ int clc_next(ClcRing* ring, int k, int lane) {
  const int slot = k & 3;
  mbar_wait(&ring->full[slot], static_cast<uint32_t>((k - 1) >> 2) & 1u);
  uint32_t valid, x, y, z;
  asm volatile(
      "{\n\t"
      ".reg .pred p1;\n\t"
      ".reg .b128 r;\n\t"
      "ld.shared.b128 r, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %3, 1, 0, p1;\n\t"
      "mov.u32 %0, 0;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t"
      "}"
      : "=r"(x), "=r"(y), "=r"(z), "=r"(valid)
      : "r"(smem_u32(&ring->resp[slot]))
      : "memory");
  fence_proxy_async_smem();   // this (generic-proxy) read precedes the async-proxy write of the slot's next response
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(&ring->empty[slot], 0);
  return valid ? static_cast<int>(x >> 1) : -1;
}

template <typename T, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_2sm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  using E = Elem<T>;
  constexpr int BN = 256;
  constexpr int STAGES = S2_STAGES;
  constexpr int MN_BOX = E::MN_BOX_ROWS * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * S2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S2_STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  both CTAs' TMA -> leader's MMA (only the leader's copy is used)
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  leader's MMA -> each CTA's producer
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       leader's MMA -> each CTA's epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]       both CTAs' epilogues -> leader's MMA (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  ClcRing* ring = reinterpret_cast<ClcRing*>(bars + 32);   // 256 bytes behind the barriers, 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = crank == 0;
  const bool clc = p.clc != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS * 32);   // the epilogue threads of both CTAs
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&ring->full[s], 1);
      mbar_init(&ring->empty[s], CLC_CONSUMERS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the prologue above touched only shared memory, TMEM and kernel parameters
  pdl_trigger();

  const int m_units = (p.m_tiles + 1) / 2;
  const int tiles_mn = m_units * p.n_tiles;
  const int num_tiles = tiles_mn * p.k_splits;
  const int unit0 = blockIdx.x / 2, unit_stride = gridDim.x / 2;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp + elect.sync) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      int k = 0;
      for (int tile = unit0; tile >= 0 && tile < num_tiles; tile = clc ? clc_next(ring, k, lane) : tile + unit_stride) {
        ++k;   // the tile AFTER this one is number k: request it now, one tile ahead of its use
        if (clc && elect_one()) {
          const int slot = k & 3;
          if (leader && k >= 5) mbar_wait(&ring->empty[slot], static_cast<uint32_t>(((k - 1) >> 2) - 1) & 1u);
          mbar_arrive_expect_tx(&ring->full[slot], 16);
          if (leader) clc_try_cancel_pair(&ring->resp[slot], &ring->full[slot]);
        }
        __syncwarp();
        const int n_blk = tile % p.n_tiles;
        const int m_blk = ((tile / p.n_tiles) % m_units) * 2 + crank;
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S2_STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            uint64_t* fb = &full_bar[stage];
            if (leader) mbar_arrive_expect_tx(fb, 2 * S2_STAGE_BYTES);   // bytes of BOTH CTAs
            load_operand<T, A_MN, BM>(sa, kb, m_blk * BM, [&](uint8_t* d, int c0, int c1) { tma_load_2d_2sm(d, &tmap_a, fb, c0, c1); });
            load_operand<T, B_MN, 128>(sb, kb, n_blk * BN + crank * 128, [&](uint8_t* d, int c0, int c1) { tma_load_2d_2sm(d, &tmap_b, fb, c0, c1); });
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc<T>(256, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int k = 0;
      for (int tile = unit0; tile >= 0 && tile < num_tiles; tile = clc ? clc_next(ring, k, lane) : tile + unit_stride) {
        ++k;
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S2_STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            const uint64_t a0 = A_MN ? umma_desc_mn<T>(sa, MN_BOX) : umma_desc_kmajor(sa);
            const uint64_t b0 = B_MN ? umma_desc_mn<T>(sb, MN_BOX) : umma_desc_kmajor(sb);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_2sm<T>(d_tmem, a0 + static_cast<uint64_t>((A_MN ? k * E::MN_K_ADV : k * 32) >> 4),
                             b0 + static_cast<uint64_t>((B_MN ? k * E::MN_K_ADV : k * 32) >> 4), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_2sm_mc(&empty_bar[stage], 3);   // release the stage in both CTAs
            if (kb + 1 == kb1) umma_commit_2sm_mc(&tfull_bar[acc], 3);   // accumulators complete -> both epilogues
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = epi_smem + (warp - 2) * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    int k = 0;
    for (int tile = unit0; tile >= 0 && tile < num_tiles; tile = clc ? clc_next(ring, k, lane) : tile + unit_stride) {
      ++k;
      const int n_blk = tile % p.n_tiles;
      const int m_blk = ((tile / p.n_tiles) % m_units) * 2 + crank;
      if (p.flavour == EPI_AUX_ADD || p.flavour == EPI_AUX_MASK_ROUND) prefetch_aux_tile<T, BN>(p, m_blk, n_blk, quarter, half, lane);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<T, BN>(p, stg, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN, m_blk, n_blk, quarter,
                           half, lane);
      tc_fence_before();
      mbar_arrive_cluster(&tempty_bar[acc], 0);   // the leader's MMA warp owns the accumulator hand-back barrier
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ================================================================================ host side
// profile tag bit fields: bn/64 [0,4)  variant [4,8)  flavour [8,12)  mode [12,14)  dtype [14,16)  K [16,34)  N [34,48)  M [48,63)... M, K in units of 8
template <typename T>
long long gemm_tag(const GemmParams& p, bool a_mn, bool b_mn, int bn, int variant) {
  const long long mode = a_mn ? 2 : (b_mn ? 1 : 0);
  const long long dt = Elem<T>::k16 ? (Elem<T>::FMT == 0 ? 1 : 2) : 0;
  return (bn / 64) | (static_cast<long long>(variant) << 4) | (static_cast<long long>(p.flavour) << 8) | (mode << 12) | (dt << 14) |
         (static_cast<long long>(p.K / 8) << 16) | (static_cast<long long>(p.N) << 34) | (static_cast<long long>(p.M / 8) << 48);
}

template <typename Kern>
int launch_clustered(Kern kern, int grid, int smem, int cluster, cudaStream_t stream, const CUtensorMap& ta,
                     const CUtensorMap& tb, const GemmParams& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_allowed(stream);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  ST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return ST_OK;
}

template <typename T, int BN, bool A_MN, bool B_MN, int CL>
int launch_gemm(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<T, BN, A_MN, B_MN, CL>;
  static bool attr_set = false;  // per instantiation; benign race (idempotent)
  if (!attr_set) {
    ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int units = ((p.m_tiles + CL - 1) / CL) * p.n_tiles * p.k_splits;
  int cap = get_option("gemm_max_ctas");
  if (cap <= 0) cap = num_sms();
  cap = cap / CL > 0 ? cap / CL : 1;
  const int grid = (units < cap ? units : cap) * CL;
  ProfScope prof(stream, PROF_GEMM, 2.0 * p.M * static_cast<double>(p.N) * p.K, gemm_tag<T>(p, A_MN, B_MN, BN, CL));
  if (CL == 1) {
    ST_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, p));
  } else {
    ST_TRY(launch_clustered(kern, grid, Cfg::SMEM_BYTES, CL, stream, ta, tb, p));
  }
  ST_CHECK_LAUNCH();
  return ST_OK;
}

template <typename T, bool A_MN, bool B_MN>
int launch_gemm_2sm(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p) {
  auto kern = gemm_2sm_kernel<T, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM_BYTES));
    attr_set = true;
  }
  const int units = ((p.m_tiles + 1) / 2) * p.n_tiles * p.k_splits;
  int cap = get_option("gemm_max_ctas");
  if (cap <= 0) cap = num_sms();
  cap = cap / 2 > 0 ? cap / 2 : 1;
  GemmParams q = p;
  q.clc = (get_option("gemm_clc") && units > cap) ? 1 : 0;     // more tiles than resident pairs: deal them out dynamically
  const int grid = (q.clc ? units : (units < cap ? units : cap)) * 2;
  ProfScope prof(stream, PROF_GEMM, 2.0 * p.M * static_cast<double>(p.N) * p.K, gemm_tag<T>(p, A_MN, B_MN, 256, 3));
  ST_TRY(launch_clustered(kern, grid, S2_SMEM_BYTES, 2, stream, ta, tb, q));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

template <typename T, int BN, int CL>
int dispatch_mode(cudaStream_t stream, GemmMode mode, const CUtensorMap& ta, const CUtensorMap& tb,
                  const GemmParams& p) {
  switch (mode) {
    case GEMM_NT: return launch_gemm<T, BN, false, false, CL>(stream, ta, tb, p);
    case GEMM_NN: return launch_gemm<T, BN, false, true, CL>(stream, ta, tb, p);
    case GEMM_TN: return launch_gemm<T, BN, true, true, CL>(stream, ta, tb, p);
  }
  set_error("gemm: bad mode %d", static_cast<int>(mode));
  return ST_ERR_INVALID;
}

// Tile width: 256 when that still gives every SM a tile; otherwise narrower tiles spread a small problem
// (decoder-side GEMMs with M = B*L ~ 1600 rows) over more SMs.
inline int pick_bn(int M, int N, int k_splits) {
  if (N <= 64) return 64;
  const int m_tiles = (M + BM - 1) / BM;
  const int sms = num_sms();
  const int forced = get_option("gemm_bn");
  if (forced == 64 || forced == 128 || forced == 256) return (forced > 64 && N <= 64) ? 64 : forced;
  // measured (tools/gemm_small_sweep.py): the wide tile wins as soon as it gives half the SMs a tile — its MMAs amortise the
  // fixed per-instruction cost over 4x the columns (M = 1600, N = 1536 / 2048: 14.7 us against 16.8 us with 128-wide tiles)
  if (N > 128 && 2 * m_tiles * ((N + 255) / 256) * k_splits >= sms) return 256;
  if (N > 64 && m_tiles * ((N + 127) / 128) * k_splits >= sms) return 128;
  return 64;
}

// Fast (compile-time specialised, vector) epilogue for the flag combinations the composite ops use.
template <typename T>
int pick_flavour(const GemmEpilogue& ep, const void* C, int64_t ldc, int c_lp, int N) {
  constexpr bool k16 = Elem<T>::k16;
  auto al = [](const void* q, int a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  const int c_align = (k16 && c_lp) ? 8 : 16, aux_align = k16 ? 8 : 16;
  const bool vec_ok = (ldc & 3) == 0 && al(C, c_align) && (N & 3) == 0 && (!ep.aux || ((ep.ldaux & 3) == 0 && al(ep.aux, aux_align))) &&
                      (!ep.bias || al(ep.bias, 16));
  if (!vec_ok || get_option("gemm_generic_epilogue")) return EPI_GENERIC;
  const bool drop = ep.drop_thresh != 0;
  const bool round = !k16 && ep.round_tf32;   // a 16-bit output is rounded by its conversion
  if (ep.atomic) return (!ep.bias && !ep.aux_mode && !ep.relu && !drop && !round && !c_lp) ? EPI_ATOMIC : EPI_GENERIC;
  if (k16) {
    if (ep.relu) return (ep.aux_mode == 0 && c_lp) ? EPI_RELU_DROP : EPI_GENERIC;
    if (drop) return EPI_GENERIC;
    if (ep.aux_mode == 1) return EPI_AUX_ADD;
    if (ep.aux_mode == 2) return c_lp ? EPI_AUX_MASK_ROUND : EPI_GENERIC;
    return EPI_PLAIN;
  }
  if (ep.relu) return ep.aux_mode != 0 ? EPI_GENERIC : (round ? EPI_RELU_DROP_ROUND : EPI_RELU_DROP);   // thresh 0 keeps everything
  if (drop) return EPI_GENERIC;
  if (ep.aux_mode == 1) return round ? EPI_GENERIC : EPI_AUX_ADD;
  if (ep.aux_mode == 2) return round ? EPI_AUX_MASK_ROUND : EPI_GENERIC;
  return round ? EPI_ROUND : EPI_PLAIN;
}

// ALLOW_MC: also instantiate the 2-CTA multicast variant (opt-in through option "gemm_cluster" = 2; TF32 only)
template <typename T, bool ALLOW_MC>
int gemm_run(cudaStream_t stream, GemmMode mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
             int64_t ldc, int c_lp, int M, int N, int K, const GemmEpilogue& ep, int k_splits) {
  using E = Elem<T>;
  constexpr int DT = E::k16 ? (E::FMT == 0 ? 1 : 2) : 0;
  constexpr int EPV = 16 / E::BYTES;   // elements per 16 bytes
  ST_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  ST_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  ST_REQUIRE((lda % EPV) == 0 && (ldb % EPV) == 0, "gemm: lda=%lld ldb=%lld must be multiples of %d elements",
             (long long)lda, (long long)ldb, EPV);
  ST_REQUIRE(k_splits >= 1 && (k_splits == 1 || ep.atomic), "gemm: split-K needs the atomic epilogue");
  ST_REQUIRE(!(ep.atomic && c_lp), "gemm: split-K accumulation needs an fp32 output");

  GemmParams p;
  p.C = C; p.ldc = ldc; p.c_lp = (E::k16 && c_lp) ? 1 : 0; p.M = M; p.N = N; p.K = K;
  p.kblocks_total = (K + E::ROW - 1) / E::ROW;
  if (k_splits > p.kblocks_total) k_splits = p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + k_splits - 1) / k_splits;
  p.k_splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  const int BN = pick_bn(M, N, p.k_splits);
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.ep = ep;
  p.flavour = pick_flavour<T>(ep, C, ldc, p.c_lp, N);
  p.rows16 = 0;
  p.clc = 0;
  if (E::k16 && p.c_lp && (p.flavour == EPI_PLAIN || p.flavour == EPI_RELU_DROP) && !ep.colsum && !ep.amax_out && (N & 7) == 0 &&
      (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (!ep.bias || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0))
    p.rows16 = get_option("gemm_rows16");
  float* deferred_colsum = nullptr;
  if (ep.colsum && (p.flavour == EPI_GENERIC || p.flavour == EPI_ATOMIC || (reinterpret_cast<uintptr_t>(ep.colsum) & 15))) {
    deferred_colsum = ep.colsum;   // the generic epilogue has no fused column sums: separate pass below
    p.ep.colsum = nullptr;
  }
  if (deferred_colsum) {
    ST_REQUIRE(k_splits == 1 && !ep.atomic && !p.c_lp, "gemm: column sums of a split-K / 16-bit output need the fast epilogue");
  }
  // Variant: 0 = independent CTAs, 2 = 2-CTA cluster with multicast B (opt-in: measured slower), 3 = CTA-pair MMA.
  // The pair MMA is used when the grid is full anyway (>= 2 tiles per SM): then bytes per SM, not latency, limit.
  int variant = 0;
  const bool pairable = (BN == 256 && p.m_tiles >= 2);
  if (pairable && p.m_tiles * p.n_tiles * p.k_splits >= 2 * num_sms()) variant = 3;
  const int forced = get_option("gemm_cluster");
  if (forced == 1) variant = 0;
  if ((forced == 2 || forced == 3) && pairable) variant = forced;
  if (!ALLOW_MC && variant == 2) variant = 3;
  const int b_rows = (variant == 2 || variant == 3) ? BN / 2 : BN;

  CUtensorMap ta, tb;
  {
    // A: K-major -> dims {K, M}; MN-major (TN) -> stored [K, M], dims {M, K}
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(lda) * E::BYTES};
    uint32_t box[2];
    if (mode == GEMM_TN) { dims[0] = M; dims[1] = K; box[0] = E::ROW; box[1] = E::MN_BOX_ROWS; }
    else                 { dims[0] = K; dims[1] = M; box[0] = E::ROW; box[1] = BM; }
    ST_TRY(make_tmap(&ta, DT, A, 2, dims, strides, box, (!E::k16 && mode == GEMM_TN) ? 1 : 0));
  }
  {
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(ldb) * E::BYTES};
    uint32_t box[2];
    if (mode == GEMM_NT) { dims[0] = K; dims[1] = N; box[0] = E::ROW; box[1] = static_cast<uint32_t>(b_rows); }
    else                 { dims[0] = N; dims[1] = K; box[0] = E::ROW; box[1] = E::MN_BOX_ROWS; }
    ST_TRY(make_tmap(&tb, DT, B, 2, dims, strides, box, (!E::k16 && mode != GEMM_NT) ? 1 : 0));
  }
  int status = ST_ERR_INVALID;
  if (variant == 3) {
    switch (mode) {
      case GEMM_NT: status = launch_gemm_2sm<T, false, false>(stream, ta, tb, p); break;
      case GEMM_NN: status = launch_gemm_2sm<T, false, true>(stream, ta, tb, p); break;
      case GEMM_TN: status = launch_gemm_2sm<T, true, true>(stream, ta, tb, p); break;
    }
  } else {
    switch (BN) {
      case 256:
        if constexpr (ALLOW_MC) {
          status = variant == 2 ? dispatch_mode<T, 256, 2>(stream, mode, ta, tb, p) : dispatch_mode<T, 256, 1>(stream, mode, ta, tb, p);
        } else {
          status = dispatch_mode<T, 256, 1>(stream, mode, ta, tb, p);
        }
        break;
      case 128: status = dispatch_mode<T, 128, 1>(stream, mode, ta, tb, p); break;
      default:  status = dispatch_mode<T, 64, 1>(stream, mode, ta, tb, p); break;
    }
  }
  if (status == ST_OK && deferred_colsum) status = colsum_add(stream, static_cast<const float*>(C), ldc, M, N, deferred_colsum);
  return status;
}

}  // namespace

}  // namespace st
