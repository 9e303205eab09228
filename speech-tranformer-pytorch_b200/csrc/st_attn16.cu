// st_attn16.cu — the attention core with 16-bit operands (fp16 or bf16 on tcgen05 kind::f16, fp32 accumulation, fp32
// softmax statistics).  Same mathematics as st_attn.cu / st_attn_bwd.cu (transformer/Attention.py:78-90 and its autograd
// backward); what the narrower operands change:
//
//   * one MMA covers K = 16 elements (32 bytes) instead of 8: half the tcgen05 instructions for the same tile — the TF32
//     kernels are bound by instruction count at N = d_k = 64;
//   * a [rows][64] tile of 16-bit elements under the plain 128-byte swizzle is BOTH a valid K-major operand (d_k is the
//     contraction dimension: Q K^T, dO V^T) and a valid MN-major operand (the rows are the contraction dimension: P V,
//     dS K, P^T dO, dS^T Q), so every streamed tile is fetched ONCE (the TF32 kernels need a second copy in the
//     32-byte-atom swizzle): 16 KB per 64-row step instead of 48-64 KB;
//   * probabilities / score gradients go back to TMEM as packed pairs (two K elements per 32-bit column), each thread into
//     the first half of the columns it has just read, so no thread overwrites columns another one still has to read;
//   * resident 128-row tiles (Q, dO / K, V) are copied to TMEM as they lie in memory (a 64-element row = 32 columns);
//   * fp16 only: dS is multiplied by a power of two before rounding (AttnDev::ds_boost, undone in the epilogue) so that
//     small score gradients stay in fp16's normal range.
//
//   forward : CTA per (128-query tile, head, batch), 2 threads per row, 2 CTAs per SM, online softmax with one sweep
//   dQ      : CTA per 128-query tile streaming 64-key tiles through a TMA ring (producer warp / MMA warp / 16 compute warps)
//   dK, dV  : CTA per 128-key tile streaming 64-query tiles, same organisation
#include <math.h>
#include <utility>

#include "st_attn.cuh"

namespace st {

namespace {

template <int V> struct IC { static constexpr int value = V; };
template <class F, int... I>
__device__ __forceinline__ void cfor_impl(F& f, std::integer_sequence<int, I...>) { (f(IC<I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void cfor(F& f) { cfor_impl(f, std::make_integer_sequence<int, N>{}); }

template <typename T> constexpr int dtype_of() { return Elem<T>::FMT == 0 ? ST_DTYPE_F16 : ST_DTYPE_BF16; }

__device__ __forceinline__ float chunk_max32(const uint32_t (&r)[32], uint32_t mb) {
  float mx = -INFINITY;
  if (mb == 0u) {
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!((mb >> i) & 1u)) mx = fmaxf(mx, __uint_as_float(r[i]));
  }
  return mx;
}

// out <- packed T(keep ? exp2(r*scale_log2 - m) : 0) with masked entries 0; returns the (pre-dropout) row-sum part.
template <typename T, bool MASK, bool DROP>
__device__ __forceinline__ float probs32_t(const uint32_t (&r)[32], uint32_t (&out)[16], uint32_t mb, float scale_log2, float m_use,
                                           uint32_t thresh, uint32_t rowkey, const uint32_t* ckey) {
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    uint4 ck = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) ck = *reinterpret_cast<const uint4*>(ckey + i);
    const uint32_t cks[4] = {ck.x, ck.y, ck.z, ck.w};
    float pv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float x = fast_exp2(fmaf(__uint_as_float(r[i + t]), scale_log2, -m_use));
      if (MASK && ((mb >> (i + t)) & 1u)) x = 0.f;
      l += x;
      if (DROP && !dropout_keep_xor(rowkey, cks[t], thresh)) x = 0.f;
      pv[t] = x;
    }
    out[i / 2] = pack2<T>(pv[0], pv[1]);
    out[i / 2 + 1] = pack2<T>(pv[2], pv[3]);
  }
  return l;
}
template <typename T>
__device__ __forceinline__ float probs32(const uint32_t (&r)[32], uint32_t (&out)[16], uint32_t mb, float scale_log2, float m_use,
                                         uint32_t thresh, uint32_t rowkey, const uint32_t* ckey) {
  if (mb == 0u) {
    return thresh ? probs32_t<T, false, true>(r, out, mb, scale_log2, m_use, thresh, rowkey, ckey)
                  : probs32_t<T, false, false>(r, out, mb, scale_log2, m_use, thresh, rowkey, ckey);
  }
  return thresh ? probs32_t<T, true, true>(r, out, mb, scale_log2, m_use, thresh, rowkey, ckey)
                : probs32_t<T, true, false>(r, out, mb, scale_log2, m_use, thresh, rowkey, ckey);
}

// store 2*NW consecutive T elements held as NW packed words (NW a multiple of 4, dst 16-byte aligned)
template <int NW>
__device__ __forceinline__ void store_words(void* dst, const uint32_t (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW; i += 4) reinterpret_cast<uint4*>(dst)[i / 4] = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
}

// ================================================================================ forward
template <typename T, int DK>
__global__ void __launch_bounds__(256, 2)
attn16_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const AttnDev p) {
  constexpr int BQ = 128, BKV = 128;
  constexpr int G = DK / 64;                // 64-column groups (one 128-byte swizzle row each)
  constexpr int Q_BYTES = BQ * DK * 2;
  constexpr int KV_BYTES = BKV * DK * 2;
  constexpr uint32_t TCOLS = 256;           // S / packed P at [0, BKV), O at [BKV, BKV + DK)
  constexpr int SH = BKV / 2;               // score columns per thread
  constexpr int NCH = SH / 32;              // 32-column chunks per thread
  constexpr int OH = DK / 2;                // output columns per thread
  static_assert(BKV + DK <= 256 && (DK == 64 || DK == 128), "tile configuration");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  __shared__ uint64_t bar_q, bar_k, bar_v, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_part[2][BQ];
  __shared__ __align__(16) uint32_t s_ckey[BKV];
  __shared__ uint32_t s_mb[2][BKV / 32];    // mask bits of the current / next key tile when every row shares them
  __shared__ int s_extent;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int rit = quarter * 32 + lane;      // row in tile
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + rit;
  const bool row_ok = row < p.Lq;
  const int n_kv_all = (p.Lk + BKV - 1) / BKV;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(&bar_q, 1); mbar_init(&bar_k, 1); mbar_init(&bar_v, 1); mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  pdl_wait();      // nothing above reads or writes global memory (programmatic dependent launch, st_host.h)
  pdl_trigger();

  const int extent = block_key_extent(p, b, &s_extent);
  // key tiles entirely inside this utterance's padding contribute exactly nothing: skipped (at least one tile is always
  // processed so that a fully masked row still produces the reference's NaN); with a causal mask also the tiles above
  // the diagonal of this query tile
  int n_kv = extent >= p.Lk ? n_kv_all : max(1, (extent + BKV - 1) / BKV);
  if (p.causal) n_kv = min(n_kv, (min(q0 + BQ, p.Lq) + BKV - 1) / BKV);
  const bool shared_mask = mask_is_row_invariant(p);
  if (shared_mask && tid < BKV / 32) s_mb[0][tid] = mask_bits_row(p, b, 0, true, tid * 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
  constexpr uint32_t T_S = 0, T_O = BKV;

  uint32_t k_loads = 0, s_count = 0;  // phase counters for bar_k / bar_s (warp 0 issues, all wait)

  auto load_k = [&](int j) {
    mbar_arrive_expect_tx(&bar_k, KV_BYTES);
    tma_load_4d(sK, &tmap_k, &bar_k, 0, j * BKV, h * G, b);
  };
  auto load_v = [&](int j) {
    mbar_arrive_expect_tx(&bar_v, KV_BYTES);
    tma_load_4d(sV, &tmap_v, &bar_v, 0, j * BKV, h * G, b);
  };
  auto issue_s = [&]() {  // S = Q K^T  (both K-major)
    constexpr uint32_t idesc = umma_idesc<T>(128, BKV, false, false);
    const uint64_t aq0 = umma_desc_kmajor(smem_u32(sQ)), bk0 = umma_desc_kmajor(smem_u32(sK));
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks)
      umma_f16_ss(tmem + T_S, aq0 + static_cast<uint64_t>(((ks / 4) * (BQ * 128) + (ks % 4) * 32) >> 4),
                  bk0 + static_cast<uint64_t>(((ks / 4) * (BKV * 128) + (ks % 4) * 32) >> 4), idesc, ks > 0 ? 1u : 0u);
    umma_commit(&bar_s);
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_q, Q_BYTES);
      tma_load_4d(sQ, &tmap_q, &bar_q, 0, q0, h * G, b);
      load_k(0);
      load_v(0);
    }
    __syncwarp();
    mbar_wait(&bar_q, 0);
    mbar_wait(&bar_k, 0);
    tc_fence_after();
    if (elect_one()) issue_s();
    __syncwarp();
  }
  k_loads = 1;

  // Online softmax with ONE read of S per tile; O accumulates in TMEM in units of exp2(-m_ref), the reference maximum only
  // advances (and O / l are rescaled) when a tile exceeds it by more than TAU — see st_attn.cu.
  constexpr float TAU = 8.f;
  float m_ref = -INFINITY, l_run = 0.f;
  const uint64_t rng_row = (static_cast<uint64_t>(b) * p.H + h) * p.Lq + (row_ok ? row : 0);
  const uint32_t drop_key = p.drop_thresh ? dropout_row_key(p.drop_seed, rng_row) : 0u;

  for (int j = 0; j < n_kv; ++j) {
    mbar_wait(&bar_s, s_count & 1);   // S_j complete; the tensor pipe is in order, so P·V(j-1) has completed too
    ++s_count;
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        if (j + 1 < n_kv) load_k(j + 1);   // S_j has consumed K_j
        if (j > 0) load_v(j);              // P·V(j-1) has consumed V_{j-1}
      }
      __syncwarp();
    }
    if (p.drop_thresh && tid < BKV) s_ckey[tid] = dropout_col_key(p.drop_seed, static_cast<uint32_t>(j * BKV + tid));
    if (shared_mask && tid < BKV / 32 && j + 1 < n_kv)
      s_mb[(j + 1) & 1][tid] = mask_bits_row(p, b, 0, true, (j + 1) * BKV + tid * 32);

    uint32_t mbits[NCH];
    uint32_t r[NCH][32];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col0 = half * SH + c * 32;
      mbits[c] = shared_mask ? s_mb[j & 1][col0 / 32] : mask_bits_row(p, b, row, row_ok, j * BKV + col0);
      tmem_ld32(t_lane + T_S + col0, r[c]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NCH; ++c) mx = fmaxf(mx, chunk_max32(r[c], mbits[c]));
    s_part[half][rit] = mx;
    __syncthreads();
    mx = fmaxf(s_part[0][rit], s_part[1][rit]) * p.scale_log2;
    const bool need = mx > m_ref + TAU || (m_ref == -INFINITY && mx > -INFINITY);
    if (__any_sync(0xffffffffu, need)) {
      const float alpha = need ? fast_exp2(m_ref - mx) : 1.f;   // m_ref == -inf -> 0
      if (j > 0) {
#pragma unroll
        for (int c = 0; c < OH / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(t_lane + T_O + half * OH + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(need ? __uint_as_float(o[i]) * alpha : __uint_as_float(o[i]));
          tmem_st32(t_lane + T_O + half * OH + c * 32, o);
        }
      }
      if (need) { l_run *= alpha; m_ref = mx; }
    }
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    // probabilities -> TMEM as packed pairs, into the first 16 columns of each 32-column chunk this thread owns
    float l_tile = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col0 = half * SH + c * 32;
      uint32_t pk[16];
      l_tile += probs32<T>(r[c], pk, mbits[c], p.scale_log2, m_use, p.drop_thresh, drop_key, s_ckey + col0);
      tmem_st16(t_lane + T_S + col0, pk);
    }
    l_run += l_tile;
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {  // O += P V   (A = packed P in TMEM, B = V as an MN-major operand), then S_{j+1} right behind it
      tc_fence_after();
      mbar_wait(&bar_v, j & 1);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc<T>(128, DK, false, true);
        const uint64_t bv0 = umma_desc_mn<T>(smem_u32(sV), BKV * 128);
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)   // keys [16 ks, 16 ks + 16): packed columns (ks / 2) * 32 + (ks % 2) * 8
          umma_f16_ts(tmem + T_O, tmem + T_S + (ks / 2) * 32 + (ks % 2) * 8, bv0 + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                      (j > 0 || ks > 0) ? 1u : 0u);
        if (j + 1 == n_kv) umma_commit(&bar_o);
      }
      __syncwarp();
      if (j + 1 < n_kv) {
        mbar_wait(&bar_k, k_loads & 1);      // K_{j+1}
        tc_fence_after();
        if (elect_one()) issue_s();
        __syncwarp();
      }
    }
    if (j + 1 < n_kv) ++k_loads;
  }

  // ---- finalize: ctx = O / l (NaN for a fully masked row: 0 * inf), lse
  mbar_wait(&bar_o, 0);
  tc_fence_after();
  __syncthreads();
  s_part[half][rit] = l_run;
  __syncthreads();
  const float l_tot = s_part[0][rit] + s_part[1][rit];
  const float inv_l = (p.drop_thresh ? p.drop_scale : 1.f) / l_tot;   // inverted-dropout scale folded in here
  const float lse2 = m_ref + log2f(l_tot);
  {
    T* dst = static_cast<T*>(p.ctx) + (static_cast<int64_t>(b) * p.Lq + (row_ok ? row : 0)) * p.ldctx + h * DK + half * OH;
#pragma unroll
    for (int c = 0; c < OH / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(t_lane + T_O + half * OH + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack2<T>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
        store_words<16>(dst + c * 32, w);
      }
    }
    if (row_ok && half == 0) p.lse2[(static_cast<int64_t>(b) * p.H + h) * p.Lq + row] = lse2;
  }

  // ---- optional second sweep: materialise the (post-dropout) probabilities the module returns (fp32)
  if (p.attn != nullptr) {
    for (int j = 0; j < n_kv_all; ++j) {   // every key tile: masked columns of the returned weights are written as zeros
      tc_fence_before();
      __syncthreads();  // every thread is done with the S region / previous sweep step
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) load_k(j);
        __syncwarp();
        mbar_wait(&bar_k, k_loads & 1);
        tc_fence_after();
        if (elect_one()) issue_s();
        __syncwarp();
      }
      ++k_loads;
      mbar_wait(&bar_s, s_count & 1);
      ++s_count;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int k0 = j * BKV + half * SH + c * 32;
        const uint32_t mb = mask_bits_row(p, b, row, row_ok, k0);
        uint32_t r[32];
        tmem_ld32(t_lane + T_S + half * SH + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = p.attn + ((static_cast<int64_t>(b) * p.H + h) * p.Lq + row) * p.Lk + k0;
          for (int i = 0; i < 32; ++i) {
            if (k0 + i >= p.Lk) break;
            float pv = ((mb >> i) & 1u) ? 0.f : fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -lse2));
            if (l_tot == 0.f) pv = __int_as_float(0x7fc00000);  // fully masked row: NaN like softmax(-inf row)
            if (p.drop_thresh)
              pv = dropout_keep_xor(drop_key, dropout_col_key(p.drop_seed, static_cast<uint32_t>(k0 + i)), p.drop_thresh)
                       ? pv * p.drop_scale : 0.f;
            dst[i] = pv;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ delta = rowsum(dO * O) per head
template <typename T>
__global__ void __launch_bounds__(256)
attn16_delta_kernel(const T* __restrict__ dctx, int64_t lddctx, const T* __restrict__ ctx, int64_t ldctx,
                    float* __restrict__ delta, int B, int H, int Lq, int dk) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t rows = static_cast<int64_t>(B) * Lq;
  const int d = H * dk;
  const int grp = dk / 8;  // lanes per head within a 256-column chunk (8 or 16)
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const int b = static_cast<int>(row / Lq), q = static_cast<int>(row - static_cast<int64_t>(b) * Lq);
    for (int c0 = 0; c0 < d; c0 += 256) {
      const int c = c0 + lane * 8;
      float s = 0.f;
      if (c < d) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(dctx + row * lddctx + c));
        const uint4 o = __ldg(reinterpret_cast<const uint4*>(ctx + row * ldctx + c));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = unpack2<T>(aw[i]), y = unpack2<T>(ow[i]);
          s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
        }
      }
      for (int off = 1; off < grp; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (c < d && (lane % grp) == 0) delta[(static_cast<int64_t>(b) * H + c / dk) * Lq + q] = s;
    }
  }
}

// ================================================================================ backward: shared pieces
constexpr int BT = 64;                 // streamed tile height
constexpr int NCOMP = 512;             // compute threads
constexpr int NTHREADS = NCOMP + 64;   // + producer warp + MMA warp
constexpr int W_PROD = NCOMP / 32;     // warp 16
constexpr int W_MMA = W_PROD + 1;      // warp 17

// dS chunk (dQ kernel; thread = query row, 16 key columns): out <- packed T(P * (keep ? dP*dsc : 0) - P*dlt),
// P = exp2(S*scale_log2 - lse2); dsc = dropout scale * boost, dlt = delta * boost
template <typename T, bool MASK, bool DROP>
__device__ __forceinline__ void ds16_t(const uint32_t (&rs)[16], const uint32_t (&rd)[16], uint32_t (&out)[8], uint32_t mb,
                                       float scale_log2, float lse2, float dlt, float dsc, uint32_t thresh, uint32_t rowkey,
                                       const uint32_t* ckey) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    uint4 ck = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) ck = *reinterpret_cast<const uint4*>(ckey + i);
    const uint32_t cks[4] = {ck.x, ck.y, ck.z, ck.w};
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float pr = fast_exp2(fmaf(__uint_as_float(rs[i + t]), scale_log2, -lse2));
      if (MASK && ((mb >> (i + t)) & 1u)) pr = 0.f;
      float dp = __uint_as_float(rd[i + t]);
      if (DROP && !dropout_keep_xor(rowkey, cks[t], thresh)) dp = 0.f;
      v[t] = pr * fmaf(dp, dsc, -dlt);
    }
    out[i / 2] = pack2<T>(v[0], v[1]);
    out[i / 2 + 1] = pack2<T>(v[2], v[3]);
  }
}

// P^T / dS^T chunk (dK/dV kernel; thread = key row, 16 query columns); delta[] is pre-multiplied by the boost
template <typename T, bool DENSE, bool DROP>
__device__ __forceinline__ void dkv16_t(const uint32_t (&rs)[16], const uint32_t (&rd)[16], uint32_t (&op)[8], uint32_t (&ods)[8],
                                        const float* lse, const float* delta, const uint32_t* rkey, float scale_log2, float dsc,
                                        uint32_t thresh, uint32_t my_ckey, const uint8_t* mrow, int64_t ms_q, int q_first, int Lq,
                                        bool key_ok, int causal_key) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 lse4 = *reinterpret_cast<const float4*>(lse + i);
    const float4 del4 = *reinterpret_cast<const float4*>(delta + i);
    uint4 key4 = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) key4 = *reinterpret_cast<const uint4*>(rkey + i);
    const float lses[4] = {lse4.x, lse4.y, lse4.z, lse4.w};
    const float dels[4] = {del4.x, del4.y, del4.z, del4.w};
    const uint32_t rks[4] = {key4.x, key4.y, key4.z, key4.w};
    float vp[4], vd[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float pr = fast_exp2(fmaf(__uint_as_float(rs[i + t]), scale_log2, -lses[t]));
      if (DENSE) {
        const int q = q_first + i + t;
        if (!key_ok || q < causal_key || (mrow != nullptr && q < Lq && mrow[static_cast<int64_t>(q) * ms_q] != 0)) pr = 0.f;
      }
      float dp = __uint_as_float(rd[i + t]);
      float pd = pr;
      if (DROP && !dropout_keep_xor(rks[t], my_ckey, thresh)) { pd = 0.f; dp = 0.f; }
      vp[t] = pd;
      vd[t] = pr * fmaf(dp, dsc, -dels[t]);
    }
    op[i / 2] = pack2<T>(vp[0], vp[1]);
    op[i / 2 + 1] = pack2<T>(vp[2], vp[3]);
    ods[i / 2] = pack2<T>(vd[0], vd[1]);
    ods[i / 2 + 1] = pack2<T>(vd[2], vd[3]);
  }
}

// Resident 128-row tiles -> TMEM as they lie in memory: compute slice `slice` (0..3) owns one 64-element (32-word) chunk of
// one of the two resident tensors; rows >= L are zeros.
template <int DK, typename T>
__device__ __forceinline__ void resident_load16(const T* x0, int64_t ld0, const T* x1, int64_t ld1, int64_t row, bool row_ok, int h,
                                                int slice, uint32_t (&r)[32]) {
  constexpr int CH = DK / 64;
  if (slice < 2 * CH) {                  // warp-uniform
    const int which = slice / CH, c = slice % CH;
    const T* src = (which == 0 ? x0 + row * ld0 : x1 + row * ld1) + h * DK + c * 64;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok) v = __ldg(reinterpret_cast<const uint4*>(src + 2 * i));
      r[i] = v.x; r[i + 1] = v.y; r[i + 2] = v.z; r[i + 3] = v.w;
    }
  }
}
template <int DK>
__device__ __forceinline__ void resident_store16(int slice, uint32_t t_lane, uint32_t t_x0, uint32_t t_x1, const uint32_t (&r)[32]) {
  constexpr int CH = DK / 64;
  if (slice < 2 * CH) {
    const int which = slice / CH, c = slice % CH;
    tmem_st32(t_lane + (which == 0 ? t_x0 : t_x1) + c * 32, r);
    tmem_st_wait();
  }
}

// ================================================================================ dQ
template <typename T, int DK>
__global__ void __launch_bounds__(NTHREADS, 1)
attn16_bwd_dq(const T* __restrict__ q, int64_t ldq, const T* __restrict__ dctx, int64_t lddctx,
              const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v, const AttnDev p) {
  constexpr int BQ = 128;
  constexpr int G = DK / 64;
  constexpr int STAGES = 4;
  constexpr int T_BYTES = BT * DK * 2;
  constexpr int STAGE_BYTES = 2 * T_BYTES;   // K | V  (K serves S = Q K^T as a K-major and dQ += dS K as an MN-major operand)
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_S = 0, T_DP = 2 * BT, T_DQ = 4 * BT, T_Q = 4 * BT + DK, T_DO = 4 * BT + DK + DK / 2;
  static_assert(4 * BT + 2 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t res_ready, ld_full[STAGES], ld_empty[STAGES], s_full[2], ds_full[2], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) uint32_t s_ckey[STAGES][BT];
  __shared__ uint32_t s_mb[STAGES][BT / 32];
  __shared__ int s_extent;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    mbar_init(&res_ready, NCOMP); mbar_init(&acc_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ld_full[s], 1); mbar_init(&ld_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], NCOMP); }
    fence_mbar_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  pdl_wait();      // nothing above reads or writes global memory (programmatic dependent launch, st_host.h)
  pdl_trigger();

  uint32_t rres[32];
  const int r_row = q0 + (warp & 3) * 32 + lane;
  const bool r_ok = r_row < p.Lq;
  if (warp < W_PROD)
    resident_load16<DK, T>(q, ldq, dctx, lddctx, static_cast<int64_t>(b) * p.Lq + (r_ok ? r_row : 0), r_ok, h, warp >> 2, rres);

  const int extent = block_key_extent(p, b, &s_extent);
  int n_kv = extent >= p.Lk ? (p.Lk + BT - 1) / BT : max(1, (extent + BT - 1) / BT);
  if (p.causal) n_kv = min(n_kv, (min(q0 + BQ, p.Lq) + BT - 1) / BT);   // key tiles above the diagonal have dS == 0
  const bool shared_mask = mask_is_row_invariant(p);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == W_PROD) {
    // ===================== producer =====================
    if (lane == 0) { tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); }
    for (int t = 0; t < n_kv; ++t) {
      const int s = t % STAGES;
      mbar_wait(&ld_empty[s], ((t / STAGES) & 1) ^ 1);
      if (p.drop_thresh) {
        s_ckey[s][lane] = dropout_col_key(p.drop_seed, static_cast<uint32_t>(t * BT + lane));
        s_ckey[s][lane + 32] = dropout_col_key(p.drop_seed, static_cast<uint32_t>(t * BT + lane + 32));
      }
      if (lane < BT / 32) s_mb[s][lane] = shared_mask ? mask_bits_row(p, b, 0, true, t * BT + lane * 32) : 0u;
      __syncwarp();
      if (lane == 0) {
        uint8_t* st = sStage + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&ld_full[s], STAGE_BYTES);   // release: the metadata stores above become visible with it
        tma_load_4d(st, &tmap_k, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + T_BYTES, &tmap_v, &ld_full[s], 0, t * BT, h * G, b);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (converged warp + elect.sync) =====================
    {
      const uint32_t st0 = smem_u32(sStage);
      auto koff = [](int ks, int rows) { return static_cast<uint64_t>(((ks / 4) * (rows * 128) + (ks % 4) * 32) >> 4); };
      constexpr int U = STAGES;   // tiles per unrolled round: stage and buffer indices are constants (STAGES is even)
      const uint64_t dk0 = umma_desc_kmajor(st0), dkm0 = umma_desc_mn<T>(st0, BT * 128), dv0 = umma_desc_kmajor(st0 + T_BYTES);
      auto issue_a = [&](uint32_t ld_parity, auto S, auto TB) {   // S(t) = Q K^T, dP(t) = dO V^T   (A = resident tile in TMEM)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ld_full[s], ld_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, BT, false, false);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ts(tmem + T_S + tb * BT, tmem + T_Q + ks * 8, dk0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ts(tmem + T_DP + tb * BT, tmem + T_DO + ks * 8, dv0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          umma_commit(&s_full[tb]);
        }
        __syncwarp();
      };
      auto issue_b = [&](uint32_t ds_parity, bool first, auto S, auto TB) {   // dQ += dS(t) K(t)   (A = packed dS, B = K MN-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ds_full[tb], ds_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, DK, false, true);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)   // keys [16 ks, +16): the 8 packed columns slice ks wrote at 16 ks
            umma_f16_ts(tmem + T_DQ, tmem + T_DP + tb * BT + ks * 16, dkm0 + so + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                        (!first || ks > 0) ? 1u : 0u);
          umma_commit(&ld_empty[s]);   // stage s free once everything issued so far has retired
        }
        __syncwarp();
      };
      mbar_wait(&res_ready, 0);
      tc_fence_after();
      issue_a(0u, IC<0>{}, IC<0>{});
      for (int t0 = 0; t0 < n_kv; t0 += U) {
        const uint32_t ldp = static_cast<uint32_t>(t0 / STAGES), dsp = static_cast<uint32_t>(t0 >> 1);
        auto round = [&](auto Uc) {
          constexpr int u = decltype(Uc)::value;
          const int t = t0 + u;
          if (t < n_kv) {
            if (t + 1 < n_kv) issue_a((ldp + (u + 1) / STAGES) & 1u, IC<(u + 1) % STAGES>{}, IC<(u + 1) & 1>{});
            issue_b((dsp + (u >> 1)) & 1u, t == 0, IC<u % STAGES>{}, IC<(u & 1)>{});
          }
        };
        cfor<U>(round);
      }
      if (elect_one()) umma_commit(&acc_full);
    }
    __syncwarp();
  } else {
    // ===================== compute warps =====================
    const int quarter = warp & 3, slice = warp >> 2;
    const int row = q0 + quarter * 32 + lane;
    const bool row_ok = row < p.Lq;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int col0 = slice * 16;
    const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (row_ok ? row : 0);
    resident_store16<DK>(slice, t_lane, T_Q, T_DO, rres);
    tc_fence_before();
    mbar_arrive(&res_ready);
    const float dlt = (row_ok ? p.delta[stat] : 0.f) * p.ds_boost;
    const float lse2 = row_ok ? p.lse2[stat] : INFINITY;
    const uint32_t drop_key = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(stat)) : 0u;
    const float dsc = (p.drop_thresh ? p.drop_scale : 1.f) * p.ds_boost;
    for (int t = 0; t < n_kv; ++t) {
      const int s = t % STAGES, tb = t & 1;
      mbar_wait(&s_full[tb], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t word = shared_mask ? s_mb[s][col0 >> 5] : mask_bits_row(p, b, row, row_ok, t * BT + (col0 & ~31));
      const uint32_t mb = (word >> (col0 & 31)) & 0xFFFFu;
      uint32_t rs[16], rd[16], out[8];
      tmem_ld16(t_lane + T_S + tb * BT + col0, rs);
      tmem_ld16(t_lane + T_DP + tb * BT + col0, rd);
      tmem_ld_wait();
      const uint32_t* ck = s_ckey[s] + col0;
      if (mb == 0u) {
        if (p.drop_thresh) ds16_t<T, false, true>(rs, rd, out, mb, p.scale_log2, lse2, dlt, dsc, p.drop_thresh, drop_key, ck);
        else ds16_t<T, false, false>(rs, rd, out, mb, p.scale_log2, lse2, dlt, dsc, p.drop_thresh, drop_key, ck);
      } else {
        if (p.drop_thresh) ds16_t<T, true, true>(rs, rd, out, mb, p.scale_log2, lse2, dlt, dsc, p.drop_thresh, drop_key, ck);
        else ds16_t<T, true, false>(rs, rd, out, mb, p.scale_log2, lse2, dlt, dsc, p.drop_thresh, drop_key, ck);
      }
      tmem_st8(t_lane + T_DP + tb * BT + col0, out);   // into the first half of the columns this thread has just read
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[tb]);
    }
    // ---- epilogue: dQ = scale / boost * accumulator
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const float oscale = p.scale / p.ds_boost;
#pragma unroll 1
    for (int c0 = col0; c0 < DK; c0 += 64) {
      uint32_t r[16];
      tmem_ld16(t_lane + T_DQ + c0, r);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack2<T>(__uint_as_float(r[2 * i]) * oscale, __uint_as_float(r[2 * i + 1]) * oscale);
        store_words<8>(static_cast<T*>(p.dq) + (static_cast<int64_t>(b) * p.Lq + row) * p.lddq + h * DK + c0, w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ dK, dV
template <typename T, int DK>
__global__ void __launch_bounds__(NTHREADS, 1)
attn16_bwd_dkv(const T* __restrict__ k, int64_t ldk, const T* __restrict__ v, int64_t ldv,
               const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do, const AttnDev p) {
  constexpr int BKV = 128;
  constexpr int G = DK / 64;
  constexpr int STAGES = 4;
  constexpr int T_BYTES = BT * DK * 2;
  constexpr int STAGE_BYTES = 2 * T_BYTES;   // Q | dO, each used as a K-major (recompute) and as an MN-major (gradient) operand
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_ST = 0, T_DPT = 2 * BT, T_DV = 4 * BT, T_DK = 4 * BT + DK, T_K = 4 * BT + 2 * DK, T_V = 4 * BT + 2 * DK + DK / 2;
  static_assert(4 * BT + 3 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t res_ready, ld_full[STAGES], ld_empty[STAGES], s_full[2], ds_full[2], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[STAGES][BT];
  __shared__ __align__(16) float s_delta[STAGES][BT];
  __shared__ __align__(16) uint32_t s_rkey[STAGES][BT];
  __shared__ int s_extent;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.Lq + BT - 1) / BT;
  // with a causal mask, query tiles entirely before this key tile see none of its keys
  const int t_first = p.causal ? min(kv0 / BT, n_q - 1) : 0;
  pdl_wait();      // before the first global access and before the early exit below (programmatic dependent launch)
  pdl_trigger();
  {
    // a key tile entirely inside this utterance's padding: dK = dV = 0 for its rows, nothing to compute
    const int extent = block_key_extent(p, b, &s_extent);
    if (extent > 0 && kv0 >= extent) {
      const int rows = min(BKV, p.Lk - kv0);
      for (int i = tid; i < rows * (DK / 8); i += NTHREADS) {
        const int r = i / (DK / 8), c = (i - r * (DK / 8)) * 8;
        const int64_t grow = static_cast<int64_t>(b) * p.Lk + kv0 + r;
        *reinterpret_cast<uint4*>(static_cast<T*>(p.dk) + grow * p.lddk + h * DK + c) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(static_cast<T*>(p.dv) + grow * p.lddv + h * DK + c) = make_uint4(0u, 0u, 0u, 0u);
      }
      return;
    }
  }

  if (tid == 0) {
    mbar_init(&res_ready, NCOMP); mbar_init(&acc_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ld_full[s], 1); mbar_init(&ld_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], NCOMP); }
    fence_mbar_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_t = n_q - t_first;   // tiles this CTA walks over: t_first .. n_q - 1

  if (warp == W_PROD) {
    // ===================== producer =====================
    if (lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); }
    for (int t = 0; t < n_t; ++t) {
      const int s = t % STAGES;
      const int qt = (t_first + t) * BT;
      mbar_wait(&ld_empty[s], ((t / STAGES) & 1) ^ 1);
#pragma unroll
      for (int e = lane; e < BT; e += 32) {   // per-query statistics of this tile
        const int q = qt + e;
        const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (q < p.Lq ? q : 0);
        s_lse[s][e] = q < p.Lq ? p.lse2[o] : INFINITY;   // +inf => probability 0 for padded query rows
        s_delta[s][e] = q < p.Lq ? p.delta[o] * p.ds_boost : 0.f;
        s_rkey[s][e] = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(o)) : 0u;
      }
      __syncwarp();
      if (lane == 0) {
        uint8_t* st = sStage + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&ld_full[s], STAGE_BYTES);
        tma_load_4d(st, &tmap_q, &ld_full[s], 0, qt, h * G, b);
        tma_load_4d(st + T_BYTES, &tmap_do, &ld_full[s], 0, qt, h * G, b);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (converged warp + elect.sync) =====================
    {
      const uint32_t st0 = smem_u32(sStage);
      auto koff = [](int ks, int rows) { return static_cast<uint64_t>(((ks / 4) * (rows * 128) + (ks % 4) * 32) >> 4); };
      constexpr int U = STAGES;
      const uint64_t dq0 = umma_desc_kmajor(st0), dqm0 = umma_desc_mn<T>(st0, BT * 128);
      const uint64_t ddo0 = umma_desc_kmajor(st0 + T_BYTES), ddom0 = umma_desc_mn<T>(st0 + T_BYTES, BT * 128);
      auto issue_a = [&](uint32_t ld_parity, auto S, auto TB) {   // S^T(t) = K Q^T, dP^T(t) = V dO^T   (A = resident tile in TMEM)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ld_full[s], ld_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, BT, false, false);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ts(tmem + T_ST + tb * BT, tmem + T_K + ks * 8, dq0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ts(tmem + T_DPT + tb * BT, tmem + T_V + ks * 8, ddo0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          umma_commit(&s_full[tb]);
        }
        __syncwarp();
      };
      auto issue_b = [&](uint32_t ds_parity, bool first, auto S, auto TB) {   // dV += P^T dO, dK += dS^T Q   (A packed in TMEM, B MN-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ds_full[tb], ds_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, DK, false, true);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)
            umma_f16_ts(tmem + T_DV, tmem + T_ST + tb * BT + ks * 16, ddom0 + so + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                        (!first || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)
            umma_f16_ts(tmem + T_DK, tmem + T_DPT + tb * BT + ks * 16, dqm0 + so + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                        (!first || ks > 0) ? 1u : 0u);
          umma_commit(&ld_empty[s]);
        }
        __syncwarp();
      };
      mbar_wait(&res_ready, 0);
      tc_fence_after();
      issue_a(0u, IC<0>{}, IC<0>{});
      for (int t0 = 0; t0 < n_t; t0 += U) {
        const uint32_t ldp = static_cast<uint32_t>(t0 / STAGES), dsp = static_cast<uint32_t>(t0 >> 1);
        auto round = [&](auto Uc) {
          constexpr int u = decltype(Uc)::value;
          const int t = t0 + u;
          if (t < n_t) {
            if (t + 1 < n_t) issue_a((ldp + (u + 1) / STAGES) & 1u, IC<(u + 1) % STAGES>{}, IC<(u + 1) & 1>{});
            issue_b((dsp + (u >> 1)) & 1u, t == 0, IC<u % STAGES>{}, IC<(u & 1)>{});
          }
        };
        cfor<U>(round);
      }
      if (elect_one()) umma_commit(&acc_full);
    }
    __syncwarp();
  } else {
    // ===================== compute warps =====================
    const int quarter = warp & 3, slice = warp >> 2;
    const int key = kv0 + quarter * 32 + lane;
    const bool key_ok = key < p.Lk;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int col0 = slice * 16;
    const int64_t grow = static_cast<int64_t>(b) * p.Lk + (key_ok ? key : 0);
    {
      uint32_t rres[32];
      resident_load16<DK, T>(k, ldk, v, ldv, grow, key_ok, h, slice, rres);
      resident_store16<DK>(slice, t_lane, T_K, T_V, rres);
      tc_fence_before();
      mbar_arrive(&res_ready);
    }
    const bool mask_per_key = (p.mask != nullptr) && (p.ms_q == 0);
    bool key_masked = !key_ok || key >= key_limit(p, b);
    if (mask_per_key && !key_masked) key_masked = p.mask[b * p.ms_b + static_cast<int64_t>(key) * p.ms_k] != 0;
    const bool mask_dense = ((p.mask != nullptr) && !mask_per_key) || p.causal;
    const int causal_key = p.causal ? key : -1;
    const uint8_t* mrow = (p.mask != nullptr && !mask_per_key) ? p.mask + b * p.ms_b + static_cast<int64_t>(key_ok ? key : 0) * p.ms_k : nullptr;
    const uint32_t my_ckey = p.drop_thresh ? dropout_col_key(p.drop_seed, static_cast<uint32_t>(key_ok ? key : 0)) : 0u;
    const float dscale = p.drop_thresh ? p.drop_scale : 1.f;
    const float dsc = dscale * p.ds_boost;
    for (int t = 0; t < n_t; ++t) {
      const int s = t % STAGES, tb = t & 1;
      mbar_wait(&s_full[tb], (t >> 1) & 1);
      tc_fence_after();
      uint32_t rs[16], rd[16], op[8], ods[8];
      // tcgen05.ld/st are warp-collective (.sync.aligned): every lane executes them, whatever its key's mask state
      tmem_ld16(t_lane + T_ST + tb * BT + col0, rs);
      tmem_ld16(t_lane + T_DPT + tb * BT + col0, rd);
      tmem_ld_wait();
      if (!mask_dense && key_masked) {   // this key is padding for every query: P = dS = 0
#pragma unroll
        for (int i = 0; i < 8; ++i) { op[i] = 0u; ods[i] = 0u; }
      } else {
        const float* ls = s_lse[s] + col0;
        const float* de = s_delta[s] + col0;
        const uint32_t* rk = s_rkey[s] + col0;
        const int qf = (t_first + t) * BT + col0;
        if (mask_dense) {
          if (p.drop_thresh) dkv16_t<T, true, true>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
          else dkv16_t<T, true, false>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
        } else {
          if (p.drop_thresh) dkv16_t<T, false, true>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
          else dkv16_t<T, false, false>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
        }
      }
      tmem_st8(t_lane + T_ST + tb * BT + col0, op);     // into the first half of the columns this thread has just read
      tmem_st8(t_lane + T_DPT + tb * BT + col0, ods);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[tb]);
    }
    // ---- epilogue: dV = dropout-scale * acc, dK = softmax-scale / boost * acc
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const float kscale = p.scale / p.ds_boost;
#pragma unroll 1
    for (int c0 = col0; c0 < DK; c0 += 64) {
      uint32_t rv[16], rk[16];
      tmem_ld16(t_lane + T_DV + c0, rv);
      tmem_ld16(t_lane + T_DK + c0, rk);
      tmem_ld_wait();
      if (key_ok) {
        uint32_t wv[8], wk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          wv[i] = pack2<T>(__uint_as_float(rv[2 * i]) * dscale, __uint_as_float(rv[2 * i + 1]) * dscale);
          wk[i] = pack2<T>(__uint_as_float(rk[2 * i]) * kscale, __uint_as_float(rk[2 * i + 1]) * kscale);
        }
        store_words<8>(static_cast<T*>(p.dv) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddv + h * DK + c0, wv);
        store_words<8>(static_cast<T*>(p.dk) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddk + h * DK + c0, wk);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ dK, dV for ONE query tile (Lq <= 64)
// Decoder cross-attention (50 target positions against 1000 encoder frames) and decoder self-attention.  With a single
// query tile the kernel above is one short tile per CTA — barrier / TMEM set-up, the resident K/V load and the output write
// are all exposed (2048 CTAs x ~7 us at B=32, h=8, Lk=1000: 95 us).  Here a CTA keeps the query-side tiles (Q, dO — one
// copy each serves as K-major and as MN-major operand — log-sum-exp, delta, dropout keys) resident and walks over SEVERAL
// key tiles: K/V tiles double-buffered in shared memory (TMA), S^T / dP^T and the dV / dK accumulators double-buffered in
// TMEM, so the loads of tile i+1, the MMAs of tile i and the output write of tile i-1 overlap.  Same mathematics and thread
// mapping as attn16_bwd_dkv (16-bit twin of st_attn_bwd.cu attn_bwd_dkv_small).
template <typename T, int DK>
__global__ void __launch_bounds__(NTHREADS, 1)
attn16_bwd_dkv_small(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_k_res, const __grid_constant__ CUtensorMap tmap_v_res,
                     const AttnDev p) {
  constexpr int BKV = 128;
  constexpr int G = DK / 64;
  constexpr int T_BYTES = BT * DK * 2;
  constexpr int RES_BYTES = BKV * DK * 2;
  constexpr uint32_t TCOLS = 512;
  // TMEM: S^T x2 | dP^T x2 | dV x2 | dK x2
  constexpr uint32_t T_ST = 0, T_DPT = 2 * BT, T_DV = 4 * BT, T_DK = 4 * BT + 2 * DK;
  static_assert(4 * BT + 4 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sBase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRes = sBase;                                  // [2 buffers][K | V]
  uint8_t* sQ = sBase + 4 * RES_BYTES;                    // Q | dO
  __shared__ uint64_t q_full, res_full[2], res_empty[2], s_full[2], ds_full[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[BT];
  __shared__ __align__(16) float s_delta[BT];
  __shared__ __align__(16) uint32_t s_rkey[BT];
  __shared__ int s_extent;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kt_all = (p.Lk + BKV - 1) / BKV;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1); mbar_init(&res_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&ds_full[i], NCOMP);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NCOMP);
    }
    fence_mbar_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  pdl_wait();      // nothing above reads or writes global memory (programmatic dependent launch, st_host.h)
  pdl_trigger();

  const int extent = block_key_extent(p, b, &s_extent);
  // key tiles up to the utterance's last valid key are computed; tiles entirely inside its padding get zeros (below)
  const int n_kt = (extent > 0 && extent < p.Lk) ? (extent + BKV - 1) / BKV : n_kt_all;
  const int first = blockIdx.x, step = gridDim.x;
  const int n_it = first < n_kt ? (n_kt - first + step - 1) / step : 0;   // key tiles of this CTA: first, first + step, ...
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == W_PROD) {
    // ===================== producer =====================
    if (n_it > 0) {
#pragma unroll
      for (int e = lane; e < BT; e += 32) {   // per-query statistics of the single query tile
        const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (e < p.Lq ? e : 0);
        s_lse[e] = e < p.Lq ? p.lse2[o] : INFINITY;   // +inf => probability 0 for padded query rows
        s_delta[e] = e < p.Lq ? p.delta[o] * p.ds_boost : 0.f;
        s_rkey[e] = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(o)) : 0u;
      }
      __syncwarp();
      if (elect_one()) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k_res); tma_prefetch_desc(&tmap_v_res);
        mbar_arrive_expect_tx(&q_full, 2 * T_BYTES);   // release: the statistics above become visible with it
        tma_load_4d(sQ, &tmap_q, &q_full, 0, 0, h * G, b);
        tma_load_4d(sQ + T_BYTES, &tmap_do, &q_full, 0, 0, h * G, b);
      }
      __syncwarp();
      for (int it = 0; it < n_it; ++it) {
        const int rb = it & 1, kv0 = (first + it * step) * BKV;
        if (it >= 2) mbar_wait(&res_empty[rb], ((it >> 1) - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&res_full[rb], 2 * RES_BYTES);
          tma_load_4d(sRes + rb * 2 * RES_BYTES, &tmap_k_res, &res_full[rb], 0, kv0, h * G, b);               // rows >= Lk: zeros
          tma_load_4d(sRes + rb * 2 * RES_BYTES + RES_BYTES, &tmap_v_res, &res_full[rb], 0, kv0, h * G, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (converged warp + elect.sync) =====================
    if (n_it > 0) {
      const uint32_t sq0 = smem_u32(sQ), sr0 = smem_u32(sRes);
      auto koff = [](int ks, int rows) { return static_cast<uint64_t>(((ks / 4) * (rows * 128) + (ks % 4) * 32) >> 4); };
      const uint64_t dqk = umma_desc_kmajor(sq0), dqm = umma_desc_mn<T>(sq0, BT * 128);
      const uint64_t ddok = umma_desc_kmajor(sq0 + T_BYTES), ddom = umma_desc_mn<T>(sq0 + T_BYTES, BT * 128);
      auto issue_a = [&](int it) {   // S^T = K Q^T, dP^T = V dO^T   (A = K / V tile in shared memory, B = Q / dO K-major)
        const int rb = it & 1;
        mbar_wait(&res_full[rb], (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, BT, false, false);
          const uint64_t ak0 = umma_desc_kmajor(sr0 + rb * 2 * RES_BYTES), av0 = umma_desc_kmajor(sr0 + rb * 2 * RES_BYTES + RES_BYTES);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ss(tmem + T_ST + rb * BT, ak0 + koff(ks, BKV), dqk + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < DK / 16; ++ks)
            umma_f16_ss(tmem + T_DPT + rb * BT, av0 + koff(ks, BKV), ddok + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          umma_commit(&s_full[rb]);
          umma_commit(&res_empty[rb]);   // only these two products read the K / V tile: its buffer is free as soon as they retire,
        }                                // so the producer runs two key tiles ahead and the TMA latency stays hidden
        __syncwarp();
      };
      auto issue_b = [&](int it) {   // dV = P^T dO, dK = dS^T Q   (A packed in TMEM, B MN-major)
        const int rb = it & 1;
        mbar_wait(&ds_full[rb], (it >> 1) & 1);
        if (it >= 2) mbar_wait(&acc_empty[rb], ((it >> 1) - 1) & 1);   // the epilogue of tile it-2 has drained this accumulator pair
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc<T>(128, DK, false, true);
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)   // queries [16 ks, +16): the 8 packed columns slice ks wrote at 16 ks
            umma_f16_ts(tmem + T_DV + rb * DK, tmem + T_ST + rb * BT + ks * 16, ddom + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                        ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)
            umma_f16_ts(tmem + T_DK + rb * DK, tmem + T_DPT + rb * BT + ks * 16, dqm + static_cast<uint64_t>((ks * 2048) >> 4), idesc,
                        ks > 0 ? 1u : 0u);
          umma_commit(&acc_full[rb]);
        }
        __syncwarp();
      };
      mbar_wait(&q_full, 0);
      tc_fence_after();
      issue_a(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_a(it + 1);
        issue_b(it);
      }
    }
  } else {
    // ===================== compute warps =====================
    const int quarter = warp & 3, slice = warp >> 2;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int col0 = slice * 16;
    const bool mask_per_key = (p.mask != nullptr) && (p.ms_q == 0);
    const bool mask_dense = ((p.mask != nullptr) && !mask_per_key) || p.causal;
    const float dscale = p.drop_thresh ? p.drop_scale : 1.f;
    const float dsc = dscale * p.ds_boost;
    const float kscale = p.scale / p.ds_boost;
    auto epilogue = [&](int it) {   // dV = dropout-scale * acc, dK = softmax-scale / boost * acc for the key tile of iteration `it`
      const int rb = it & 1;
      const int key = (first + it * step) * BKV + quarter * 32 + lane;
      const bool key_ok = key < p.Lk;
      mbar_wait(&acc_full[rb], (it >> 1) & 1);
      tc_fence_after();
      if (col0 < DK) {
        uint32_t rv[16], rk[16];
        tmem_ld16(t_lane + T_DV + rb * DK + col0, rv);
        tmem_ld16(t_lane + T_DK + rb * DK + col0, rk);
        tmem_ld_wait();
        if (key_ok) {
          uint32_t wv[8], wk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            wv[i] = pack2<T>(__uint_as_float(rv[2 * i]) * dscale, __uint_as_float(rv[2 * i + 1]) * dscale);
            wk[i] = pack2<T>(__uint_as_float(rk[2 * i]) * kscale, __uint_as_float(rk[2 * i + 1]) * kscale);
          }
          store_words<8>(static_cast<T*>(p.dv) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddv + h * DK + col0, wv);
          store_words<8>(static_cast<T*>(p.dk) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddk + h * DK + col0, wk);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[rb]);
    };
    if (n_it > 0) mbar_wait(&q_full, 0);   // the statistics in shared memory are valid from here on
    for (int it = 0; it < n_it; ++it) {
      const int rb = it & 1;
      const int key = (first + it * step) * BKV + quarter * 32 + lane;
      const bool key_ok = key < p.Lk;
      bool key_masked = !key_ok || key >= key_limit(p, b);
      if (mask_per_key && !key_masked) key_masked = p.mask[b * p.ms_b + static_cast<int64_t>(key) * p.ms_k] != 0;
      const int causal_key = p.causal ? key : -1;
      const uint8_t* mrow = (p.mask != nullptr && !mask_per_key) ? p.mask + b * p.ms_b + static_cast<int64_t>(key_ok ? key : 0) * p.ms_k : nullptr;
      const uint32_t my_ckey = p.drop_thresh ? dropout_col_key(p.drop_seed, static_cast<uint32_t>(key_ok ? key : 0)) : 0u;
      mbar_wait(&s_full[rb], (it >> 1) & 1);
      tc_fence_after();
      uint32_t rs[16], rd[16], op[8], ods[8];
      tmem_ld16(t_lane + T_ST + rb * BT + col0, rs);
      tmem_ld16(t_lane + T_DPT + rb * BT + col0, rd);
      tmem_ld_wait();
      if (!mask_dense && key_masked) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { op[i] = 0u; ods[i] = 0u; }
      } else {
        const float* ls = s_lse + col0;
        const float* de = s_delta + col0;
        const uint32_t* rk = s_rkey + col0;
        if (mask_dense) {
          if (p.drop_thresh) dkv16_t<T, true, true>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
          else dkv16_t<T, true, false>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
        } else {
          if (p.drop_thresh) dkv16_t<T, false, true>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
          else dkv16_t<T, false, false>(rs, rd, op, ods, ls, de, rk, p.scale_log2, dsc, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
        }
      }
      tmem_st8(t_lane + T_ST + rb * BT + col0, op);     // into the first half of the columns this thread has just read
      tmem_st8(t_lane + T_DPT + rb * BT + col0, ods);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[rb]);
      if (it > 0) epilogue(it - 1);       // overlaps the gradient MMAs of this tile
    }
    if (n_it > 0) epilogue(n_it - 1);
    // key tiles entirely inside the utterance's padding: zero gradients
    for (int kt = n_kt + first; kt < n_kt_all; kt += step) {
      const int rows = min(BKV, p.Lk - kt * BKV);
      for (int i = tid; i < rows * (DK / 8); i += NCOMP) {
        const int r = i / (DK / 8), c = (i - r * (DK / 8)) * 8;
        const int64_t grow = static_cast<int64_t>(b) * p.Lk + kt * BKV + r;
        *reinterpret_cast<uint4*>(static_cast<T*>(p.dk) + grow * p.lddk + h * DK + c) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(static_cast<T*>(p.dv) + grow * p.lddv + h * DK + c) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ host side
int check_attn16(const AttnArgs& a, const char* who) {
  ST_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "%s: empty problem", who);
  ST_REQUIRE(a.dk == 64, "%s: 16-bit operands need d_k = 64 (got %d)", who, a.dk);
  ST_REQUIRE((a.ldq & 7) == 0 && (a.ldk & 7) == 0 && (a.ldv & 7) == 0 && (a.ldctx & 7) == 0,
             "%s: leading dimensions must be multiples of 8 elements", who);
  ST_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
               reinterpret_cast<uintptr_t>(a.ctx)) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  return ST_OK;
}

template <typename T, int DK>
int launch_fwd16(cudaStream_t s, const AttnArgs& a) {
  constexpr int DT = dtype_of<T>();
  CUtensorMap tq, tk, tv;
  const int cols = a.H * DK;
  ST_TRY(make_act_tmap16(&tq, DT, a.q, a.ldq, cols, a.Lq, a.B, 128, DK));
  ST_TRY(make_act_tmap16(&tk, DT, a.k, a.ldk, cols, a.Lk, a.B, 128, DK));
  ST_TRY(make_act_tmap16(&tv, DT, a.v, a.ldv, cols, a.Lk, a.B, 128, DK));
  constexpr int SMEM = 3 * 128 * DK * 2 + 1024;
  auto kern = attn16_fwd_kernel<T, DK>;
  static bool attr = false;
  if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
  dim3 grid((a.Lq + 127) / 128, a.H, a.B);
  ProfScope prof(s, PROF_ATTN_FWD, 4.0 * a.B * a.H * static_cast<double>(a.Lq) * a.Lk * DK);
  AttnDev dev = attn_to_dev(a);
  ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(256), SMEM, s, tq, tk, tv, dev));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

template <typename T, int DK>
int launch_bwd16(cudaStream_t s, const AttnBwdArgs& a) {
  constexpr int DT = dtype_of<T>();
  const AttnArgs& f = a.f;
  const int cols = f.H * DK;
  AttnDev p = attn_to_dev(f);
  p.delta = a.delta; p.dq = a.dq; p.lddq = a.lddq; p.dk = a.dk_; p.lddk = a.lddk; p.dv = a.dv; p.lddv = a.lddv;
  p.ds_boost = DT == ST_DTYPE_F16 ? 256.f : 1.f;
  {
    const int64_t rows = static_cast<int64_t>(f.B) * f.Lq;
    const int64_t blocks = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
    ProfScope prof(s, PROF_ATTN_DELTA, 2.0 * rows * cols * 2);
    ST_CHECK_CUDA(launch_pdl(attn16_delta_kernel<T>, dim3(static_cast<unsigned>(blocks < cap ? blocks : cap)), dim3(256), 0, s,
                             static_cast<const T*>(a.dctx), a.lddctx, static_cast<const T*>(f.ctx), f.ldctx, a.delta, f.B, f.H,
                             f.Lq, DK));
    ST_CHECK_LAUNCH();
  }
  constexpr int SMEM = 4 * 2 * BT * DK * 2 + 1024;
  Fork fk(s);                           // dK/dV and dQ only share their inputs: two streams (st_host.h)
  cudaStream_t s_dq = fk.branch(1);
  if (f.Lq <= BT && !get_option("attn_dkv_no_small")) {   // single query tile: the persistent multi-key-tile kernel
    CUtensorMap tq, tdo, tkr, tvr;
    ST_TRY(make_act_tmap16(&tq, DT, f.q, f.ldq, cols, f.Lq, f.B, BT, DK));
    ST_TRY(make_act_tmap16(&tdo, DT, a.dctx, a.lddctx, cols, f.Lq, f.B, BT, DK));
    ST_TRY(make_act_tmap16(&tkr, DT, f.k, f.ldk, cols, f.Lk, f.B, 128, DK));
    ST_TRY(make_act_tmap16(&tvr, DT, f.v, f.ldv, cols, f.Lk, f.B, 128, DK));
    constexpr int SMEM_SMALL = 4 * 128 * DK * 2 + 2 * BT * DK * 2 + 1024;
    auto ks = attn16_bwd_dkv_small<T, DK>;
    static bool attr_small = false;
    if (!attr_small) { ST_CHECK_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL)); attr_small = true; }
    const int n_kt = (f.Lk + 127) / 128;
    int split = get_option("attn_dkv_small_split");
    if (split <= 0) split = 1;
    dim3 grid(n_kt < split ? n_kt : split, f.H, f.B);
    ProfScope prof(s, PROF_ATTN_DKV, 4.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    ST_CHECK_CUDA(launch_pdl(ks, grid, dim3(NTHREADS), SMEM_SMALL, s, tq, tdo, tkr, tvr, p));
    ST_CHECK_LAUNCH();
  } else {
    CUtensorMap tq, tdo;
    ST_TRY(make_act_tmap16(&tq, DT, f.q, f.ldq, cols, f.Lq, f.B, BT, DK));
    ST_TRY(make_act_tmap16(&tdo, DT, a.dctx, a.lddctx, cols, f.Lq, f.B, BT, DK));
    auto kern = attn16_bwd_dkv<T, DK>;
    static bool attr = false;
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    dim3 grid((f.Lk + 127) / 128, f.H, f.B);
    ProfScope prof(s, PROF_ATTN_DKV, 4.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), SMEM, s, static_cast<const T*>(f.k), f.ldk, static_cast<const T*>(f.v), f.ldv,
                             tq, tdo, p));
    ST_CHECK_LAUNCH();
  }
  {
    CUtensorMap tk, tv;
    ST_TRY(make_act_tmap16(&tk, DT, f.k, f.ldk, cols, f.Lk, f.B, BT, DK));
    ST_TRY(make_act_tmap16(&tv, DT, f.v, f.ldv, cols, f.Lk, f.B, BT, DK));
    auto kern = attn16_bwd_dq<T, DK>;
    static bool attr = false;
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    dim3 grid((f.Lq + 127) / 128, f.H, f.B);
    ProfScope prof(s_dq, PROF_ATTN_DQ, 6.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), SMEM, s_dq, static_cast<const T*>(f.q), f.ldq, static_cast<const T*>(a.dctx),
                             a.lddctx, tk, tv, p));
    ST_CHECK_LAUNCH();
  }
  return fk.join();
}

}  // namespace

int attn16_fwd(cudaStream_t s, const AttnArgs& a) {
  ST_TRY(check_attn16(a, "attn_fwd"));
  ST_REQUIRE(a.lse != nullptr, "attn_fwd: lse buffer is required");
  if (a.dtype == ST_DTYPE_F16) return launch_fwd16<__half, 64>(s, a);
  if (a.dtype == ST_DTYPE_BF16) return launch_fwd16<__nv_bfloat16, 64>(s, a);
  set_error("attn_fwd: bad dtype %d", a.dtype);
  return ST_ERR_INVALID;
}

int attn16_bwd(cudaStream_t s, const AttnBwdArgs& a) {
  ST_TRY(check_attn16(a.f, "attn_bwd"));
  ST_REQUIRE(a.dctx && a.delta && a.dq && a.dk_ && a.dv && a.f.lse, "attn_bwd: null buffer");
  ST_REQUIRE((a.lddctx & 7) == 0 && (a.lddq & 7) == 0 && (a.lddk & 7) == 0 && (a.lddv & 7) == 0,
             "attn_bwd: leading dimensions must be multiples of 8 elements");
  if (a.f.dtype == ST_DTYPE_F16) return launch_bwd16<__half, 64>(s, a);
  if (a.f.dtype == ST_DTYPE_BF16) return launch_bwd16<__nv_bfloat16, 64>(s, a);
  set_error("attn_bwd: bad dtype %d", a.f.dtype);
  return ST_ERR_INVALID;
}

}  // namespace st
