// st_attn.cu — fused multi-head attention core on tcgen05 (TF32 operands, fp32 accumulate).
//
// Reference: transformer/Attention.py:78-90 — split heads, scores = Q K^T / sqrt(d_k),
// masked_fill_(mask, -inf), softmax, dropout, P V, merge heads — and its autograd backward.
// The (B, h, Lq, Lk) score tensor the reference materialises (1 GB per layer at B=32, T=1000) never
// leaves the SM here: S lives in TMEM, the softmax runs out of TMEM, P goes back to TMEM as the
// A operand of the P·V MMA.
//
// Forward  : one CTA per (128-query tile, head, batch), flash-style online softmax over key tiles;
//            256 threads = 2 threads per query row (each owns half of the tile's columns), 2 CTAs per SM
//            so one CTA's MMAs overlap the other's softmax.
// Backward : two kernels, no atomics —
//   dKV : one CTA per (128-key tile, head, batch), loops over query tiles; S^T = K Q^T and
//         dP^T = V dO^T are recomputed, dV += P^T dO and dK += dS^T Q accumulate in TMEM;
//   dQ  : one CTA per (128-query tile, head, batch), loops over key tiles; dQ += dS K in TMEM.
//   Both use 4 threads per tile row (one 32-column slice each, 512 threads).
// These kernels are bound by the softmax ALU work (exp2, dropout hash, TF32 rounding: ~16-20
// instructions per score against 2 MMAs of 512 cycles per 128x128 tile), so the layout maximises
// resident warps rather than tile size.
// A tile that is consumed both along d (K-major operand, 16-byte swizzle) and along the sequence
// (MN-major operand, which for TF32 must use the 32-byte-atom swizzle) is fetched by two TMA loads.
//
// TMEM lane = tile row; warp w may touch lanes 32*(w%4)..+31, so warps w, w+4, w+8, ... share rows and
// split columns.  Masks are arbitrary byte tensors with strides (stride 0 over queries for the
// key-padding masks of Utils.py:41-57); masked => probability exactly 0; a fully masked row gives
// NaN, as the reference does.
#include <math.h>

#include "st_attn.cuh"

namespace st {

namespace {

// in-kernel timeline of CTA (0,0,0), thread 0 of the forward kernel (option "attn_trace"): [tile][event] clock64 stamps
constexpr int FTRACE_TILES = 16, FTRACE_EVENTS = 8;
__device__ unsigned long long g_ftrace[FTRACE_TILES * FTRACE_EVENTS];
#define ST_FTRACE(tile, ev)                                                                                   \
  do {                                                                                                        \
    if (ftrace_on && (tile) < FTRACE_TILES) g_ftrace[(tile) * FTRACE_EVENTS + (ev)] = clock64();              \
  } while (0)

__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], uint32_t mb) {
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

// r <- tf32(keep ? exp2(r*scale_log2 - m) : 0) with masked entries 0; returns the (pre-dropout) row-sum part.
// The inverted-dropout scale is NOT applied here (it is folded into the final normalisation of O).
// ckey: this chunk's 32 per-column dropout keys in shared memory (16-byte aligned); rowkey: the row's key.
template <bool MASK, bool DROP>
__device__ __forceinline__ float chunk_probs_t(uint32_t (&r)[32], uint32_t mb, float scale_log2, float m_use,
                                               uint32_t thresh, uint32_t rowkey, const uint32_t* ckey) {
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    uint4 ck = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) ck = *reinterpret_cast<const uint4*>(ckey + i);
    const uint32_t cks[4] = {ck.x, ck.y, ck.z, ck.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float pv = fast_exp2(fmaf(__uint_as_float(r[i + t]), scale_log2, -m_use));
      if (MASK && ((mb >> (i + t)) & 1u)) pv = 0.f;
      l += pv;
      if (DROP && !dropout_keep_xor(rowkey, cks[t], thresh)) pv = 0.f;
      r[i + t] = tf32_rna_mma_bits(pv);
    }
  }
  return l;
}
// dispatch on (any masked entry in this chunk?, dropout on?) so the common unmasked path carries no per-element tests
__device__ __forceinline__ float chunk_probs(uint32_t (&r)[32], uint32_t mb, float scale_log2, float m_use,
                                             uint32_t thresh, uint32_t rowkey, const uint32_t* ckey) {
  if (mb == 0u) {
    return thresh ? chunk_probs_t<false, true>(r, mb, scale_log2, m_use, thresh, rowkey, ckey)
                  : chunk_probs_t<false, false>(r, mb, scale_log2, m_use, thresh, rowkey, ckey);
  }
  return thresh ? chunk_probs_t<true, true>(r, mb, scale_log2, m_use, thresh, rowkey, ckey)
                : chunk_probs_t<true, false>(r, mb, scale_log2, m_use, thresh, rowkey, ckey);
}

// dS chunk of the dQ kernel: rd <- tf32(P * (keep ? dP*dscale : 0) - P*delta), P = exp2(S*scale_log2 - lse2)
template <bool MASK, bool DROP>
__device__ __forceinline__ void chunk_ds_t(const uint32_t (&rs)[32], uint32_t (&rd)[32], uint32_t mb, float scale_log2,
                                           float lse2, float delta, float dscale, uint32_t thresh, uint32_t rowkey,
                                           const uint32_t* ckey) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    uint4 ck = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) ck = *reinterpret_cast<const uint4*>(ckey + i);
    const uint32_t cks[4] = {ck.x, ck.y, ck.z, ck.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float pr = fast_exp2(fmaf(__uint_as_float(rs[i + t]), scale_log2, -lse2));
      if (MASK && ((mb >> (i + t)) & 1u)) pr = 0.f;
      float dp = __uint_as_float(rd[i + t]);
      if (DROP && !dropout_keep_xor(rowkey, cks[t], thresh)) dp = 0.f;
      rd[i + t] = tf32_rna_mma_bits(pr * fmaf(dp, dscale, -delta));
    }
  }
}
__device__ __forceinline__ void chunk_ds(const uint32_t (&rs)[32], uint32_t (&rd)[32], uint32_t mb, float scale_log2,
                                         float lse2, float delta, float dscale, uint32_t thresh, uint32_t rowkey,
                                         const uint32_t* ckey) {
  if (mb == 0u) {
    if (thresh) chunk_ds_t<false, true>(rs, rd, mb, scale_log2, lse2, delta, dscale, thresh, rowkey, ckey);
    else chunk_ds_t<false, false>(rs, rd, mb, scale_log2, lse2, delta, dscale, thresh, rowkey, ckey);
  } else {
    if (thresh) chunk_ds_t<true, true>(rs, rd, mb, scale_log2, lse2, delta, dscale, thresh, rowkey, ckey);
    else chunk_ds_t<true, false>(rs, rd, mb, scale_log2, lse2, delta, dscale, thresh, rowkey, ckey);
  }
}

// P^T / dS^T chunk of the dK/dV kernel (thread = key row, 32 query columns):
//   rs <- tf32(keep ? P : 0),  rd <- tf32(P * ((keep ? dP : 0) * dscale - delta_q)),  P = exp2(S*scale_log2 - lse2_q)
// lse / delta / rkey: this chunk's per-query statistics and dropout row keys in shared memory (16-byte aligned).
template <bool DENSE, bool DROP>
__device__ __forceinline__ void chunk_dkv_t(uint32_t (&rs)[32], uint32_t (&rd)[32], const float* lse, const float* delta,
                                            const uint32_t* rkey, float scale_log2, float dscale, uint32_t thresh,
                                            uint32_t my_ckey, const uint8_t* mrow, int64_t ms_q, int q_first, int Lq,
                                            bool key_ok, int causal_key = -1) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 lse4 = *reinterpret_cast<const float4*>(lse + i);
    const float4 del4 = *reinterpret_cast<const float4*>(delta + i);
    uint4 key4 = make_uint4(0u, 0u, 0u, 0u);
    if (DROP) key4 = *reinterpret_cast<const uint4*>(rkey + i);
    const float lses[4] = {lse4.x, lse4.y, lse4.z, lse4.w};
    const float dels[4] = {del4.x, del4.y, del4.z, del4.w};
    const uint32_t rks[4] = {key4.x, key4.y, key4.z, key4.w};
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
      rs[i + t] = tf32_rna_mma_bits(pd);
      rd[i + t] = tf32_rna_mma_bits(pr * fmaf(dp, dscale, -dels[t]));
    }
  }
}

// ================================================================================ forward
template <int DK, int BKV>
__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnDev p) {
  constexpr int BQ = 128;
  constexpr int G = DK / 32;                // 32-column groups (one TMA box each)
  constexpr int Q_BYTES = BQ * DK * 4;
  constexpr int KV_BYTES = BKV * DK * 4;
  constexpr uint32_t TCOLS = 256;           // S/P at [0,BKV), O at [BKV, BKV+DK)
  constexpr int SH = BKV / 2;               // score columns per thread
  constexpr int NCH = SH / 32;              // 32-column chunks per thread
  constexpr int OH = DK / 2;                // output columns per thread
  static_assert(BKV + DK <= 256 && NCH >= 1, "tile configuration");

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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int rit = quarter * 32 + lane;      // row in tile
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + rit;
  const bool row_ok = row < p.Lq;
  const int n_kv_all = (p.Lk + BKV - 1) / BKV;
  __shared__ int s_extent;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(&bar_q, 1); mbar_init(&bar_k, 1); mbar_init(&bar_v, 1); mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  pdl_wait();      // nothing above reads or writes global memory (programmatic dependent launch, st_host.h)
  pdl_trigger();

  const int extent = block_key_extent(p, b, &s_extent);
  // key tiles that lie entirely in the padding of this utterance contribute exactly nothing: skip them (at least one
  // tile is always processed so that a fully masked row still produces the reference's NaN)
  const int n_kv = extent >= p.Lk ? n_kv_all : max(1, (extent + BKV - 1) / BKV);
  // key-padding masks (stride 0 over queries, Utils.py:53-54) and "no mask" give every row of the tile the same
  // bits: compute them once per tile (one thread per 32-key chunk), one tile ahead
  const bool shared_mask = mask_is_row_invariant(p);
  if (shared_mask && tid < BKV / 32) s_mb[0][tid] = mask_bits_row(p, b, 0, true, tid * 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
  constexpr uint32_t T_S = 0, T_O = BKV;

  uint32_t k_loads = 0, s_count = 0;  // phase counters for bar_k / bar_s (thread 0 issues, all wait)

  auto load_k = [&](int j) {
    mbar_arrive_expect_tx(&bar_k, KV_BYTES);
    tma_load_4d(sK, &tmap_k, &bar_k, 0, j * BKV, h * G, b);
  };
  auto load_v = [&](int j) {
    mbar_arrive_expect_tx(&bar_v, KV_BYTES);
    tma_load_4d(sV, &tmap_v, &bar_v, 0, j * BKV, h * G, b);
  };
  auto issue_s = [&]() {  // S = Q K^T  (both K-major)
    constexpr uint32_t idesc = umma_idesc_tf32(128, BKV, false, false);
    const uint64_t aq0 = umma_desc_kmajor(smem_u32(sQ)), bk0 = umma_desc_kmajor(smem_u32(sK));
#pragma unroll
    for (int ks = 0; ks < DK / 8; ++ks)
      umma_tf32_ss(tmem + T_S, aq0 + static_cast<uint64_t>(((ks / 4) * (BQ * 128) + (ks % 4) * 32) >> 4),
                   bk0 + static_cast<uint64_t>(((ks / 4) * (BKV * 128) + (ks % 4) * 32) >> 4), idesc, ks > 0 ? 1u : 0u);
    umma_commit(&bar_s);
  };

  // tcgen05.mma / TMA are issued by ONE lane of a CONVERGED warp 0, chosen with elect.sync: under a plain `tid == 0`
  // branch the compiler wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop (60-120 cycles each, tools/mma_bench.py).
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

  // Online softmax, ONE read of S per tile.  The output accumulates in TMEM (P·V with the accumulate flag) in units of
  // exp2(-m_ref): the reference maximum m_ref is only advanced — and O / l rescaled — when a tile's maximum exceeds it by
  // more than TAU (probabilities stay <= 2^TAU, harmless in fp32 / TF32), which after the first tiles is rare.  This
  // removes the second tcgen05.ld sweep over S, the per-tile read-back of O and its 32-FMA register update, and lets
  // S(j+1) be queued right behind P·V(j) on the tensor pipe.
  constexpr float TAU = 8.f;
  float m_ref = -INFINITY, l_run = 0.f;
  const bool ftrace_on = p.trace && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const uint64_t rng_row = (static_cast<uint64_t>(b) * p.H + h) * p.Lq + (row_ok ? row : 0);
  const uint32_t drop_key = p.drop_thresh ? dropout_row_key(p.drop_seed, rng_row) : 0u;

  for (int j = 0; j < n_kv; ++j) {
    ST_FTRACE(j, 0);
    mbar_wait(&bar_s, s_count & 1);   // S_j complete; the tensor pipe is in order, so P·V(j-1) has completed too
    ++s_count;
    tc_fence_after();
    ST_FTRACE(j, 1);
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

    // ---- this thread's half of the score row, read once
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
    ST_FTRACE(j, 2);
#pragma unroll
    for (int c = 0; c < NCH; ++c) mx = fmaxf(mx, chunk_max(r[c], mbits[c]));
    s_part[half][rit] = mx;
    __syncthreads();
    ST_FTRACE(j, 3);
    mx = fmaxf(s_part[0][rit], s_part[1][rit]) * p.scale_log2;
    // ---- advance the reference maximum only when needed; rescale l and (warp-collectively) this thread's half of O
    const bool need = mx > m_ref + TAU || (m_ref == -INFINITY && mx > -INFINITY);
    if (__any_sync(0xffffffffu, need)) {
      const float alpha = need ? fast_exp2(m_ref - mx) : 1.f;   // m_ref == -inf -> 0
      if (j > 0) {
        if constexpr (OH >= 32) {
#pragma unroll
          for (int c = 0; c < OH / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(t_lane + T_O + half * OH + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(need ? __uint_as_float(o[i]) * alpha : __uint_as_float(o[i]));
            tmem_st32(t_lane + T_O + half * OH + c * 32, o);
          }
        } else {
          uint32_t o[16];
          tmem_ld16(t_lane + T_O + half * OH, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(need ? __uint_as_float(o[i]) * alpha : __uint_as_float(o[i]));
          tmem_st16(t_lane + T_O + half * OH, o);
        }
      }
      if (need) { l_run *= alpha; m_ref = mx; }
    }
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    // ---- probabilities -> TMEM (A operand of P·V), row sum
    float l_tile = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col0 = half * SH + c * 32;
      l_tile += chunk_probs(r[c], mbits[c], p.scale_log2, m_use, p.drop_thresh, drop_key, s_ckey + col0);
      tmem_st32(t_lane + T_S + col0, r[c]);
    }
    l_run += l_tile;
    tmem_st_wait();
    ST_FTRACE(j, 4);
    tc_fence_before();
    __syncthreads();
    ST_FTRACE(j, 5);

    if (warp == 0) {  // O += P V   (A = P in TMEM, B = V MN-major), then S_{j+1} right behind it
      tc_fence_after();
      mbar_wait(&bar_v, j & 1);
      tc_fence_after();
      ST_FTRACE(j, 6);
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
        const uint64_t bv0 = umma_desc_mnmajor(smem_u32(sV), BKV * 128);
#pragma unroll
        for (int ks = 0; ks < BKV / 8; ++ks)
          umma_tf32_ts(tmem + T_O, tmem + T_S + ks * 8, bv0 + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
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
    ST_FTRACE(j, 7);
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
    float* dst = static_cast<float*>(p.ctx) + (static_cast<int64_t>(b) * p.Lq + (row_ok ? row : 0)) * p.ldctx + h * DK + half * OH;
    if constexpr (OH >= 32) {
#pragma unroll
      for (int c = 0; c < OH / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(t_lane + T_O + half * OH + c * 32, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + i) =
                make_float4(tf32_rna(__uint_as_float(o[i]) * inv_l), tf32_rna(__uint_as_float(o[i + 1]) * inv_l),
                            tf32_rna(__uint_as_float(o[i + 2]) * inv_l), tf32_rna(__uint_as_float(o[i + 3]) * inv_l));
        }
      }
    } else {
      uint32_t o[16];
      tmem_ld16(t_lane + T_O + half * OH, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + i) =
              make_float4(tf32_rna(__uint_as_float(o[i]) * inv_l), tf32_rna(__uint_as_float(o[i + 1]) * inv_l),
                          tf32_rna(__uint_as_float(o[i + 2]) * inv_l), tf32_rna(__uint_as_float(o[i + 3]) * inv_l));
      }
    }
    if (row_ok && half == 0) p.lse2[(static_cast<int64_t>(b) * p.H + h) * p.Lq + row] = lse2;
  }

  // ---- optional second sweep: materialise the (post-dropout) probabilities the module returns
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
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ dctx, int64_t lddctx, const float* __restrict__ ctx, int64_t ldctx,
                  float* __restrict__ delta, int B, int H, int Lq, int dk) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t rows = static_cast<int64_t>(B) * Lq;
  const int d = H * dk;
  const int grp = dk / 4;  // lanes per head within a 128-column chunk (8, 16 or 32)
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const int b = static_cast<int>(row / Lq), q = static_cast<int>(row - static_cast<int64_t>(b) * Lq);
    for (int c0 = 0; c0 < d; c0 += 128) {
      const int c = c0 + lane * 4;
      float s = 0.f;
      if (c < d) {
        const float4 a = *reinterpret_cast<const float4*>(dctx + row * lddctx + c);
        const float4 o = *reinterpret_cast<const float4*>(ctx + row * ldctx + c);
        s = (a.x * o.x + a.y * o.y) + (a.z * o.z + a.w * o.w);
      }
      for (int off = 1; off < grp; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (c < d && (lane % grp) == 0) delta[(static_cast<int64_t>(b) * H + c / dk) * Lq + q] = s;
    }
  }
}

// ================================================================================ backward: dK, dV
template <int DK, int BQ>
__global__ void __launch_bounds__(4 * BQ, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_q_k, const __grid_constant__ CUtensorMap tmap_q_mn,
                    const __grid_constant__ CUtensorMap tmap_do_k, const __grid_constant__ CUtensorMap tmap_do_mn,
                    const AttnDev p) {
  constexpr int BKV = 128;
  constexpr int NS = BQ / 32;               // column slices = threads per key row
  constexpr int G = DK / 32;
  constexpr int KV_BYTES = BKV * DK * 4;
  constexpr int QT_BYTES = BQ * DK * 4;
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_ST = 0, T_DPT = BQ, T_DV = 2 * BQ, T_DK = 2 * BQ + DK;
  static_assert(2 * BQ + 2 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQk = sV + KV_BYTES;
  uint8_t* sQm = sQk + QT_BYTES;
  uint8_t* sDOk = sQm + QT_BYTES;
  uint8_t* sDOm = sDOk + QT_BYTES;
  __shared__ uint64_t bar_kv, bar_ld, bar_s, bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[BQ];
  __shared__ __align__(16) float s_delta[BQ];
  __shared__ __align__(16) uint32_t s_key[BQ];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, slice = warp >> 2;
  const int kv0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int key = kv0 + quarter * 32 + lane;
  const bool key_ok = key < p.Lk;
  const int n_q = (p.Lq + BQ - 1) / BQ;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_q_k);
    tma_prefetch_desc(&tmap_q_mn); tma_prefetch_desc(&tmap_do_k); tma_prefetch_desc(&tmap_do_mn);
    mbar_init(&bar_kv, 1); mbar_init(&bar_ld, 1); mbar_init(&bar_s, 1); mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_kv, 2 * KV_BYTES);
      tma_load_4d(sK, &tmap_k, &bar_kv, 0, kv0, h * G, b);
      tma_load_4d(sV, &tmap_v, &bar_kv, 0, kv0, h * G, b);

  }
  // key-padding masks (stride 0 over queries) are a per-thread constant
  const bool mask_per_key = (p.mask != nullptr) && (p.ms_q == 0);
  bool key_masked = !key_ok || key >= key_limit(p, b);
  if (mask_per_key && !key_masked) key_masked = p.mask[b * p.ms_b + static_cast<int64_t>(key) * p.ms_k] != 0;
  const bool mask_dense = ((p.mask != nullptr) && !mask_per_key) || p.causal;
  const int causal_key = p.causal ? key : -1;
  const uint32_t my_ckey = p.drop_thresh ? dropout_col_key(p.drop_seed, static_cast<uint32_t>(key_ok ? key : 0)) : 0u;
  const float dscale = p.drop_thresh ? p.drop_scale : 1.f;

  for (int it = 0; it < n_q; ++it) {
    const int q0 = it * BQ;
    if (tid == 0) {
      if (it > 0) { mbar_wait(&bar_acc, (it - 1) & 1); tc_fence_after(); }  // previous MMAs done with Q/dO tiles
      mbar_arrive_expect_tx(&bar_ld, 4 * QT_BYTES);
        tma_load_4d(sQk, &tmap_q_k, &bar_ld, 0, q0, h * G, b);
        tma_load_4d(sQm, &tmap_q_mn, &bar_ld, 0, q0, h * G, b);
        tma_load_4d(sDOk, &tmap_do_k, &bar_ld, 0, q0, h * G, b);
        tma_load_4d(sDOm, &tmap_do_mn, &bar_ld, 0, q0, h * G, b);

    }
    if (tid >= 3 * BQ) {  // the last BQ threads (never the issuing thread 0)
      const int t = tid - 3 * BQ;
      const int q = q0 + t;
      const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (q < p.Lq ? q : 0);
      s_lse[t] = q < p.Lq ? p.lse2[o] : INFINITY;  // +inf => probability 0 for padded query rows
      s_delta[t] = q < p.Lq ? p.delta[o] : 0.f;
      s_key[t] = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(o)) : 0u;
    }
    if (tid == 0) {
      if (it == 0) mbar_wait(&bar_kv, 0);
      mbar_wait(&bar_ld, it & 1);
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_tf32(128, BQ, false, false);
      const uint32_t ak = smem_u32(sK), av = smem_u32(sV), bq = smem_u32(sQk), bdo = smem_u32(sDOk);
#pragma unroll
      for (int ks = 0; ks < DK / 8; ++ks)  // S^T = K Q^T
        umma_tf32_ss(tmem + T_ST, umma_desc_kmajor(ak + (ks / 4) * (BKV * 128) + (ks % 4) * 32),
                     umma_desc_kmajor(bq + (ks / 4) * (BQ * 128) + (ks % 4) * 32), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < DK / 8; ++ks)  // dP^T = V dO^T
        umma_tf32_ss(tmem + T_DPT, umma_desc_kmajor(av + (ks / 4) * (BKV * 128) + (ks % 4) * 32),
                     umma_desc_kmajor(bdo + (ks / 4) * (BQ * 128) + (ks % 4) * 32), idesc, ks > 0 ? 1u : 0u);
      umma_commit(&bar_s);
    }
    __syncthreads();  // s_lse / s_delta / s_key visible
    mbar_wait(&bar_s, it & 1);
    tc_fence_after();

    {
      const int c = slice;  // this thread's 32 query columns
      uint32_t rs[32], rd[32];
      tmem_ld32(t_lane + T_ST + c * 32, rs);
      tmem_ld32(t_lane + T_DPT + c * 32, rd);
      tmem_ld_wait();
      // The inverted-dropout scale of P (-> dV) and the softmax scale of dS (-> dK) are applied once in the epilogue.
      if (!mask_dense && key_masked) {  // this key is padding for every query: P = dS = 0
#pragma unroll
        for (int i = 0; i < 32; ++i) { rs[i] = 0u; rd[i] = 0u; }
      } else {
        const uint8_t* mrow = (p.mask != nullptr && !mask_per_key) ? p.mask + b * p.ms_b + static_cast<int64_t>(key_ok ? key : 0) * p.ms_k : nullptr;
        if (mask_dense) {
          if (p.drop_thresh) chunk_dkv_t<true, true>(rs, rd, s_lse + c * 32, s_delta + c * 32, s_key + c * 32, p.scale_log2, dscale,
                                                    p.drop_thresh, my_ckey, mrow, p.ms_q, q0 + c * 32, p.Lq, !key_masked, causal_key);
          else chunk_dkv_t<true, false>(rs, rd, s_lse + c * 32, s_delta + c * 32, s_key + c * 32, p.scale_log2, dscale,
                                        p.drop_thresh, my_ckey, mrow, p.ms_q, q0 + c * 32, p.Lq, !key_masked, causal_key);
        } else {
          if (p.drop_thresh) chunk_dkv_t<false, true>(rs, rd, s_lse + c * 32, s_delta + c * 32, s_key + c * 32, p.scale_log2, dscale,
                                                     p.drop_thresh, my_ckey, mrow, p.ms_q, q0 + c * 32, p.Lq, !key_masked, causal_key);
          else chunk_dkv_t<false, false>(rs, rd, s_lse + c * 32, s_delta + c * 32, s_key + c * 32, p.scale_log2, dscale,
                                         p.drop_thresh, my_ckey, mrow, p.ms_q, q0 + c * 32, p.Lq, !key_masked, causal_key);
        }
      }
      tmem_st32(t_lane + T_ST + c * 32, rs);
      tmem_st32(t_lane + T_DPT + c * 32, rd);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
      const uint32_t bdo = smem_u32(sDOm), bq = smem_u32(sQm);
#pragma unroll
      for (int ks = 0; ks < BQ / 8; ++ks)  // dV += P^T dO
        umma_tf32_ts(tmem + T_DV, tmem + T_ST + ks * 8, umma_desc_mnmajor(bdo + ks * 1024, BQ * 128), idesc,
                     (it > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < BQ / 8; ++ks)  // dK += dS^T Q
        umma_tf32_ts(tmem + T_DK, tmem + T_DPT + ks * 8, umma_desc_mnmajor(bq + ks * 1024, BQ * 128), idesc,
                     (it > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&bar_acc);
    }
  }
  mbar_wait(&bar_acc, (n_q - 1) & 1);
  tc_fence_after();
#pragma unroll 1
  for (int c = slice; c < DK / 32; c += NS) {
    uint32_t rv[32], rk[32];
    tmem_ld32(t_lane + T_DV + c * 32, rv);
    tmem_ld32(t_lane + T_DK + c * 32, rk);
    tmem_ld_wait();
    if (key_ok) {
      float* dvp = static_cast<float*>(p.dv) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddv + h * DK + c * 32;
      float* dkp = static_cast<float*>(p.dk) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddk + h * DK + c * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        *reinterpret_cast<float4*>(dvp + i) =
            make_float4(tf32_rna(__uint_as_float(rv[i]) * dscale), tf32_rna(__uint_as_float(rv[i + 1]) * dscale),
                        tf32_rna(__uint_as_float(rv[i + 2]) * dscale), tf32_rna(__uint_as_float(rv[i + 3]) * dscale));
        *reinterpret_cast<float4*>(dkp + i) =
            make_float4(tf32_rna(__uint_as_float(rk[i]) * p.scale), tf32_rna(__uint_as_float(rk[i + 1]) * p.scale),
                        tf32_rna(__uint_as_float(rk[i + 2]) * p.scale), tf32_rna(__uint_as_float(rk[i + 3]) * p.scale));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ backward: dQ
template <int DK, int BKV>
__global__ void __launch_bounds__(4 * BKV, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_k_k, const __grid_constant__ CUtensorMap tmap_k_mn,
                   const __grid_constant__ CUtensorMap tmap_v_k, const AttnDev p) {
  constexpr int BQ = 128;
  constexpr int NS = BKV / 32;
  constexpr int G = DK / 32;
  constexpr int Q_BYTES = BQ * DK * 4;
  constexpr int KT_BYTES = BKV * DK * 4;
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_S = 0, T_DP = BKV, T_DQ = 2 * BKV;
  static_assert(2 * BKV + DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + Q_BYTES;
  uint8_t* sKk = sDO + Q_BYTES;
  uint8_t* sKm = sKk + KT_BYTES;
  uint8_t* sVk = sKm + KT_BYTES;
  __shared__ uint64_t bar_q, bar_ld, bar_s, bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) uint32_t s_ckey[BKV];
  __shared__ uint32_t s_mb[2][BKV / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, slice = warp >> 2;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + quarter * 32 + lane;
  const bool row_ok = row < p.Lq;
  const int n_kv = (p.Lk + BKV - 1) / BKV;
  const bool shared_mask = mask_is_row_invariant(p);
  if (shared_mask && tid < BKV / 32) s_mb[0][tid] = mask_bits_row(p, b, 0, true, tid * 32);

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k_k);
    tma_prefetch_desc(&tmap_k_mn); tma_prefetch_desc(&tmap_v_k);
    mbar_init(&bar_q, 1); mbar_init(&bar_ld, 1); mbar_init(&bar_s, 1); mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_q, 2 * Q_BYTES);
      tma_load_4d(sQ, &tmap_q, &bar_q, 0, q0, h * G, b);
      tma_load_4d(sDO, &tmap_do, &bar_q, 0, q0, h * G, b);

  }
  const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (row_ok ? row : 0);
  const float lse2 = row_ok ? p.lse2[stat] : INFINITY;
  const float delta = row_ok ? p.delta[stat] : 0.f;
  const uint32_t drop_key = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(stat)) : 0u;
  const float dscale = p.drop_thresh ? p.drop_scale : 1.f;

  for (int j = 0; j < n_kv; ++j) {
    if (tid == 0) {
      if (j > 0) { mbar_wait(&bar_acc, (j - 1) & 1); tc_fence_after(); }
      mbar_arrive_expect_tx(&bar_ld, 3 * KT_BYTES);
        tma_load_4d(sKk, &tmap_k_k, &bar_ld, 0, j * BKV, h * G, b);
        tma_load_4d(sKm, &tmap_k_mn, &bar_ld, 0, j * BKV, h * G, b);
        tma_load_4d(sVk, &tmap_v_k, &bar_ld, 0, j * BKV, h * G, b);

      if (j == 0) mbar_wait(&bar_q, 0);
      mbar_wait(&bar_ld, j & 1);
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_tf32(128, BKV, false, false);
      const uint32_t aq = smem_u32(sQ), ado = smem_u32(sDO), bk = smem_u32(sKk), bv = smem_u32(sVk);
#pragma unroll
      for (int ks = 0; ks < DK / 8; ++ks)  // S = Q K^T
        umma_tf32_ss(tmem + T_S, umma_desc_kmajor(aq + (ks / 4) * (BQ * 128) + (ks % 4) * 32),
                     umma_desc_kmajor(bk + (ks / 4) * (BKV * 128) + (ks % 4) * 32), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < DK / 8; ++ks)  // dP = dO V^T
        umma_tf32_ss(tmem + T_DP, umma_desc_kmajor(ado + (ks / 4) * (BQ * 128) + (ks % 4) * 32),
                     umma_desc_kmajor(bv + (ks / 4) * (BKV * 128) + (ks % 4) * 32), idesc, ks > 0 ? 1u : 0u);
      umma_commit(&bar_s);
    }
    if (shared_mask && tid < BKV / 32 && j + 1 < n_kv)   // next tile's shared mask bits (read after the barrier below)
      s_mb[(j + 1) & 1][tid] = mask_bits_row(p, b, 0, true, (j + 1) * BKV + tid * 32);
    if (p.drop_thresh) {  // per-column dropout keys of this key tile (previous readers passed the barrier below)
      if (tid < BKV) s_ckey[tid] = dropout_col_key(p.drop_seed, static_cast<uint32_t>(j * BKV + tid));
      __syncthreads();
    }
    mbar_wait(&bar_s, j & 1);
    tc_fence_after();
    {
      const int c = slice;
      const int k0 = j * BKV + c * 32;
      const uint32_t mb = shared_mask ? s_mb[j & 1][c] : mask_bits_row(p, b, row, row_ok, k0);
      uint32_t rs[32], rd[32];
      tmem_ld32(t_lane + T_S + c * 32, rs);
      tmem_ld32(t_lane + T_DP + c * 32, rd);
      tmem_ld_wait();
      // dS without the softmax scale (applied once to dQ in the epilogue)
      chunk_ds(rs, rd, mb, p.scale_log2, lse2, delta, dscale, p.drop_thresh, drop_key, s_ckey + c * 32);
      tmem_st32(t_lane + T_DP + c * 32, rd);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {  // dQ += dS K   (A = dS in TMEM, B = K MN-major)
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
      const uint32_t bk = smem_u32(sKm);
#pragma unroll
      for (int ks = 0; ks < BKV / 8; ++ks)
        umma_tf32_ts(tmem + T_DQ, tmem + T_DP + ks * 8, umma_desc_mnmajor(bk + ks * 1024, BKV * 128), idesc,
                     (j > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&bar_acc);
    }
  }
  mbar_wait(&bar_acc, (n_kv - 1) & 1);
  tc_fence_after();
#pragma unroll 1
  for (int c = slice; c < DK / 32; c += NS) {
    uint32_t r[32];
    tmem_ld32(t_lane + T_DQ + c * 32, r);
    tmem_ld_wait();
    if (row_ok) {
      float* dst = static_cast<float*>(p.dq) + (static_cast<int64_t>(b) * p.Lq + row) * p.lddq + h * DK + c * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + i) =
            make_float4(tf32_rna(__uint_as_float(r[i]) * p.scale), tf32_rna(__uint_as_float(r[i + 1]) * p.scale),
                        tf32_rna(__uint_as_float(r[i + 2]) * p.scale), tf32_rna(__uint_as_float(r[i + 3]) * p.scale));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

}  // namespace

// ================================================================================ host side
// 4-D tensor map over a (B, L, H*dk) activation addressed as rows of `ld` floats, viewed as
// {32 columns, L rows, H*dk/32 column groups, B}: one box {32, box_rows, dk/32, 1} fetches a whole head tile and lands it
// as [column group][row][32 floats] — the canonical 128-byte-swizzled UMMA operand layout — with ONE TMA instruction.
int make_act_tmap(CUtensorMap* m, const void* base, int64_t ld, int cols, int L, int B, int box_rows, int atom32, int dk) {
  const uint64_t dims[4] = {32, static_cast<uint64_t>(L), static_cast<uint64_t>(cols / 32), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ld) * 4, 128, static_cast<uint64_t>(L) * static_cast<uint64_t>(ld) * 4};
  const uint32_t box[4] = {32, static_cast<uint32_t>(box_rows), static_cast<uint32_t>(dk / 32), 1};
  return make_tmap_f32(m, base, 4, dims, strides, box, atom32);
}

int make_act_tmap16(CUtensorMap* m, int dtype, const void* base, int64_t ld, int cols, int L, int B, int box_rows, int dk) {
  const uint64_t dims[4] = {64, static_cast<uint64_t>(L), static_cast<uint64_t>(cols / 64), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, 128, static_cast<uint64_t>(L) * static_cast<uint64_t>(ld) * 2};
  const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), static_cast<uint32_t>(dk / 64), 1};
  return make_tmap(m, dtype, base, 4, dims, strides, box, 0);
}

int check_attn(const AttnArgs& a, const char* who) {
  ST_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "%s: empty problem", who);
  ST_REQUIRE(a.dk == 32 || a.dk == 64 || a.dk == 128, "%s: d_k must be 32, 64 or 128 (got %d)", who, a.dk);
  ST_REQUIRE((a.ldq & 3) == 0 && (a.ldk & 3) == 0 && (a.ldv & 3) == 0 && (a.ldctx & 3) == 0,
             "%s: leading dimensions must be multiples of 4", who);
  ST_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
               reinterpret_cast<uintptr_t>(a.ctx)) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  return ST_OK;
}

// Measured on B200 (6+6 x 512, B=32, T=1000): accumulating the bias column sums from the dK/dV and dQ epilogues costs more
// (same-address L2 reductions from 2048 CTAs: dkv +14 %, dq +5 %) than the separate 35 us column-sum pass it replaces,
// so it is opt-in (option "attn_fuse_bias"); the GEMM-epilogue fusion of the FFN bias gradient is always on.
int attn_read_fwd_trace(unsigned long long* host_out, int n) {
  const int total = FTRACE_TILES * FTRACE_EVENTS;
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  ST_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_ftrace, sizeof(unsigned long long) * (n < total ? n : total)));
  return total;
}

bool attn_bwd_fuses_bias(int dk) { return get_option("attn_fuse_bias") && dk <= 64 && !get_option("attn_bwd_simple"); }

AttnDev attn_to_dev(const AttnArgs& a) {
  AttnDev p{};
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk;
  p.mask = a.mask; p.ms_b = a.ms_b; p.ms_q = a.ms_q; p.ms_k = a.ms_k;
  p.k_len = a.k_len; p.causal = a.causal; p.ds_boost = 1.f;
  p.scale = a.scale; p.scale_log2 = a.scale * kLog2e;
  p.drop_thresh = a.drop.thresh32; p.drop_scale = a.drop.scale32; p.drop_seed = a.drop.seed;
  p.ctx = a.ctx; p.ldctx = a.ldctx; p.lse2 = a.lse; p.attn = a.attn;
  return p;
}

namespace {

template <int DK, int BKV>
int launch_fwd(cudaStream_t s, const AttnArgs& a) {
  CUtensorMap tq, tk, tv;
  const int cols = a.H * DK;
  ST_TRY(make_act_tmap(&tq, a.q, a.ldq, cols, a.Lq, a.B, 128, 0, DK));
  ST_TRY(make_act_tmap(&tk, a.k, a.ldk, cols, a.Lk, a.B, BKV, 0, DK));
  ST_TRY(make_act_tmap(&tv, a.v, a.ldv, cols, a.Lk, a.B, BKV, 1, DK));
  constexpr int SMEM = 128 * DK * 4 + 2 * BKV * DK * 4 + 1024;
  auto kern = attn_fwd_kernel<DK, BKV>;
  static bool attr = false;
  if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
  dim3 grid((a.Lq + 127) / 128, a.H, a.B);
  ProfScope prof(s, PROF_ATTN_FWD, 4.0 * a.B * a.H * static_cast<double>(a.Lq) * a.Lk * DK);
  AttnDev dev = attn_to_dev(a);
  dev.trace = get_option("attn_trace");
  ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(256), SMEM, s, tq, tk, tv, dev));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

template <int DK, int BQ, int BKV>
int launch_bwd(cudaStream_t s, const AttnBwdArgs& a) {
  const AttnArgs& f = a.f;
  const int cols = f.H * DK;
  AttnDev p = attn_to_dev(f);
  p.trace = get_option("attn_trace");
  p.dbq = a.dbq; p.dbk = a.dbk; p.dbv = a.dbv;
  p.delta = a.delta; p.dq = a.dq; p.lddq = a.lddq; p.dk = a.dk_; p.lddk = a.lddk; p.dv = a.dv; p.lddv = a.lddv;
  const bool pipelined = DK <= 64 && !get_option("attn_bwd_simple");
  {   // delta = rowsum(dO * O): a separate HBM pass (forming it inside the dQ kernel measured slower)
    const int64_t rows = static_cast<int64_t>(f.B) * f.Lq;
    const int64_t blocks = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
    ProfScope prof(s, PROF_ATTN_DELTA, 2.0 * rows * cols * 4);
    ST_CHECK_CUDA(launch_pdl(attn_delta_kernel, dim3(static_cast<unsigned>(blocks < cap ? blocks : cap)), dim3(256), 0, s,
                             static_cast<const float*>(a.dctx), a.lddctx, static_cast<const float*>(f.ctx), f.ldctx, a.delta, f.B,
                             f.H, f.Lq, DK));
    ST_CHECK_LAUNCH();
  }
  if (pipelined) return attn_bwd_pipelined(s, a, p);  // st_attn_bwd.cu
  {
    CUtensorMap tk, tv, tqk, tqm, tdk, tdm;
    ST_TRY(make_act_tmap(&tk, f.k, f.ldk, cols, f.Lk, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tv, f.v, f.ldv, cols, f.Lk, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tqk, f.q, f.ldq, cols, f.Lq, f.B, BQ, 0, DK));
    ST_TRY(make_act_tmap(&tqm, f.q, f.ldq, cols, f.Lq, f.B, BQ, 1, DK));
    ST_TRY(make_act_tmap(&tdk, a.dctx, a.lddctx, cols, f.Lq, f.B, BQ, 0, DK));
    ST_TRY(make_act_tmap(&tdm, a.dctx, a.lddctx, cols, f.Lq, f.B, BQ, 1, DK));
    constexpr int SMEM = 2 * 128 * DK * 4 + 4 * BQ * DK * 4 + 1024;
    auto kern = attn_bwd_dkv_kernel<DK, BQ>;
    static bool attr = false;
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    dim3 grid((f.Lk + 127) / 128, f.H, f.B);
    // algorithmic share of the attention backward carried by this kernel: dV and dK (S, dP recompute not counted)
    ProfScope prof(s, PROF_ATTN_DKV, 4.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    kern<<<grid, 4 * BQ, SMEM, s>>>(tk, tv, tqk, tqm, tdk, tdm, p);
    ST_CHECK_LAUNCH();
  }
  {
    CUtensorMap tq, tdo, tkk, tkm, tvk;
    ST_TRY(make_act_tmap(&tq, f.q, f.ldq, cols, f.Lq, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tdo, a.dctx, a.lddctx, cols, f.Lq, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tkk, f.k, f.ldk, cols, f.Lk, f.B, BKV, 0, DK));
    ST_TRY(make_act_tmap(&tkm, f.k, f.ldk, cols, f.Lk, f.B, BKV, 1, DK));
    ST_TRY(make_act_tmap(&tvk, f.v, f.ldv, cols, f.Lk, f.B, BKV, 0, DK));
    constexpr int SMEM = 2 * 128 * DK * 4 + 3 * BKV * DK * 4 + 1024;
    auto kern = attn_bwd_dq_kernel<DK, BKV>;
    static bool attr = false;
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    dim3 grid((f.Lq + 127) / 128, f.H, f.B);
    // algorithmic share: dQ plus the (single) S and dP products of the textbook backward
    ProfScope prof(s, PROF_ATTN_DQ, 6.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    kern<<<grid, 4 * BKV, SMEM, s>>>(tq, tdo, tkk, tkm, tvk, p);
    ST_CHECK_LAUNCH();
  }
  return ST_OK;
}

}  // namespace

int attn_fwd(cudaStream_t s, const AttnArgs& a) {
  if (a.dtype != ST_DTYPE_F32) return attn16_fwd(s, a);
  ST_TRY(check_attn(a, "attn_fwd"));
  ST_REQUIRE(a.lse != nullptr, "attn_fwd: lse buffer is required");
  switch (a.dk) {
    case 32: return launch_fwd<32, 128>(s, a);
    case 64: return launch_fwd<64, 128>(s, a);
    default: return launch_fwd<128, 64>(s, a);
  }
}

int attn_bwd(cudaStream_t s, const AttnBwdArgs& a) {
  if (a.f.dtype != ST_DTYPE_F32) return attn16_bwd(s, a);
  ST_TRY(check_attn(a.f, "attn_bwd"));
  ST_REQUIRE(a.dctx && a.delta && a.dq && a.dk_ && a.dv && a.f.lse, "attn_bwd: null buffer");
  ST_REQUIRE((a.lddctx & 3) == 0 && (a.lddq & 3) == 0 && (a.lddk & 3) == 0 && (a.lddv & 3) == 0,
             "attn_bwd: leading dimensions must be multiples of 4");
  switch (a.f.dk) {
    case 32: return launch_bwd<32, 128, 128>(s, a);
    case 64: return launch_bwd<64, 128, 128>(s, a);
    default: return launch_bwd<128, 32, 32>(s, a);
  }
}

}  // namespace st
