// st_attn_bwd.cu — warp-specialised, software-pipelined attention backward for d_k in {32, 64}.
//
// Same mathematics as the simple kernels in st_attn.cu (autograd of transformer/Attention.py:82-90), but
// organised like the GEMM: the CTA has a TMA-producer warp, a single-thread tcgen05.mma issuer warp and
// 16 compute warps that only ever wait on mbarriers (no __syncthreads in the loop):
//
//   resident : the CTA's own 128-row tiles (Q, dO in the dQ kernel; K, V in the dK/dV kernel) are the A operands
//              of the "recompute" MMAs.  They are written ONCE into TMEM by the compute threads (lane = row,
//              column = feature) and consumed with the .ts MMA form, so they occupy no shared memory and are not
//              re-read from smem by every MMA
//   producer : multi-stage smem ring (4 / 3 stages) of the streamed 64-row tiles, one 4-D TMA box per operand copy,
//              + per-tile metadata (dropout keys, mask bits, log-sum-exp / delta), released by tcgen05.commit
//   MMA warp : A(t) = the two recompute MMAs of tile t into TMEM buffer t&1  (S, dP  or  S^T, dP^T),
//              B(t) = the gradient MMAs consuming what the compute warps wrote back (A operand in TMEM);
//              issue order A(0) A(1) B(0) A(2) B(1) ... so A(t+1) runs under the compute phase of tile t
//   compute  : 4 threads per tile row (16 columns each): tcgen05.ld -> exp2 / dropout / dS -> tcgen05.st,
//              then arrive on the "written back" barrier
//
//   dQ  kernel: one CTA per 128-query tile, streams 64-key tiles:   dQ += dS K
//   dKV kernel: one CTA per 128-key tile,   streams 64-query tiles: dV += P^T dO,  dK += dS^T Q
// The kernels are bound by the bytes streamed into each SM (K-major + MN-major copies of every streamed tile);
// the deep ring keeps those loads in flight behind the MMAs and the softmax arithmetic.
// TMEM (512 columns): 2 x (64 + 64) double-buffered S/dP pair, the fp32 gradient accumulators, the resident tiles.
#include <utility>

#include "st_attn.cuh"

namespace st {

namespace {

// Optional in-kernel timeline (option "attn_trace"): CTA (0,0,0) of the dK/dV kernel records clock64() at its pipeline
// hand-offs — slot layout [role][tile][event]; read back with st_debug_read_trace (tools/trace_attn.py).
constexpr int TRACE_TILES = 24, TRACE_EVENTS = 6;
__device__ unsigned long long g_trace[2 * TRACE_TILES * TRACE_EVENTS];
#define ST_TRACE(role, tile, ev)                                                                      \
  do {                                                                                                \
    if (trace_on && (tile) < TRACE_TILES) g_trace[((role) * TRACE_TILES + (tile)) * TRACE_EVENTS + (ev)] = clock64(); \
  } while (0)

// per-CTA wall-clock record of the dK/dV kernel (option "attn_trace"): {globaltimer at entry, at exit, SM id, main-loop cycles}
constexpr int TRACE_CTAS = 4096;
__device__ unsigned long long g_cta_trace[TRACE_CTAS * 4];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t sm_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %smid;" : "=r"(r));
  return r;
}

// compile-time loop: f(IC<0>{}), f(IC<1>{}), ... — the MMA-issuer loops are unrolled over lcm(ring stages, 2) tiles so that
// the ring stage and the TMEM buffer of every phase are compile-time constants: descriptors and barrier addresses become
// `base + constant`, which keeps the issuing warp's instruction count (it shares a scheduler with four compute warps that
// saturate it) to a few dozen per phase instead of ~150.
template <int V> struct IC { static constexpr int value = V; };
template <class F, int... I>
__device__ __forceinline__ void cfor_impl(F& f, std::integer_sequence<int, I...>) { (f(IC<I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void cfor(F& f) { cfor_impl(f, std::make_integer_sequence<int, N>{}); }

constexpr int BT = 64;                 // streamed tile height
constexpr int NCOMP = 512;             // compute threads
constexpr int NTHREADS = NCOMP + 64;   // + producer warp + MMA warp
constexpr int W_PROD = NCOMP / 32;     // warp 16
constexpr int W_MMA = W_PROD + 1;      // warp 17

// dS chunk (dQ kernel; thread = query row, 16 key columns)
template <bool MASK, bool DROP>
__device__ __forceinline__ void ds16_t(const uint32_t (&rs)[16], uint32_t (&rd)[16], uint32_t mb, float scale_log2, float lse2,
                                       float delta, float dscale, uint32_t thresh, uint32_t rowkey, const uint32_t* ckey) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
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

// P^T / dS^T chunk (dKV kernel; thread = key row, 16 query columns)
template <bool DENSE, bool DROP>
__device__ __forceinline__ void dkv16_t(uint32_t (&rs)[16], uint32_t (&rd)[16], const float* lse, const float* delta,
                                        const uint32_t* rkey, float scale_log2, float dscale, uint32_t thresh,
                                        uint32_t my_ckey, const uint8_t* mrow, int64_t ms_q, int q_first, int Lq, bool key_ok,
                                        int causal_key = -1) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
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

// Column sums of a 128-row output tile (this thread: one row, 16 consecutive columns) added to out[0..16): the bias
// gradient of the projection that produced the attention input.  Epilogue only (once per CTA).
__device__ __forceinline__ void bias_colsum16(const float (&v)[16], float* out, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == i) atomicAdd(out + i, s);
  }
}

// Resident tiles -> TMEM: compute slice `slice` (0..3) owns one 32-column chunk of one of the two resident tensors
// (x0 at columns [t_x0, t_x0+DK), x1 at [t_x1, ...)); the thread writes its row's 32 floats (zeros for rows >= L).
template <int DK>
__device__ __forceinline__ void resident_to_tmem(const float* x0, int64_t ld0, const float* x1, int64_t ld1, int64_t row,
                                                 bool row_ok, int h, int slice, uint32_t t_lane, uint32_t t_x0, uint32_t t_x1) {
  constexpr int CH = DK / 32;            // 32-column chunks per tensor
  if (slice < 2 * CH) {                  // warp-uniform
    const int which = slice / CH, c = slice % CH;
    const float* src = (which == 0 ? x0 + row * ld0 : x1 + row * ld1) + h * DK + c * 32;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) v = __ldg(reinterpret_cast<const float4*>(src + i));
      r[i] = __float_as_uint(v.x); r[i + 1] = __float_as_uint(v.y); r[i + 2] = __float_as_uint(v.z); r[i + 3] = __float_as_uint(v.w);
    }
    tmem_st32(t_lane + (which == 0 ? t_x0 : t_x1) + c * 32, r);
    tmem_st_wait();
  }
}

// The same in two halves, so that a kernel can issue the global loads early (they then fly under its mask scan and
// start-up barrier) and write the rows to TMEM once the allocation is known.
template <int DK>
__device__ __forceinline__ void resident_load(const float* x0, int64_t ld0, const float* x1, int64_t ld1, int64_t row, bool row_ok,
                                              int h, int slice, uint32_t (&r)[32]) {
  constexpr int CH = DK / 32;
  if (slice < 2 * CH) {                  // warp-uniform
    const int which = slice / CH, c = slice % CH;
    const float* src = (which == 0 ? x0 + row * ld0 : x1 + row * ld1) + h * DK + c * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) v = __ldg(reinterpret_cast<const float4*>(src + i));
      r[i] = __float_as_uint(v.x); r[i + 1] = __float_as_uint(v.y); r[i + 2] = __float_as_uint(v.z); r[i + 3] = __float_as_uint(v.w);
    }
  }
}
template <int DK>
__device__ __forceinline__ void resident_store(int slice, uint32_t t_lane, uint32_t t_x0, uint32_t t_x1, const uint32_t (&r)[32]) {
  constexpr int CH = DK / 32;
  if (slice < 2 * CH) {
    const int which = slice / CH, c = slice % CH;
    tmem_st32(t_lane + (which == 0 ? t_x0 : t_x1) + c * 32, r);
    tmem_st_wait();
  }
}

// ================================================================================ dQ
// RS = true: the resident Q / dO tiles live in SHARED memory (one TMA box each) and the recompute MMAs use the .ss form.
// The tensor-memory read port (tcgen05.ld of S / dP by the compute warps + the A operands of .ts MMAs) is the busiest
// resource of these kernels; with N = 64 a .ts MMA re-reads a 4 KB A slice from TMEM for only 64 output columns.
template <int DK, bool RS>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dq_pipe(const float* __restrict__ q, int64_t ldq, const float* __restrict__ dctx, int64_t lddctx,
                 const __grid_constant__ CUtensorMap tmap_k_k, const __grid_constant__ CUtensorMap tmap_k_mn,
                 const __grid_constant__ CUtensorMap tmap_v_k, const __grid_constant__ CUtensorMap tmap_q_res,
                 const __grid_constant__ CUtensorMap tmap_do_res, const AttnDev p) {
  constexpr int BQ = 128;
  constexpr int G = DK / 32;
  constexpr int STAGES = RS ? 3 : 4;
  constexpr int T_BYTES = BT * DK * 4;
  constexpr int STAGE_BYTES = 3 * T_BYTES;   // Kk | Km | Vk
  constexpr int RES_BYTES = BQ * DK * 4;     // one resident tile (RS only)
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_S = 0, T_DP = 2 * BT, T_DQ = 4 * BT, T_Q = 4 * BT + DK, T_DO = 4 * BT + 2 * DK;
  static_assert(4 * BT + 3 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sBase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRes = sBase;                                   // RS: Q | dO resident tiles (K-major, 128B swizzle)
  uint8_t* sStage = sBase + (RS ? 2 * RES_BYTES : 0);
  __shared__ uint64_t res_ready, ld_full[STAGES], ld_empty[STAGES], s_full[2], ds_full[2], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) uint32_t s_ckey[STAGES][BT];
  __shared__ uint32_t s_mb[STAGES][BT / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  __shared__ int s_extent;

  if (tid == 0) {
    mbar_init(&res_ready, RS ? 1 : NCOMP); mbar_init(&acc_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ld_full[s], 1); mbar_init(&ld_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], NCOMP); }
    fence_mbar_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  pdl_wait();      // nothing above reads or writes global memory (programmatic dependent launch, st_host.h)
  pdl_trigger();

  // the compute threads' resident Q / dO rows: loads issued now, written to TMEM after the barrier below — this kernel runs
  // 14 CTAs per SM back to back, so a microsecond of start-up latency is paid 14 times per launch
  uint32_t rres[32];
  const int r_row = q0 + (warp & 3) * 32 + lane;
  const bool r_ok = r_row < p.Lq;
  if (!RS && warp < W_PROD)
    resident_load<DK>(q, ldq, dctx, lddctx, static_cast<int64_t>(b) * p.Lq + (r_ok ? r_row : 0), r_ok, h, warp >> 2, rres);

  const int extent = block_key_extent(p, b, &s_extent);
  // key tiles entirely inside this utterance's padding have dS == 0: skipped (see block_key_extent)
  const int n_kv = extent >= p.Lk ? (p.Lk + BT - 1) / BT : max(1, (extent + BT - 1) / BT);
  const bool shared_mask = mask_is_row_invariant(p);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == W_PROD) {
    // ===================== producer =====================
    if (lane == 0) { tma_prefetch_desc(&tmap_k_k); tma_prefetch_desc(&tmap_k_mn); tma_prefetch_desc(&tmap_v_k); }
    if (RS && lane == 0) {
      mbar_arrive_expect_tx(&res_ready, 2 * RES_BYTES);
      tma_load_4d(sRes, &tmap_q_res, &res_ready, 0, q0, h * G, b);              // rows >= Lq are zero-filled by TMA
      tma_load_4d(sRes + RES_BYTES, &tmap_do_res, &res_ready, 0, q0, h * G, b);
    }
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
        tma_load_4d(st, &tmap_k_k, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + T_BYTES, &tmap_k_mn, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + 2 * T_BYTES, &tmap_v_k, &ld_full[s], 0, t * BT, h * G, b);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    // The WARP stays converged (every lane waits on the barriers); one lane, chosen by elect.sync, issues.  Issuing
    // under `lane == 0` instead makes the compiler wrap every tcgen05.mma in an ELECT / BRA.U.ANY loop (~60-120 cycles
    // per MMA, measured: tools/mma_bench.py), which starves the tensor pipe for these N = 64 MMAs (floor 32 cycles).
    {
      const uint32_t st0 = smem_u32(sStage);
      const uint32_t rq = smem_u32(sRes), rdo = rq + RES_BYTES;
      auto koff = [](int ks, int rows) { return static_cast<uint64_t>(((ks / 4) * (rows * 128) + (ks % 4) * 32) >> 4); };
      constexpr int U = (STAGES % 2 == 0) ? STAGES : 2 * STAGES;   // tiles per unrolled round: stage and buffer indices are constants
      const uint64_t dk0 = umma_desc_kmajor(st0), dkm0 = umma_desc_mnmajor(st0 + T_BYTES, BT * 128), dv0 = umma_desc_kmajor(st0 + 2 * T_BYTES);
      const uint64_t aq0 = umma_desc_kmajor(rq), ado0 = umma_desc_kmajor(rdo);
      auto issue_a = [&](uint32_t ld_parity, auto S, auto TB) {   // S(t) = Q K^T, dP(t) = dO V^T   (A = resident tile, B K-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ld_full[s], ld_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, BT, false, false);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks) {
            if (RS) umma_tf32_ss(tmem + T_S + tb * BT, aq0 + koff(ks, BQ), dk0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
            else umma_tf32_ts(tmem + T_S + tb * BT, tmem + T_Q + ks * 8, dk0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks) {
            if (RS) umma_tf32_ss(tmem + T_DP + tb * BT, ado0 + koff(ks, BQ), dv0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
            else umma_tf32_ts(tmem + T_DP + tb * BT, tmem + T_DO + ks * 8, dv0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[tb]);
        }
        __syncwarp();
      };
      auto issue_b = [&](uint32_t ds_parity, bool first, auto S, auto TB) {   // dQ += dS(t) K(t)   (A = dS in TMEM, B = K MN-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        mbar_wait(&ds_full[tb], ds_parity);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < BT / 8; ++ks)
            umma_tf32_ts(tmem + T_DQ, tmem + T_DP + tb * BT + ks * 8, dkm0 + so + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
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
    if (!RS) {
      resident_store<DK>(slice, t_lane, T_Q, T_DO, rres);
      tc_fence_before();
      mbar_arrive(&res_ready);
    }
    const float delta = row_ok ? p.delta[stat] : 0.f;
    const float lse2 = row_ok ? p.lse2[stat] : INFINITY;
    const uint32_t drop_key = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(stat)) : 0u;
    const float dscale = p.drop_thresh ? p.drop_scale : 1.f;
    for (int t = 0; t < n_kv; ++t) {
      const int s = t % STAGES, tb = t & 1;
      mbar_wait(&s_full[tb], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t word = shared_mask ? s_mb[s][col0 >> 5] : mask_bits_row(p, b, row, row_ok, t * BT + (col0 & ~31));
      const uint32_t mb = (word >> (col0 & 31)) & 0xFFFFu;
      uint32_t rs[16], rd[16];
      tmem_ld16(t_lane + T_S + tb * BT + col0, rs);
      tmem_ld16(t_lane + T_DP + tb * BT + col0, rd);
      tmem_ld_wait();
      const uint32_t* ck = s_ckey[s] + col0;
      if (mb == 0u) {
        if (p.drop_thresh) ds16_t<false, true>(rs, rd, mb, p.scale_log2, lse2, delta, dscale, p.drop_thresh, drop_key, ck);
        else ds16_t<false, false>(rs, rd, mb, p.scale_log2, lse2, delta, dscale, p.drop_thresh, drop_key, ck);
      } else {
        if (p.drop_thresh) ds16_t<true, true>(rs, rd, mb, p.scale_log2, lse2, delta, dscale, p.drop_thresh, drop_key, ck);
        else ds16_t<true, false>(rs, rd, mb, p.scale_log2, lse2, delta, dscale, p.drop_thresh, drop_key, ck);
      }
      tmem_st16(t_lane + T_DP + tb * BT + col0, rd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[tb]);
    }
    // ---- epilogue: dQ = scale * accumulator
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    if (col0 < DK) {
      uint32_t r[16];
      tmem_ld16(t_lane + T_DQ + col0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = row_ok ? __uint_as_float(r[i]) * p.scale : 0.f;
      if (row_ok) {
        float* dst = static_cast<float*>(p.dq) + (static_cast<int64_t>(b) * p.Lq + row) * p.lddq + h * DK + col0;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(tf32_rna(v[i]), tf32_rna(v[i + 1]), tf32_rna(v[i + 2]), tf32_rna(v[i + 3]));
      }
      if (p.dbq) bias_colsum16(v, p.dbq + h * DK + col0, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ================================================================================ dK, dV
template <int DK, bool RS>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dkv_pipe(const float* __restrict__ k, int64_t ldk, const float* __restrict__ v, int64_t ldv,
                  const __grid_constant__ CUtensorMap tmap_q_k, const __grid_constant__ CUtensorMap tmap_q_mn,
                  const __grid_constant__ CUtensorMap tmap_do_k, const __grid_constant__ CUtensorMap tmap_do_mn,
                  const __grid_constant__ CUtensorMap tmap_k_res, const __grid_constant__ CUtensorMap tmap_v_res,
                  const AttnDev p) {
  constexpr int BKV = 128;
  constexpr int G = DK / 32;
  constexpr int STAGES = (RS && DK == 64) ? 2 : 3;
  constexpr int T_BYTES = BT * DK * 4;
  constexpr int STAGE_BYTES = 4 * T_BYTES;   // Qk | Qm | dOk | dOm
  constexpr int RES_BYTES = BKV * DK * 4;
  constexpr uint32_t TCOLS = 512;
  constexpr uint32_t T_ST = 0, T_DPT = 2 * BT, T_DV = 4 * BT, T_DK = 4 * BT + DK, T_K = 4 * BT + 2 * DK, T_V = 4 * BT + 3 * DK;
  static_assert(4 * BT + 4 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sBase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRes = sBase;                                   // RS: K | V resident tiles (K-major, 128B swizzle)
  uint8_t* sStage = sBase + (RS ? 2 * RES_BYTES : 0);
  __shared__ uint64_t res_ready, ld_full[STAGES], ld_empty[STAGES], s_full[2], ds_full[2], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[STAGES][BT];
  __shared__ __align__(16) float s_delta[STAGES][BT];
  __shared__ __align__(16) uint32_t s_rkey[STAGES][BT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.Lq + BT - 1) / BT;
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  pdl_wait();      // before the first global access and before the early exit below (programmatic dependent launch)
  pdl_trigger();
  if (p.trace && tid == 0 && cta_lin < TRACE_CTAS) { g_cta_trace[cta_lin * 4] = global_ns(); g_cta_trace[cta_lin * 4 + 2] = sm_id(); }
  {
    // a key tile entirely inside this utterance's padding: dK = dV = 0 for its rows, nothing to compute
    __shared__ int s_extent;
    const int extent = block_key_extent(p, b, &s_extent);
    if (extent > 0 && kv0 >= extent) {
      const int rows = min(BKV, p.Lk - kv0);
      for (int i = tid; i < rows * (DK / 4); i += NTHREADS) {
        const int r = i / (DK / 4), c = (i - r * (DK / 4)) * 4;
        const int64_t grow = static_cast<int64_t>(b) * p.Lk + kv0 + r;
        *reinterpret_cast<float4*>(static_cast<float*>(p.dk) + grow * p.lddk + h * DK + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(static_cast<float*>(p.dv) + grow * p.lddv + h * DK + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (p.trace && tid == 0 && cta_lin < TRACE_CTAS) g_cta_trace[cta_lin * 4 + 1] = global_ns();
      return;
    }
  }
  const bool trace_on = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  if (tid == 0) {
    mbar_init(&res_ready, RS ? 1 : NCOMP); mbar_init(&acc_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ld_full[s], 1); mbar_init(&ld_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], NCOMP); }
    fence_mbar_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == W_PROD) {
    // ===================== producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q_k); tma_prefetch_desc(&tmap_q_mn); tma_prefetch_desc(&tmap_do_k); tma_prefetch_desc(&tmap_do_mn);
      if (RS) {
        mbar_arrive_expect_tx(&res_ready, 2 * RES_BYTES);
        tma_load_4d(sRes, &tmap_k_res, &res_ready, 0, kv0, h * G, b);             // rows >= Lk are zero-filled by TMA
        tma_load_4d(sRes + RES_BYTES, &tmap_v_res, &res_ready, 0, kv0, h * G, b);
      }
    }
    for (int t = 0; t < n_q; ++t) {
      const int s = t % STAGES;
      mbar_wait(&ld_empty[s], ((t / STAGES) & 1) ^ 1);
#pragma unroll
      for (int e = lane; e < BT; e += 32) {   // per-query statistics of this tile
        const int q = t * BT + e;
        const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Lq + (q < p.Lq ? q : 0);
        s_lse[s][e] = q < p.Lq ? p.lse2[o] : INFINITY;   // +inf => probability 0 for padded query rows
        s_delta[s][e] = q < p.Lq ? p.delta[o] : 0.f;
        s_rkey[s][e] = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(o)) : 0u;
      }
      __syncwarp();
      if (lane == 0) {
        uint8_t* st = sStage + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&ld_full[s], STAGE_BYTES);
        tma_load_4d(st, &tmap_q_k, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + T_BYTES, &tmap_q_mn, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + 2 * T_BYTES, &tmap_do_k, &ld_full[s], 0, t * BT, h * G, b);
        tma_load_4d(st + 3 * T_BYTES, &tmap_do_mn, &ld_full[s], 0, t * BT, h * G, b);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (converged warp + elect.sync, see the dQ kernel) =====================
    {
      const uint32_t st0 = smem_u32(sStage);
      const uint32_t rk = smem_u32(sRes), rv = rk + RES_BYTES;
      auto koff = [](int ks, int rows) { return static_cast<uint64_t>(((ks / 4) * (rows * 128) + (ks % 4) * 32) >> 4); };
      constexpr int U = (STAGES % 2 == 0) ? STAGES : 2 * STAGES;   // tiles per unrolled round (see the dQ kernel)
      const uint64_t dq0 = umma_desc_kmajor(st0), dqm0 = umma_desc_mnmajor(st0 + T_BYTES, BT * 128);
      const uint64_t ddo0 = umma_desc_kmajor(st0 + 2 * T_BYTES), ddom0 = umma_desc_mnmajor(st0 + 3 * T_BYTES, BT * 128);
      const uint64_t ak0 = umma_desc_kmajor(rk), av0 = umma_desc_kmajor(rv);
      auto issue_a = [&](int t, uint32_t ld_parity, auto S, auto TB) {   // S^T(t) = K Q^T, dP^T(t) = V dO^T   (A = resident tile, B K-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        ST_TRACE(0, t, 0);
        mbar_wait(&ld_full[s], ld_parity);
        ST_TRACE(0, t, 1);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, BT, false, false);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks) {
            if (RS) umma_tf32_ss(tmem + T_ST + tb * BT, ak0 + koff(ks, BKV), dq0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
            else umma_tf32_ts(tmem + T_ST + tb * BT, tmem + T_K + ks * 8, dq0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks) {
            if (RS) umma_tf32_ss(tmem + T_DPT + tb * BT, av0 + koff(ks, BKV), ddo0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
            else umma_tf32_ts(tmem + T_DPT + tb * BT, tmem + T_V + ks * 8, ddo0 + so + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[tb]);
        }
        __syncwarp();
        ST_TRACE(0, t, 2);
      };
      auto issue_b = [&](int t, uint32_t ds_parity, auto S, auto TB) {   // dV += P^T dO, dK += dS^T Q   (A in TMEM, B MN-major)
        constexpr int s = decltype(S)::value, tb = decltype(TB)::value;
        ST_TRACE(0, t, 3);
        mbar_wait(&ds_full[tb], ds_parity);
        ST_TRACE(0, t, 4);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
          constexpr uint64_t so = static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < BT / 8; ++ks)
            umma_tf32_ts(tmem + T_DV, tmem + T_ST + tb * BT + ks * 8, ddom0 + so + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
                         (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < BT / 8; ++ks)
            umma_tf32_ts(tmem + T_DK, tmem + T_DPT + tb * BT + ks * 8, dqm0 + so + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
                         (t > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&ld_empty[s]);
        }
        __syncwarp();
        ST_TRACE(0, t, 5);
      };
      mbar_wait(&res_ready, 0);
      tc_fence_after();
      issue_a(0, 0u, IC<0>{}, IC<0>{});
      for (int t0 = 0; t0 < n_q; t0 += U) {
        const uint32_t ldp = static_cast<uint32_t>(t0 / STAGES), dsp = static_cast<uint32_t>(t0 >> 1);
        auto round = [&](auto Uc) {
          constexpr int u = decltype(Uc)::value;
          const int t = t0 + u;
          if (t < n_q) {
            if (t + 1 < n_q) issue_a(t + 1, (ldp + (u + 1) / STAGES) & 1u, IC<(u + 1) % STAGES>{}, IC<(u + 1) & 1>{});
            issue_b(t, (dsp + (u >> 1)) & 1u, IC<u % STAGES>{}, IC<(u & 1)>{});
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
    if (!RS) {
      resident_to_tmem<DK>(k, ldk, v, ldv, grow, key_ok, h, slice, t_lane, T_K, T_V);
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
    for (int t = 0; t < n_q; ++t) {
      const int s = t % STAGES, tb = t & 1;
      if (tid == 0) ST_TRACE(1, t, 0);
      mbar_wait(&s_full[tb], (t >> 1) & 1);
      if (tid == 0) ST_TRACE(1, t, 1);
      tc_fence_after();
      uint32_t rs[16], rd[16];
      // tcgen05.ld/st are warp-collective (.sync.aligned): every lane executes them, whatever its key's mask state
      tmem_ld16(t_lane + T_ST + tb * BT + col0, rs);
      tmem_ld16(t_lane + T_DPT + tb * BT + col0, rd);
      tmem_ld_wait();
      if (tid == 0) ST_TRACE(1, t, 2);
      if (!mask_dense && key_masked) {   // this key is padding for every query: P = dS = 0
#pragma unroll
        for (int i = 0; i < 16; ++i) { rs[i] = 0u; rd[i] = 0u; }
      } else {
        const float* ls = s_lse[s] + col0;
        const float* de = s_delta[s] + col0;
        const uint32_t* rk = s_rkey[s] + col0;
        const int qf = t * BT + col0;
        if (mask_dense) {
          if (p.drop_thresh) dkv16_t<true, true>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
          else dkv16_t<true, false>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
        } else {
          if (p.drop_thresh) dkv16_t<false, true>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
          else dkv16_t<false, false>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, qf, p.Lq, !key_masked, causal_key);
        }
      }
      if (tid == 0) ST_TRACE(1, t, 3);
      tmem_st16(t_lane + T_ST + tb * BT + col0, rs);
      tmem_st16(t_lane + T_DPT + tb * BT + col0, rd);
      tmem_st_wait();
      if (tid == 0) ST_TRACE(1, t, 4);
      tc_fence_before();
      mbar_arrive(&ds_full[tb]);
      if (tid == 0) ST_TRACE(1, t, 5);
    }
    // ---- epilogue: dV = dropout-scale * acc, dK = softmax-scale * acc
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    if (col0 < DK) {
      uint32_t rv[16], rk[16];
      tmem_ld16(t_lane + T_DV + col0, rv);
      tmem_ld16(t_lane + T_DK + col0, rk);
      tmem_ld_wait();
      float vv[16], vk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        vv[i] = key_ok ? __uint_as_float(rv[i]) * dscale : 0.f;
        vk[i] = key_ok ? __uint_as_float(rk[i]) * p.scale : 0.f;
      }
      if (key_ok) {
        float* dvp = static_cast<float*>(p.dv) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddv + h * DK + col0;
        float* dkp = static_cast<float*>(p.dk) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddk + h * DK + col0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          *reinterpret_cast<float4*>(dvp + i) = make_float4(tf32_rna(vv[i]), tf32_rna(vv[i + 1]), tf32_rna(vv[i + 2]), tf32_rna(vv[i + 3]));
          *reinterpret_cast<float4*>(dkp + i) = make_float4(tf32_rna(vk[i]), tf32_rna(vk[i + 1]), tf32_rna(vk[i + 2]), tf32_rna(vk[i + 3]));
        }
      }
      if (p.dbv) bias_colsum16(vv, p.dbv + h * DK + col0, lane);
      if (p.dbk) bias_colsum16(vk, p.dbk + h * DK + col0, lane);
    }
  }
  if (p.trace && tid == 0 && cta_lin < TRACE_CTAS) g_cta_trace[cta_lin * 4 + 1] = global_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}


// ================================================================================ dK, dV for ONE query tile (Lq <= 64)
// Decoder cross-attention (50 target positions against 1000 encoder frames) and decoder self-attention.  With a single
// query tile the generic kernel above is one short tile per CTA — barrier / TMEM set-up, operand latency and the output
// write are all exposed (2048 CTAs x ~9 us at B=32, h=8, Lk=1000).  Here a CTA keeps the query-side tiles (Q, dO in both
// layouts, log-sum-exp, delta, dropout keys) resident and walks over SEVERAL key tiles: K/V tiles double-buffered in shared
// memory (TMA), S^T / dP^T and the dV / dK accumulators double-buffered in TMEM, so the loads of tile i+1, the MMAs of
// tile i and the output write of tile i-1 overlap.  Same mathematics and thread mapping as attn_bwd_dkv_pipe.
template <int DK>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dkv_small(const __grid_constant__ CUtensorMap tmap_q_k, const __grid_constant__ CUtensorMap tmap_q_mn,
                   const __grid_constant__ CUtensorMap tmap_do_k, const __grid_constant__ CUtensorMap tmap_do_mn,
                   const __grid_constant__ CUtensorMap tmap_k_res, const __grid_constant__ CUtensorMap tmap_v_res,
                   const AttnDev p) {
  constexpr int BKV = 128;
  constexpr int G = DK / 32;
  constexpr int T_BYTES = BT * DK * 4;
  constexpr int RES_BYTES = BKV * DK * 4;
  constexpr uint32_t TCOLS = 512;
  // TMEM: S^T x2 | dP^T x2 | dV x2 | dK x2
  constexpr uint32_t T_ST = 0, T_DPT = 2 * BT, T_DV = 4 * BT, T_DK = 4 * BT + 2 * DK;
  static_assert(4 * BT + 4 * DK <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sBase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRes = sBase;                                  // [2 buffers][K | V]
  uint8_t* sQ = sBase + 4 * RES_BYTES;                    // Qk | Qm | dOk | dOm
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
        s_delta[e] = e < p.Lq ? p.delta[o] : 0.f;
        s_rkey[e] = p.drop_thresh ? dropout_row_key(p.drop_seed, static_cast<uint64_t>(o)) : 0u;
      }
      __syncwarp();
      if (elect_one()) {
        tma_prefetch_desc(&tmap_q_k); tma_prefetch_desc(&tmap_k_res); tma_prefetch_desc(&tmap_v_res);
        mbar_arrive_expect_tx(&q_full, 4 * T_BYTES);   // release: the statistics above become visible with it
        tma_load_4d(sQ, &tmap_q_k, &q_full, 0, 0, h * G, b);
        tma_load_4d(sQ + T_BYTES, &tmap_q_mn, &q_full, 0, 0, h * G, b);
        tma_load_4d(sQ + 2 * T_BYTES, &tmap_do_k, &q_full, 0, 0, h * G, b);
        tma_load_4d(sQ + 3 * T_BYTES, &tmap_do_mn, &q_full, 0, 0, h * G, b);
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
      const uint64_t dqk = umma_desc_kmajor(sq0), dqm = umma_desc_mnmajor(sq0 + T_BYTES, BT * 128);
      const uint64_t ddok = umma_desc_kmajor(sq0 + 2 * T_BYTES), ddom = umma_desc_mnmajor(sq0 + 3 * T_BYTES, BT * 128);
      auto issue_a = [&](int it) {   // S^T = K Q^T, dP^T = V dO^T   (A = K / V tile in shared memory, B = Q / dO K-major)
        const int rb = it & 1;
        mbar_wait(&res_full[rb], (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, BT, false, false);
          const uint64_t ak0 = umma_desc_kmajor(sr0 + rb * 2 * RES_BYTES), av0 = umma_desc_kmajor(sr0 + rb * 2 * RES_BYTES + RES_BYTES);
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks)
            umma_tf32_ss(tmem + T_ST + rb * BT, ak0 + koff(ks, BKV), dqk + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < DK / 8; ++ks)
            umma_tf32_ss(tmem + T_DPT + rb * BT, av0 + koff(ks, BKV), ddok + koff(ks, BT), idesc, ks > 0 ? 1u : 0u);
          umma_commit(&s_full[rb]);
          umma_commit(&res_empty[rb]);   // only these two products read the K / V tile: its buffer is free as soon as they retire,
        }                                // so the producer runs two key tiles ahead and the TMA latency stays hidden
        __syncwarp();
      };
      auto issue_b = [&](int it) {   // dV = P^T dO, dK = dS^T Q   (A in TMEM, B MN-major)
        const int rb = it & 1;
        mbar_wait(&ds_full[rb], (it >> 1) & 1);
        if (it >= 2) mbar_wait(&acc_empty[rb], ((it >> 1) - 1) & 1);   // the epilogue of tile it-2 has drained this accumulator pair
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, DK, false, true);
#pragma unroll
          for (int ks = 0; ks < BT / 8; ++ks)
            umma_tf32_ts(tmem + T_DV + rb * DK, tmem + T_ST + rb * BT + ks * 8, ddom + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
                         ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < BT / 8; ++ks)
            umma_tf32_ts(tmem + T_DK + rb * DK, tmem + T_DPT + rb * BT + ks * 8, dqm + static_cast<uint64_t>((ks * 1024) >> 4), idesc,
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
    auto epilogue = [&](int it) {   // dV = dropout-scale * acc, dK = softmax-scale * acc for the key tile of iteration `it`
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
          float* dvp = static_cast<float*>(p.dv) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddv + h * DK + col0;
          float* dkp = static_cast<float*>(p.dk) + (static_cast<int64_t>(b) * p.Lk + key) * p.lddk + h * DK + col0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
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
      uint32_t rs[16], rd[16];
      tmem_ld16(t_lane + T_ST + rb * BT + col0, rs);
      tmem_ld16(t_lane + T_DPT + rb * BT + col0, rd);
      tmem_ld_wait();
      if (!mask_dense && key_masked) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { rs[i] = 0u; rd[i] = 0u; }
      } else {
        const float* ls = s_lse + col0;
        const float* de = s_delta + col0;
        const uint32_t* rk = s_rkey + col0;
        if (mask_dense) {
          if (p.drop_thresh) dkv16_t<true, true>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
          else dkv16_t<true, false>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
        } else {
          if (p.drop_thresh) dkv16_t<false, true>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
          else dkv16_t<false, false>(rs, rd, ls, de, rk, p.scale_log2, dscale, p.drop_thresh, my_ckey, mrow, p.ms_q, col0, p.Lq, !key_masked, causal_key);
        }
      }
      tmem_st16(t_lane + T_ST + rb * BT + col0, rs);
      tmem_st16(t_lane + T_DPT + rb * BT + col0, rd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[rb]);
      if (it > 0) epilogue(it - 1);       // overlaps the gradient MMAs of this tile
    }
    if (n_it > 0) epilogue(n_it - 1);
    // key tiles entirely inside the utterance's padding: zero gradients
    for (int kt = n_kt + first; kt < n_kt_all; kt += step) {
      const int rows = min(BKV, p.Lk - kt * BKV);
      for (int i = tid; i < rows * (DK / 4); i += NCOMP) {
        const int r = i / (DK / 4), c = (i - r * (DK / 4)) * 4;
        const int64_t grow = static_cast<int64_t>(b) * p.Lk + kt * BKV + r;
        *reinterpret_cast<float4*>(static_cast<float*>(p.dk) + grow * p.lddk + h * DK + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(static_cast<float*>(p.dv) + grow * p.lddv + h * DK + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

template <int DK>
int launch_pipelined(cudaStream_t s, const AttnBwdArgs& a, const AttnDev& p) {
  const AttnArgs& f = a.f;
  const int cols = f.H * DK;
  {
    CUtensorMap tqk, tqm, tdk, tdm;
    ST_TRY(make_act_tmap(&tqk, f.q, f.ldq, cols, f.Lq, f.B, BT, 0, DK));
    ST_TRY(make_act_tmap(&tqm, f.q, f.ldq, cols, f.Lq, f.B, BT, 1, DK));
    ST_TRY(make_act_tmap(&tdk, a.dctx, a.lddctx, cols, f.Lq, f.B, BT, 0, DK));
    ST_TRY(make_act_tmap(&tdm, a.dctx, a.lddctx, cols, f.Lq, f.B, BT, 1, DK));
    CUtensorMap tkr, tvr;
    ST_TRY(make_act_tmap(&tkr, f.k, f.ldk, cols, f.Lk, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tvr, f.v, f.ldv, cols, f.Lk, f.B, 128, 0, DK));
    // Few query tiles (decoder cross-/self-attention, Lq <= 128): the CTA is prologue-bound, and fetching the resident K/V
    // tiles with one TMA box each beats the LDG -> tcgen05.st path (183 -> 134 us at B=32, h=8, Lq=50, Lk=1000); with many
    // query tiles the third ring stage matters more.  Option: 0 = this heuristic, 1 = always shared memory, 2 = always TMEM.
    const int rs_opt = get_option("attn_dkv_res_smem");
    const bool rs = rs_opt == 1 || (rs_opt == 0 && f.Lq <= 2 * BT);
    if (f.Lq <= BT && !get_option("attn_dkv_no_small")) {   // single query tile: the persistent multi-key-tile kernel
      constexpr int SMEM_SMALL = 4 * 128 * DK * 4 + 4 * BT * DK * 4 + 1024;
      auto ks = attn_bwd_dkv_small<DK>;
      static bool attr_small = false;
      if (!attr_small) { ST_CHECK_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL)); attr_small = true; }
      const int n_kt = (f.Lk + 127) / 128;
      int split = get_option("attn_dkv_small_split");
      if (split <= 0) split = 1;
      dim3 grid(n_kt < split ? n_kt : split, f.H, f.B);
      ProfScope prof(s, PROF_ATTN_DKV, 4.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
      ST_CHECK_CUDA(launch_pdl(ks, grid, dim3(NTHREADS), SMEM_SMALL, s, tqk, tqm, tdk, tdm, tkr, tvr, p));
      ST_CHECK_LAUNCH();
    } else {
    constexpr int SMEM_TS = 3 * 4 * BT * DK * 4 + 1024;
    constexpr int SMEM_RS = (DK == 64 ? 2 : 3) * 4 * BT * DK * 4 + 2 * 128 * DK * 4 + 1024;
    const int SMEM = rs ? SMEM_RS : SMEM_TS;
    auto kern = rs ? attn_bwd_dkv_pipe<DK, true> : attn_bwd_dkv_pipe<DK, false>;
    static bool attr[2] = {false, false};
    if (!attr[rs]) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr[rs] = true; }
    dim3 grid((f.Lk + 127) / 128, f.H, f.B);
    // algorithmic share of the attention backward carried by this kernel: dV and dK (S, dP recompute not counted)
    ProfScope prof(s, PROF_ATTN_DKV, 4.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), SMEM, s, static_cast<const float*>(f.k), f.ldk, static_cast<const float*>(f.v), f.ldv, tqk, tqm, tdk, tdm, tkr, tvr, p));
    ST_CHECK_LAUNCH();
    }
  }
  {
    CUtensorMap tkk, tkm, tvk;
    ST_TRY(make_act_tmap(&tkk, f.k, f.ldk, cols, f.Lk, f.B, BT, 0, DK));
    ST_TRY(make_act_tmap(&tkm, f.k, f.ldk, cols, f.Lk, f.B, BT, 1, DK));
    ST_TRY(make_act_tmap(&tvk, f.v, f.ldv, cols, f.Lk, f.B, BT, 0, DK));
    CUtensorMap tqr, tdr;
    ST_TRY(make_act_tmap(&tqr, f.q, f.ldq, cols, f.Lq, f.B, 128, 0, DK));
    ST_TRY(make_act_tmap(&tdr, a.dctx, a.lddctx, cols, f.Lq, f.B, 128, 0, DK));
    const bool rs = get_option("attn_dq_res_smem") != 0;
    constexpr int SMEM_TS = 4 * 3 * BT * DK * 4 + 1024;
    constexpr int SMEM_RS = 3 * 3 * BT * DK * 4 + 2 * 128 * DK * 4 + 1024;
    const int SMEM = rs ? SMEM_RS : SMEM_TS;
    auto kern = rs ? attn_bwd_dq_pipe<DK, true> : attn_bwd_dq_pipe<DK, false>;
    static bool attr[2] = {false, false};
    if (!attr[rs]) { ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr[rs] = true; }
    dim3 grid((f.Lq + 127) / 128, f.H, f.B);
    // algorithmic share: dQ plus the (single) S and dP products of the textbook backward
    ProfScope prof(s, PROF_ATTN_DQ, 6.0 * f.B * f.H * static_cast<double>(f.Lq) * f.Lk * DK);
    ST_CHECK_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), SMEM, s, static_cast<const float*>(f.q), f.ldq, static_cast<const float*>(a.dctx), a.lddctx, tkk, tkm, tvk, tqr, tdr, p));
    ST_CHECK_LAUNCH();
  }
  return ST_OK;
}

}  // namespace

int attn_read_trace(unsigned long long* host_out, int n) {
  const int total = 2 * TRACE_TILES * TRACE_EVENTS;
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  if (n > total) {   // the per-CTA records follow the tile timeline
    const int extra = n - total < TRACE_CTAS * 4 ? n - total : TRACE_CTAS * 4;
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(host_out + total, g_cta_trace, sizeof(unsigned long long) * extra));
  }
  ST_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * (n < total ? n : total)));
  return total;
}

int attn_bwd_pipelined(cudaStream_t s, const AttnBwdArgs& a, const AttnDev& p) {
  return a.f.dk == 32 ? launch_pipelined<32>(s, a, p) : launch_pipelined<64>(s, a, p);
}

}  // namespace st
