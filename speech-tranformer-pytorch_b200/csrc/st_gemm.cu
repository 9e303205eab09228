// st_gemm.cu — persistent, warp-specialised TF32 GEMM on tcgen05 tensor cores.
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld -> smem transpose -> bias / ReLU / dropout / residual / TF32
//                 rounding -> coalesced global stores)
//
// Output tile 128 x BN (BN in {64,128,256}); K is consumed in blocks of 32 fp32 (= one 128-byte
// swizzle atom); accumulators are double-buffered in TMEM (2*BN columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.  Operands may be K-major or MN-major (see
// st_common.cuh; MN-major TF32 tiles use the 32-byte-atom swizzle), which covers forward (NT),
// data-gradient (NN) and weight-gradient (TN) GEMMs without any transposed copies in HBM.
//
// Epilogue: a TMEM lane is an output row, so tcgen05.ld hands each thread 32 consecutive columns
// of ITS row; storing that directly would touch 32 different 128-byte lines per instruction.  Each
// epilogue warp therefore transposes 32x32 blocks through a padded smem tile and stores 4 full
// 128-byte row segments per instruction (bias and residual loads are coalesced the same way).
#include "st_common.cuh"
#include "st_gemm.cuh"
#include "st_host.h"

namespace st {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                       // fp32 elements per k-block = 128 bytes
constexpr int A_STAGE_BYTES = BM * BK * 4;   // 16 KB
constexpr int GEMM_THREADS = 192;
constexpr int EPI_LD = 36;                   // padded row length (floats) of the per-warp transpose tile
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  float* C;
  int64_t ldc;
  int M, N, K;
  int m_tiles, n_tiles, k_splits;
  int kblocks_total, kblocks_per_split;
  GemmEpilogue ep;
};

// CL = CTAs per cluster (1 or 2).  With CL == 2 the two CTAs of a cluster own vertically adjacent output tiles
// (same n_blk), each TMA-loads half of the shared B tile and multicasts it to both, which cuts the L2->SM operand
// traffic per CTA from A+B to A+B/2; a smem stage is then released by BOTH consumers (multicast tcgen05.commit).
template <int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

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
      mbar_init(&tempty_bar[s], 128);
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

  // work units: CL vertically adjacent tiles; CTA `crank` of the cluster takes tile m = unit_m * CL + crank
  const int crank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_units = (p.m_tiles + CL - 1) / CL;
  const int tiles_mn = m_units * p.n_tiles;
  const int num_tiles = tiles_mn * p.k_splits;
  const int unit0 = blockIdx.x / CL, unit_stride = gridDim.x / CL;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
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
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 32; ++i)
              tma_load_2d(sa + i * 4096, &tmap_a, &full_bar[stage], m_blk * BM + i * 32, kb * BK);
          }
          if (CL == 1) {
            if (!B_MN) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 32; ++i)
                tma_load_2d(sb + i * 4096, &tmap_b, &full_bar[stage], n_blk * BN + i * 32, kb * BK);
            }
          } else {  // this CTA fetches its half of the B tile for the whole cluster
            constexpr uint16_t kAll = (1u << CL) - 1;
            if (!B_MN) {
              tma_load_2d_mc(sb + crank * (BN / CL) * 128, &tmap_b, &full_bar[stage], kb * BK,
                             n_blk * BN + crank * (BN / CL), kAll);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 32 / CL; ++i) {
                const int g = crank * (BN / 32 / CL) + i;
                tma_load_2d_mc(sb + g * 4096, &tmap_b, &full_bar[stage], n_blk * BN + g * 32, kb * BK, kAll);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, A_MN, B_MN);
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
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            // K-major: +32 B per MMA inside the 128 B swizzle row.  MN-major: next 8-row K atom pair (+1024 B);
            // 32-wide MN groups are 4096 B apart (one TMA box each).
            const uint64_t adesc = A_MN ? umma_desc_mnmajor(sa + k * 1024, 4096) : umma_desc_kmajor(sa + k * 32);
            const uint64_t bdesc = B_MN ? umma_desc_mnmajor(sb + k * 1024, 4096) : umma_desc_kmajor(sb + k * 32);
            umma_tf32_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage once these MMAs retire (in every CTA whose TMA writes into it)
          if (CL > 1) umma_commit_mc(&empty_bar[stage], (1u << CL) - 1); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const GemmEpilogue& ep = p.ep;
    float* stg = epi_smem + (warp - 2) * (32 * EPI_LD);
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0) &&
                        (!ep.aux || (((ep.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.aux) & 15) == 0)));
    const int lcol = (lane & 7) * 4;   // this lane's 4 columns inside a 32-column chunk
    const int lrow = lane >> 3;        // and its row inside each group of 4 rows
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int n_blk = tile % p.n_tiles;
      const int m_blk = ((tile / p.n_tiles) % m_units) * CL + crank;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row_base = m_blk * BM + quarter * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col = n_blk * BN + c * 32 + lcol;
        if (n_blk * BN + c * 32 >= p.N) break;  // warp-uniform: nothing left in this tile row-block
        // residual / ReLU-mask operand: issue all 8 row loads of this chunk now so their HBM latency overlaps the
        // TMEM read and the smem transpose instead of serialising inside the store loop
        float4 aux4[8];
        if (ep.aux_mode && vec_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row_base + i * 4 + lrow;
            aux4[i] = (row < p.M && col < p.N)
                          ? __ldg(reinterpret_cast<const float4*>(ep.aux + static_cast<int64_t>(row) * ep.ldaux + col))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        {
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * EPI_LD + j * 4) =
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
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + lrow;
          const int row = row_base + rr;
          if (row >= p.M || col >= p.N) continue;
          const float4 s4 = *reinterpret_cast<const float4*>(stg + rr * EPI_LD + lcol);
          float v[4] = {s4.x + b4[0], s4.y + b4[1], s4.z + b4[2], s4.w + b4[3]};
          if (ep.relu) {
#pragma unroll
            for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], 0.f);
          }
          if (ep.drop_thresh) {
            const uint32_t key = dropout_row_key(ep.drop_seed, static_cast<uint64_t>(row));
#pragma unroll
            for (int t = 0; t < 4; t += 2) {  // col is a multiple of 4: one hash per column pair
              const uint32_t bits = dropout_pair(key, col + t);
              v[t] = ((bits & 0xFFFFu) >= ep.drop_thresh) ? v[t] * ep.drop_scale : 0.f;
              v[t + 1] = ((bits >> 16) >= ep.drop_thresh) ? v[t + 1] * ep.drop_scale : 0.f;
            }
          }
          float* cp = p.C + static_cast<int64_t>(row) * p.ldc + col;
          const float* ap = ep.aux ? ep.aux + static_cast<int64_t>(row) * ep.ldaux + col : nullptr;
          if (vec_ok) {  // all 4 columns in range (N % 4 == 0)
            if (ep.aux_mode) {
              const float a[4] = {aux4[i].x, aux4[i].y, aux4[i].z, aux4[i].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) v[t] = (ep.aux_mode == 1) ? v[t] + a[t] : (a[t] > 0.f ? v[t] * ep.aux_scale : 0.f);
            }
            if (ep.round_tf32) {
#pragma unroll
              for (int t = 0; t < 4; ++t) v[t] = tf32_rna(v[t]);
            }
            if (ep.atomic) {
#pragma unroll
              for (int t = 0; t < 4; ++t) atomicAdd(cp + t, v[t]);
            } else {
              *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (col + t >= p.N) break;
              float x = v[t];
              if (ep.aux_mode == 1) x += ap[t];
              if (ep.aux_mode == 2) x = ap[t] > 0.f ? x * ep.aux_scale : 0.f;
              if (ep.round_tf32) x = tf32_rna(x);
              if (ep.atomic) atomicAdd(cp + t, x); else cp[t] = x;
            }
          }
        }
        __syncwarp();
      }
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

template <int BN, bool A_MN, bool B_MN, int CL>
int launch_gemm(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, CL>;
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
  ProfScope prof(stream, PROF_GEMM, 2.0 * p.M * static_cast<double>(p.N) * p.K);
  if (CL == 1) {
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  }
  ST_CHECK_LAUNCH();
  return ST_OK;
}

template <int BN, int CL>
int dispatch_mode(cudaStream_t stream, GemmMode mode, const CUtensorMap& ta, const CUtensorMap& tb,
                  const GemmParams& p) {
  switch (mode) {
    case GEMM_NT: return launch_gemm<BN, false, false, CL>(stream, ta, tb, p);
    case GEMM_NN: return launch_gemm<BN, false, true, CL>(stream, ta, tb, p);
    case GEMM_TN: return launch_gemm<BN, true, true, CL>(stream, ta, tb, p);
  }
  set_error("gemm_tf32: bad mode %d", static_cast<int>(mode));
  return ST_ERR_INVALID;
}

// Tile width: 256 when that still gives every SM a tile; otherwise narrower tiles spread a small problem
// (decoder-side GEMMs with M = B*L ~ 1600 rows) over more SMs.
int pick_bn(int M, int N, int k_splits) {
  if (N <= 64) return 64;
  const int m_tiles = (M + BM - 1) / BM;
  const int sms = num_sms();
  const int forced = get_option("gemm_bn");
  if (forced == 64 || forced == 128 || forced == 256) return (forced > 64 && N <= 64) ? 64 : forced;
  if (N > 128 && m_tiles * ((N + 255) / 256) * k_splits >= sms) return 256;
  if (N > 64 && m_tiles * ((N + 127) / 128) * k_splits >= sms) return 128;
  return 64;
}

}  // namespace

int gemm_tf32(cudaStream_t stream, GemmMode mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
              int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, int k_splits) {
  ST_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tf32: empty problem M=%d N=%d K=%d", M, N, K);
  ST_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             "gemm_tf32: operands must be 16-byte aligned");
  ST_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "gemm_tf32: lda=%lld ldb=%lld must be multiples of 4 floats",
             (long long)lda, (long long)ldb);
  ST_REQUIRE(k_splits >= 1 && (k_splits == 1 || ep.atomic), "gemm_tf32: split-K needs the atomic epilogue");

  GemmParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.kblocks_total = (K + BK - 1) / BK;
  if (k_splits > p.kblocks_total) k_splits = p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + k_splits - 1) / k_splits;
  p.k_splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  const int BN = pick_bn(M, N, p.k_splits);
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.ep = ep;
  // 2-CTA clusters with a multicast B tile
  // (measured on B200: correct but 10-20 % SLOWER than independent CTAs at these shapes — the two CTAs run in
  // lock-step and the bytes delivered to each SM do not change — so it is opt-in: st_set_option("gemm_cluster", 2))
  int CL = 1;
  const int forced_cl = get_option("gemm_cluster");
  if (forced_cl == 1 || (forced_cl == 2 && BN == 256 && p.m_tiles >= 2)) CL = forced_cl;

  CUtensorMap ta, tb;
  {
    // A: K-major -> dims {K, M}; MN-major (TN) -> stored [K, M], dims {M, K}
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(lda) * 4};
    uint32_t box[2];
    if (mode == GEMM_TN) { dims[0] = M; dims[1] = K; box[0] = 32; box[1] = 32; }
    else                 { dims[0] = K; dims[1] = M; box[0] = 32; box[1] = BM; }
    ST_TRY(make_tmap_f32(&ta, A, 2, dims, strides, box, (mode == GEMM_TN) ? 1 : 0));
  }
  {
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(ldb) * 4};
    uint32_t box[2];
    if (mode == GEMM_NT) { dims[0] = K; dims[1] = N; box[0] = 32; box[1] = static_cast<uint32_t>(BN / CL); }
    else                 { dims[0] = N; dims[1] = K; box[0] = 32; box[1] = 32; }
    ST_TRY(make_tmap_f32(&tb, B, 2, dims, strides, box, (mode != GEMM_NT) ? 1 : 0));
  }
  switch (BN) {
    case 256: return CL == 2 ? dispatch_mode<256, 2>(stream, mode, ta, tb, p) : dispatch_mode<256, 1>(stream, mode, ta, tb, p);
    case 128: return dispatch_mode<128, 1>(stream, mode, ta, tb, p);
    default:  return dispatch_mode<64, 1>(stream, mode, ta, tb, p);
  }
}

}  // namespace st
