// st_selftest.cu — device self-tests for the tcgen05 building blocks, callable over the C ABI
// (st_selftest).  Each test builds inputs that are exactly representable in TF32, runs the device
// path and compares with a double-precision host computation, so a wrong descriptor / swizzle /
// TMEM lane mapping shows up as an O(1) error rather than hiding inside TF32 noise.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "st_common.cuh"
#include "st_gemm.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

float host_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & ~0x1FFFu;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

struct Lcg {
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ull + 3037000493ull) {}
  float next() {  // uniform in [-1, 1)
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return static_cast<float>(static_cast<int32_t>(s >> 32)) * (1.0f / 2147483648.0f);
  }
};

struct DevBuf {
  float* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n) { return cudaMalloc(&p, n * sizeof(float)) == cudaSuccess ? 0 : -1; }
};

// One GEMM case.  Returns max |dev - ref| / max|ref| through *err.
int gemm_case(GemmMode mode, int M, int N, int K, int splits, bool bias, bool relu, int aux_mode, bool round_out,
              double* err, int ldc_extra = 0, int max_ctas = 0) {
  // logical A[M,K], B^T... build in the storage layout each mode expects
  const int lda = (mode == GEMM_TN) ? ((M + 3) & ~3) + 4 : ((K + 3) & ~3) + 4;
  const int ldb = (mode == GEMM_NT) ? ((K + 3) & ~3) + 4 : ((N + 3) & ~3) + 8;
  const int ldc = N + ldc_extra;  // ldc_extra = 3 exercises the unaligned scalar store path
  const int a_rows = (mode == GEMM_TN) ? K : M;
  const int b_rows = (mode == GEMM_NT) ? N : K;
  std::vector<float> hA(static_cast<size_t>(a_rows) * lda), hB(static_cast<size_t>(b_rows) * ldb);
  std::vector<float> hbias(N), haux(static_cast<size_t>(M) * N), hC(static_cast<size_t>(M) * ldc, 0.f);
  Lcg rng(1234 + M * 7 + N * 13 + K * 17 + static_cast<int>(mode));
  for (auto& x : hA) x = host_tf32(rng.next());
  for (auto& x : hB) x = host_tf32(rng.next());
  for (auto& x : hbias) x = rng.next();
  for (auto& x : haux) x = rng.next();
  auto Aat = [&](int m, int k) { return mode == GEMM_TN ? hA[static_cast<size_t>(k) * lda + m] : hA[static_cast<size_t>(m) * lda + k]; };
  auto Bat = [&](int n, int k) { return mode == GEMM_NT ? hB[static_cast<size_t>(n) * ldb + k] : hB[static_cast<size_t>(k) * ldb + n]; };

  DevBuf dA, dB, dC, dbias, daux;
  if (dA.alloc(hA.size()) || dB.alloc(hB.size()) || dC.alloc(hC.size()) || dbias.alloc(N) || daux.alloc(haux.size())) {
    set_error("selftest: cudaMalloc failed");
    return ST_ERR_CUDA;
  }
  ST_CHECK_CUDA(cudaMemcpy(dA.p, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dB.p, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dbias.p, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(daux.p, haux.data(), haux.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemset(dC.p, 0, hC.size() * 4));

  GemmEpilogue ep;
  ep.bias = bias ? dbias.p : nullptr;
  ep.relu = relu ? 1 : 0;
  ep.aux = aux_mode ? daux.p : nullptr;
  ep.ldaux = N;
  ep.aux_mode = aux_mode;
  ep.round_tf32 = round_out ? 1 : 0;
  ep.atomic = splits > 1 ? 1 : 0;
  set_option("gemm_max_ctas", max_ctas);
  const int st = gemm_tf32(0, mode, dA.p, lda, dB.p, ldb, dC.p, ldc, M, N, K, ep, splits);
  set_option("gemm_max_ctas", 0);
  ST_TRY(st);
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  ST_CHECK_CUDA(cudaMemcpy(hC.data(), dC.p, hC.size() * 4, cudaMemcpyDeviceToHost));

  double max_ref = 0, max_err = 0;
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += static_cast<double>(Aat(m, k)) * static_cast<double>(Bat(n, k));
      if (bias) acc += hbias[n];
      if (relu) acc = acc > 0 ? acc : 0;
      if (aux_mode == 1) acc += haux[static_cast<size_t>(m) * N + n];
      if (aux_mode == 2) acc = haux[static_cast<size_t>(m) * N + n] > 0 ? acc : 0;
      if (round_out) acc = host_tf32(static_cast<float>(acc));
      const double got = hC[static_cast<size_t>(m) * ldc + n];
      max_ref = fmax(max_ref, fabs(acc));
      const double e = fabs(got - acc);
      max_err = fmax(max_err, isnan(got) ? 1e30 : e);
    }
  }
  *err = max_err / (max_ref > 0 ? max_ref : 1);
  if (round_out && *err < 6e-4) *err *= 0.1;  // one TF32 ulp of slack: fp32 vs double accumulation may round differently
  return ST_OK;
}

// ---- tcgen05.mma with the A operand in TMEM (.ts form) --------------------------------------
// One CTA, 128 threads.  A[128, K] is written to TMEM by its owning lanes (tcgen05.st), B comes
// from a TMA-loaded smem tile that is either K-major ([N, K] row-major in HBM) or MN-major
// ([K, N] row-major in HBM).  D[128, N] is read back with tcgen05.ld.
template <int N, int K, bool B_MN>
__global__ void __launch_bounds__(128)
ts_mma_test_kernel(const __grid_constant__ CUtensorMap tmap_b, const float* __restrict__ A, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  constexpr int TCOLS = 256;  // A at [0, K), D at [128, 128+N)
  static_assert(K <= 128 && N <= 128, "test tile");
  if (threadIdx.x == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_b, N * K * 4);
    if (!B_MN) {
      for (int g = 0; g < K / 32; ++g) tma_load_2d(smem + g * (N * 128), &tmap_b, &bar_b, g * 32, 0);  // box {32 k, N}
    } else {
      for (int g = 0; g < N / 32; ++g) tma_load_2d(smem + g * (K * 128), &tmap_b, &bar_b, g * 32, 0);  // box {32 n, K}
    }
  }
  // A row -> TMEM lane
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < K / 32; ++c) {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(A[threadIdx.x * K + c * 32 + i]);
    tmem_st32(lane_addr + c * 32, r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_wait(&bar_b, 0);
    constexpr uint32_t idesc = umma_idesc_tf32(128, N, false, B_MN);
    const uint32_t sb = smem_u32(smem);
    for (int k = 0; k < K / 8; ++k) {
      const uint64_t bdesc = B_MN ? umma_desc_mnmajor(sb + k * 1024, K * 128)
                                  : umma_desc_kmajor(sb + (k / 4) * (N * 128) + (k % 4) * 32);
      umma_tf32_ts(tmem + 128, tmem + k * 8, bdesc, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_addr + 128 + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[threadIdx.x * N + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

template <int N, int K, bool B_MN>
int ts_case(double* err) {
  std::vector<float> hA(128 * K), hB(static_cast<size_t>(N) * K), hD(128 * N);
  Lcg rng(99 + N + K + (B_MN ? 1000 : 0));
  for (auto& x : hA) x = host_tf32(rng.next());
  for (auto& x : hB) x = host_tf32(rng.next());  // K-major: [N][K];  MN-major: [K][N]
  DevBuf dA, dB, dD;
  if (dA.alloc(hA.size()) || dB.alloc(hB.size()) || dD.alloc(hD.size())) { set_error("selftest: cudaMalloc failed"); return ST_ERR_CUDA; }
  ST_CHECK_CUDA(cudaMemcpy(dA.p, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dB.p, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tb;
  uint64_t dims[2], strides[1];
  uint32_t box[2];
  if (!B_MN) { dims[0] = K; dims[1] = N; strides[0] = K * 4; box[0] = 32; box[1] = N; }
  else       { dims[0] = N; dims[1] = K; strides[0] = N * 4; box[0] = 32; box[1] = K; }
  ST_TRY(make_tmap_f32(&tb, dB.p, 2, dims, strides, box, B_MN ? 1 : 0));
  auto kern = ts_mma_test_kernel<N, K, B_MN>;
  const int smem = N * K * 4 + 1024;
  ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<1, 128, smem>>>(tb, dA.p, dD.p);
  ST_CHECK_CUDA(cudaGetLastError());
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  ST_CHECK_CUDA(cudaMemcpy(hD.data(), dD.p, hD.size() * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k)
        acc += static_cast<double>(hA[m * K + k]) * static_cast<double>(B_MN ? hB[static_cast<size_t>(k) * N + n] : hB[static_cast<size_t>(n) * K + k]);
      const double got = hD[m * N + n];
      max_ref = fmax(max_ref, fabs(acc));
      max_err = fmax(max_err, isnan(got) ? 1e30 : fabs(got - acc));
    }
  *err = max_err / (max_ref > 0 ? max_ref : 1);
  return ST_OK;
}


// ---- 16-bit operands (kind::f16) -----------------------------------------------------------------------------
template <typename T> T host_cvt(float x);
template <> __half host_cvt<__half>(float x) { return __float2half_rn(x); }
template <> __nv_bfloat16 host_cvt<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
float host_back(__half x) { return __half2float(x); }
float host_back(__nv_bfloat16 x) { return __bfloat162float(x); }

struct DevBytes {
  void* p = nullptr;
  ~DevBytes() { if (p) cudaFree(p); }
  int alloc(size_t n) { return cudaMalloc(&p, n) == cudaSuccess ? 0 : -1; }
};

// One 16-bit GEMM case: operands and aux of type T, output fp32 or T (c_lp).  Error as in gemm_case; a 16-bit output may
// differ from the double-precision reference rounded the same way by one unit in the last place.
template <typename T>
int gemm_case16(GemmMode mode, int M, int N, int K, int splits, bool bias, bool relu, int aux_mode, bool c_lp, double* err,
                int max_ctas = 0, int pad = 8) {
  constexpr int DT = Elem<T>::FMT == 0 ? ST_DTYPE_F16 : ST_DTYPE_BF16;
  const int lda = ((mode == GEMM_TN ? M : K) + 7) / 8 * 8 + pad;
  const int ldb = ((mode == GEMM_NT ? K : N) + 7) / 8 * 8 + 2 * pad;
  const int ldc = (N + 3) / 4 * 4;
  const int a_rows = (mode == GEMM_TN) ? K : M;
  const int b_rows = (mode == GEMM_NT) ? N : K;
  std::vector<T> hA(static_cast<size_t>(a_rows) * lda), hB(static_cast<size_t>(b_rows) * ldb), haux(static_cast<size_t>(M) * ldc);
  std::vector<float> hbias(N);
  Lcg rng(4321 + M * 7 + N * 13 + K * 17 + static_cast<int>(mode) + DT * 101);
  for (auto& x : hA) x = host_cvt<T>(rng.next());
  for (auto& x : hB) x = host_cvt<T>(rng.next());
  for (auto& x : hbias) x = rng.next();
  for (auto& x : haux) x = host_cvt<T>(rng.next());
  auto Aat = [&](int m, int k) { return host_back(mode == GEMM_TN ? hA[static_cast<size_t>(k) * lda + m] : hA[static_cast<size_t>(m) * lda + k]); };
  auto Bat = [&](int n, int k) { return host_back(mode == GEMM_NT ? hB[static_cast<size_t>(n) * ldb + k] : hB[static_cast<size_t>(k) * ldb + n]); };
  const size_t c_bytes = static_cast<size_t>(M) * ldc * (c_lp ? sizeof(T) : 4);
  DevBytes dA, dB, dC, dbias, daux;
  if (dA.alloc(hA.size() * sizeof(T)) || dB.alloc(hB.size() * sizeof(T)) || dC.alloc(c_bytes) || dbias.alloc(N * 4) ||
      daux.alloc(haux.size() * sizeof(T))) {
    set_error("selftest: cudaMalloc failed");
    return ST_ERR_CUDA;
  }
  ST_CHECK_CUDA(cudaMemcpy(dA.p, hA.data(), hA.size() * sizeof(T), cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dB.p, hB.data(), hB.size() * sizeof(T), cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dbias.p, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(daux.p, haux.data(), haux.size() * sizeof(T), cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemset(dC.p, 0, c_bytes));
  GemmEpilogue ep;
  ep.bias = bias ? static_cast<const float*>(dbias.p) : nullptr;
  ep.relu = relu ? 1 : 0;
  ep.aux = aux_mode ? daux.p : nullptr;
  ep.ldaux = ldc;
  ep.aux_mode = aux_mode;
  ep.atomic = splits > 1 ? 1 : 0;
  set_option("gemm_max_ctas", max_ctas);
  const int st = gemm_any(0, DT, mode, dA.p, lda, dB.p, ldb, dC.p, ldc, c_lp ? 1 : 0, M, N, K, ep, splits);
  set_option("gemm_max_ctas", 0);
  ST_TRY(st);
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  std::vector<uint8_t> hC(c_bytes);
  ST_CHECK_CUDA(cudaMemcpy(hC.data(), dC.p, c_bytes, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0;
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += static_cast<double>(Aat(m, k)) * static_cast<double>(Bat(n, k));
      if (bias) acc += hbias[n];
      if (relu) acc = acc > 0 ? acc : 0;
      const float ax = host_back(haux[static_cast<size_t>(m) * ldc + n]);
      if (aux_mode == 1) acc += ax;
      if (aux_mode == 2) acc = ax > 0 ? acc : 0;
      double got;
      if (c_lp) {
        acc = host_back(host_cvt<T>(static_cast<float>(acc)));
        got = host_back(reinterpret_cast<const T*>(hC.data())[static_cast<size_t>(m) * ldc + n]);
      } else {
        got = reinterpret_cast<const float*>(hC.data())[static_cast<size_t>(m) * ldc + n];
      }
      max_ref = fmax(max_ref, fabs(acc));
      max_err = fmax(max_err, isnan(got) ? 1e30 : fabs(got - acc));
    }
  }
  *err = max_err / (max_ref > 0 ? max_ref : 1);
  const double ulp = Elem<T>::FMT == 0 ? 1.0 / 1024 : 1.0 / 128;
  if (c_lp && *err <= 1.01 * ulp) *err *= 0.005;   // one unit in the last place of slack (fp32 vs double accumulation)
  return ST_OK;
}

// tcgen05.mma kind::f16 with the A operand in TMEM: A[128, K] is written as K/2 32-bit columns per lane, two consecutive
// K elements per column (low half first); B from a TMA-loaded smem tile, K-major ([N, K] in HBM) or MN-major ([K, N]).
// N = 64 (one 128-byte swizzle row in either layout).
template <typename T, int K, bool B_MN>
__global__ void __launch_bounds__(128)
ts16_test_kernel(const __grid_constant__ CUtensorMap tmap_b, const T* __restrict__ A, float* __restrict__ D) {
  constexpr int N = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  constexpr int TCOLS = 256;  // A at [0, K/2), D at [128, 128+N)
  static_assert(K <= 128 && K % 64 == 0, "test tile");
  if (threadIdx.x == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_b, N * K * 2);
    if (!B_MN) { for (int g = 0; g < K / 64; ++g) tma_load_2d(smem + g * (N * 128), &tmap_b, &bar_b, g * 64, 0); }   // box {64 k, N}
    else tma_load_2d(smem, &tmap_b, &bar_b, 0, 0);                                                                      // box {64 n, K}
  }
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < K / 64; ++c) {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i)
      r[i] = pack2<T>(to_f32(A[threadIdx.x * K + c * 64 + 2 * i]), to_f32(A[threadIdx.x * K + c * 64 + 2 * i + 1]));
    tmem_st32(lane_addr + c * 32, r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_wait(&bar_b, 0);
    constexpr uint32_t idesc = umma_idesc<T>(128, N, false, B_MN);
    const uint32_t sb = smem_u32(smem);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bdesc = B_MN ? umma_desc_mn<T>(sb + k * 2048, K * 128)
                                  : umma_desc_kmajor(sb + (k / 4) * (N * 128) + (k % 4) * 32);
      umma_f16_ts(tmem + 128, tmem + k * 8, bdesc, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_addr + 128 + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[threadIdx.x * N + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

template <typename T, int K, bool B_MN>
int ts16_case(double* err) {
  constexpr int N = 64;
  constexpr int DT = Elem<T>::FMT == 0 ? ST_DTYPE_F16 : ST_DTYPE_BF16;
  std::vector<T> hA(128 * K), hB(static_cast<size_t>(N) * K);
  std::vector<float> hD(128 * N);
  Lcg rng(199 + K + (B_MN ? 1000 : 0) + DT);
  for (auto& x : hA) x = host_cvt<T>(rng.next());
  for (auto& x : hB) x = host_cvt<T>(rng.next());  // K-major: [N][K];  MN-major: [K][N]
  DevBytes dA, dB, dD;
  if (dA.alloc(hA.size() * sizeof(T)) || dB.alloc(hB.size() * sizeof(T)) || dD.alloc(hD.size() * 4)) { set_error("selftest: cudaMalloc failed"); return ST_ERR_CUDA; }
  ST_CHECK_CUDA(cudaMemcpy(dA.p, hA.data(), hA.size() * sizeof(T), cudaMemcpyHostToDevice));
  ST_CHECK_CUDA(cudaMemcpy(dB.p, hB.data(), hB.size() * sizeof(T), cudaMemcpyHostToDevice));
  CUtensorMap tb;
  uint64_t dims[2], strides[1];
  uint32_t box[2];
  if (!B_MN) { dims[0] = K; dims[1] = N; strides[0] = K * 2; box[0] = 64; box[1] = N; }
  else       { dims[0] = N; dims[1] = K; strides[0] = N * 2; box[0] = 64; box[1] = K; }
  ST_TRY(make_tmap(&tb, DT, dB.p, 2, dims, strides, box, 0));
  auto kern = ts16_test_kernel<T, K, B_MN>;
  const int smem = N * K * 2 + 1024;
  ST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<1, 128, smem>>>(tb, static_cast<const T*>(dA.p), static_cast<float*>(dD.p));
  ST_CHECK_CUDA(cudaGetLastError());
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  ST_CHECK_CUDA(cudaMemcpy(hD.data(), dD.p, hD.size() * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k)
        acc += static_cast<double>(host_back(hA[m * K + k])) * static_cast<double>(host_back(B_MN ? hB[static_cast<size_t>(k) * N + n] : hB[static_cast<size_t>(n) * K + k]));
      const double got = hD[m * N + n];
      max_ref = fmax(max_ref, fabs(acc));
      max_err = fmax(max_err, isnan(got) ? 1e30 : fabs(got - acc));
    }
  *err = max_err / (max_ref > 0 ? max_ref : 1);
  return ST_OK;
}

// 16-bit cases, `which` relative to the first of them
int selftest16(int which, double* err) {
  using H = __half;
  using Bf = __nv_bfloat16;
  switch (which) {
    case 0:  return gemm_case16<H>(GEMM_NT, 128, 64, 64, 1, false, false, 0, false, err);      // one k-block, K-major both
    case 1:  return gemm_case16<Bf>(GEMM_NT, 128, 256, 256, 1, false, false, 0, false, err);   // BN = 256, 4 k-blocks
    case 2:  return gemm_case16<H>(GEMM_NT, 384, 512, 512, 1, true, true, 0, true, err, 2);    // bias + relu -> 16-bit out, 3 tiles per CTA
    case 3:  return gemm_case16<Bf>(GEMM_NT, 200, 136, 72, 1, true, false, 1, false, err);     // ragged M/N/K + 16-bit residual -> fp32
    case 4:  return gemm_case16<H>(GEMM_NN, 128, 64, 64, 1, false, false, 0, false, err);      // MN-major B
    case 5:  return gemm_case16<Bf>(GEMM_NN, 384, 512, 1536, 1, false, false, 2, true, err);   // dgrad shape + relu mask -> 16-bit
    case 6:  return gemm_case16<H>(GEMM_TN, 128, 64, 64, 1, false, false, 0, false, err);      // MN-major A and B
    case 7:  return gemm_case16<Bf>(GEMM_TN, 512, 512, 4096, 4, false, false, 0, false, err, 3); // wgrad, split-K atomics
    case 8:  return gemm_case16<H>(GEMM_TN, 136, 200, 1000, 3, false, false, 0, false, err);   // ragged wgrad
    case 9:  return gemm_case16<Bf>(GEMM_NT, 20000, 512, 64, 1, true, false, 0, true, err);    // >1 tile per CTA at full grid
    case 10: return gemm_case16<H>(GEMM_NN, 300, 4340, 512, 1, false, false, 0, false, err);   // wide N (vocabulary-like), fp32 out
    case 11: return gemm_case16<Bf>(GEMM_NN, 300, 512, 2048, 1, false, false, 1, true, err);   // dgrad + 16-bit residual -> 16-bit
    case 12: return ts16_case<H, 64, false>(err);     // A in TMEM (packed pairs), B K-major
    case 13: return ts16_case<Bf, 128, false>(err);
    case 14: return ts16_case<H, 64, true>(err);      // A in TMEM, B MN-major (same smem bytes as a K-major [K][64] tile)
    case 15: return ts16_case<Bf, 128, true>(err);
    case 16: case 17: case 18: case 19: {   // CTA-pair MMA (cta_group::2), all three operand modes
      set_option("gemm_cluster", 3);
      set_option("gemm_bn", 256);
      int st = ST_ERR_INVALID;
      if (which == 16) st = gemm_case16<H>(GEMM_NT, 256, 256, 64, 1, false, false, 0, false, err);
      if (which == 17) st = gemm_case16<Bf>(GEMM_NT, 384, 512, 512, 1, true, true, 0, true, err, 4);
      if (which == 18) st = gemm_case16<H>(GEMM_NN, 300, 768, 256, 1, false, false, 1, false, err);
      if (which == 19) st = gemm_case16<Bf>(GEMM_TN, 512, 512, 4096, 4, false, false, 0, false, err, 6);
      set_option("gemm_cluster", 0);
      set_option("gemm_bn", 0);
      return st;
    }
    default: set_error("selftest: no 16-bit case %d", which); return ST_ERR_INVALID;
  }
}
constexpr int kSelftests32 = 26, kSelftests16 = 20;

}  // namespace

// which: 0..N-1 selects a case; returns status, writes the relative error.
int selftest(int which, double* err) {
  *err = -1.0;
  if (which >= kSelftests32) return selftest16(which - kSelftests32, err);
  switch (which) {
    case 0:  return gemm_case(GEMM_NT, 128, 64, 32, 1, false, false, 0, false, err);    // single MMA k-block
    case 1:  return gemm_case(GEMM_NT, 128, 256, 128, 1, false, false, 0, false, err);  // BN=256, 4 k-blocks
    case 2:  return gemm_case(GEMM_NT, 384, 512, 512, 1, true, true, 0, true, err, 0, 2);  // 3 tiles per CTA + bias/relu/round
    case 3:  return gemm_case(GEMM_NT, 200, 136, 72, 1, true, false, 1, false, err, 3); // ragged M/N/K + residual, unaligned ldc
    case 4:  return gemm_case(GEMM_NN, 128, 64, 32, 1, false, false, 0, false, err);    // MN-major B, single block
    case 5:  return gemm_case(GEMM_NN, 384, 512, 1536, 1, false, false, 2, true, err);  // dgrad shape + relu mask
    case 6:  return gemm_case(GEMM_TN, 128, 64, 32, 1, false, false, 0, false, err);    // MN-major A and B
    case 7:  return gemm_case(GEMM_TN, 512, 512, 4096, 4, false, false, 0, false, err, 0, 3); // wgrad, split-K atomics, 3 CTAs
    case 8:  return gemm_case(GEMM_TN, 136, 200, 1000, 3, false, false, 0, false, err); // ragged wgrad
    case 9:  return gemm_case(GEMM_NT, 20000, 512, 64, 1, true, false, 0, false, err);  // >1 tile per CTA at full grid
    case 10: return ts_case<64, 32, false>(err);    // A in TMEM, B K-major
    case 11: return ts_case<128, 64, false>(err);
    case 12: return ts_case<64, 32, true>(err);     // A in TMEM, B MN-major
    case 13: return ts_case<64, 128, true>(err);
    case 14: return gemm_case(GEMM_NN, 300, 4340, 512, 1, false, false, 0, false, err); // wide N (vocab-like)
    case 15: case 16: case 17: case 18: case 19: {   // 2-CTA clusters with the multicast B tile, all three operand modes
      set_option("gemm_cluster", 2);
      int st = ST_ERR_INVALID;
      if (which == 15) st = gemm_case(GEMM_NT, 384, 512, 512, 1, true, true, 0, true, err, 0, 4);     // odd m_tiles, 2 clusters
      if (which == 16) st = gemm_case(GEMM_NN, 300, 768, 256, 1, false, false, 1, false, err);        // ragged M, residual
      if (which == 17) st = gemm_case(GEMM_TN, 512, 512, 4096, 4, false, false, 0, false, err, 0, 6); // split-K atomics
      if (which == 18) st = gemm_case(GEMM_NT, 20000, 512, 64, 1, true, false, 0, false, err);        // full grid, many units
      if (which == 19) st = gemm_case(GEMM_NT, 256, 300, 72, 1, false, false, 0, false, err, 3);      // ragged N/K tails
      set_option("gemm_cluster", 0);
      return st;
    }
    case 20: case 21: case 22: case 23: case 24: case 25: {   // CTA-pair MMA (cta_group::2), all three operand modes
      set_option("gemm_cluster", 3);
      set_option("gemm_bn", 256);
      int st = ST_ERR_INVALID;
      if (which == 20) st = gemm_case(GEMM_NT, 256, 256, 32, 1, false, false, 0, false, err);         // one pair, one k-block
      if (which == 21) st = gemm_case(GEMM_NT, 384, 512, 512, 1, true, true, 0, true, err, 0, 4);     // odd m_tiles, 2 pairs
      if (which == 22) st = gemm_case(GEMM_NN, 300, 768, 256, 1, false, false, 1, false, err);        // ragged M, residual
      if (which == 23) st = gemm_case(GEMM_TN, 512, 512, 4096, 4, false, false, 0, false, err, 0, 6); // split-K atomics
      if (which == 24) st = gemm_case(GEMM_NT, 20000, 512, 64, 1, true, false, 0, false, err);        // full grid, many units
      if (which == 25) st = gemm_case(GEMM_NT, 256, 300, 72, 1, false, false, 0, false, err, 3);      // ragged N/K tails
      set_option("gemm_cluster", 0);
      set_option("gemm_bn", 0);
      return st;
    }
    default: set_error("selftest: no case %d", which); return ST_ERR_INVALID;
  }
}

int selftest_count() { return kSelftests32 + kSelftests16; }

// ---- tcgen05.mma issue-rate microbenchmark (design input for the attention kernels; tools/mma_bench.py) --------
// One CTA issues `iters` groups of 16 TF32 MMAs (M = 128, N = n, K = 8 each, the k-offset pattern of the attention
// kernels) back to back and waits for the last commit; reports cycles per MMA.  Operand contents are zeros.
//   variant bit 0: A from TMEM (.ts) instead of shared memory;  bit 1: B MN-major (32-byte-atom swizzle) instead of K-major
namespace {
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int variant, int n, int iters, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32 + 64) * 1024 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  {   // zero the TMEM A region (columns 256..319)
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    tmem_st32(t_lane + 256, z);
    tmem_st32(t_lane + 288, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t sA = smem_u32(sm), sB = sA + 32 * 1024;
  const bool a_tmem = variant & 1, b_mn = variant & 2;
  const uint32_t idesc = umma_idesc_tf32(128, n, false, b_mn);
  const int nacc = ((variant >> 2) & 3) + 1;                       // accumulators cycled through (1, 2 or 4)
  if (!(variant & 16)) {
    // (a) the issuing THREAD: everything under `tid == 0`
    if (tid == 0) {
      uint32_t phase = 0;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ks = j & 7;
          const uint64_t bdesc = b_mn ? umma_desc_mnmajor(sB + ks * 1024, 64 * 128)
                                      : umma_desc_kmajor(sB + (ks / 4) * (n * 128) + (ks % 4) * 32);
          const uint32_t dcol = static_cast<uint32_t>((j % nacc) * n) & 255u;
          if (a_tmem) umma_tf32_ts(tmem + dcol, tmem + 256 + ks * 8, bdesc, idesc, 1u);
          else umma_tf32_ss(tmem + dcol, umma_desc_kmajor(sA + (ks / 4) * (128 * 128) + (ks % 4) * 32), bdesc, idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
      out[0] = static_cast<unsigned long long>(clock64() - t0);
    }
  } else if (warp == 0) {
    // (b) the issuing WARP stays converged; one elected lane issues (elect.sync), descriptors are base + constant
    uint32_t phase = 0;
    const uint64_t a0 = umma_desc_kmajor(sA);
    const uint64_t b0 = b_mn ? umma_desc_mnmajor(sB, 64 * 128) : umma_desc_kmajor(sB);
    const uint32_t bgrp = b_mn ? 0u : static_cast<uint32_t>(n * 128) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ks = j & 7;
          const uint64_t bdesc = b_mn ? b0 + static_cast<uint64_t>((ks * 1024) >> 4)
                                      : b0 + static_cast<uint64_t>((ks / 4) * bgrp + (((ks % 4) * 32) >> 4));
          const uint32_t dcol = static_cast<uint32_t>((j % nacc) * n) & 255u;
          if (a_tmem) umma_tf32_ts(tmem + dcol, tmem + 256 + ks * 8, bdesc, idesc, 1u);
          else umma_tf32_ss(tmem + dcol, a0 + static_cast<uint64_t>(((ks / 4) * (128 * 128) + (ks % 4) * 32) >> 4), bdesc, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (tid == 0) out[0] = static_cast<unsigned long long>(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
}  // namespace

int mma_bench(int variant, int n, int iters, double* clk_per_mma) {
  ST_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && variant >= 0 && variant < 32, "mma_bench: bad arguments");
  unsigned long long* d = nullptr;
  ST_CHECK_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
  const int smem = (32 + 64) * 1024 + 1024;
  ST_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_bench_kernel<<<1, 128, smem>>>(variant, n, iters, d);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  ST_CHECK_CUDA(e);
  *clk_per_mma = static_cast<double>(h) / (16.0 * iters);
  return ST_OK;
}

}  // namespace st
