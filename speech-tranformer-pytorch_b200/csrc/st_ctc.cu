// st_ctc.cu — CTC loss head for the joint CTC / attention objective (BASELINE.json configs[3]).
//
// The reference's train_attn_and_ctc.py is an empty file, so there is no reference code to follow; the algorithm
// is the published one (Graves et al. 2006: forward-backward over the blank-extended label sequence, log domain)
// with the conventions of torch.nn.functional.ctc_loss, which is the oracle of the parity tests:
//   extended labels l' = [blank, l_1, blank, l_2, ..., l_L, blank]   (S = 2L + 1 states)
//   alpha_t(s) = y_t(l'_s) * (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [l'_s != blank and l'_s != l'_{s-2}] alpha_{t-1}(s-2))
//   nll = -log(alpha_{T-1}(S-1) + alpha_{T-1}(S-2)),   y = softmax(logits)
//   d nll / d logit_t(v) = y_t(v) - (1 / (P * y_t(v))) * sum_{s : l'_s = v} alpha_t(s) beta_t(s)
// Three kernels, logits (B, T, V) fp32 read twice (555 MB at B=32, T=1000, V=4337 — HBM-bound), never a log-softmax
// tensor in memory:
//   ctc_lse_kernel        block per frame: log-sum-exp over the vocabulary                         (HBM)
//   ctc_alpha_beta_kernel two CTAs per utterance (alpha forward in time, beta backward), thread per state, the
//                         emission gathers run `PF` frames ahead of the recursion so the dependent chain per frame is
//                         one barrier + a handful of ALU ops                                        (latency)
//   ctc_grad_kernel       block per frame: softmax row, then the <= S label corrections             (HBM)
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

constexpr int CTC_THREADS = 256;
constexpr int PF = 8;   // emission prefetch distance of the recursion kernel (frames)

// log(exp(a) + exp(b)).  alpha / beta are sums of up to T log-probabilities (magnitude ~ 10 T): they are carried in
// double so that the cancellation in alpha + beta - log P keeps ~1e-7 relative accuracy (fp32 loses 1e-3 at T = 300),
// while the transcendental part only ever sees the small difference |a - b| and stays in fp32.
__device__ __forceinline__ double log_add(double a, double b) {
  const double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  const float d = static_cast<float>(fabs(a - b));
  return m + static_cast<double>(log1pf(__expf(-d)));
}
// log(exp(a) + exp(b) + [use_c] exp(c)) on the recursion's critical path: one warp per scheduler runs a dependent chain
// per frame, so instruction count is latency.  The sum of the shifted exponentials lies in [1, 3]: MUFU ex2 / lg2
// (absolute error ~1e-7) are exact enough because the result is ADDED to the running maximum.
__device__ __forceinline__ double log_add3(double a, double b, double c, bool use_c) {
  double m = fmax(a, b);
  if (use_c) m = fmax(m, c);
  if (m == -INFINITY) return -INFINITY;
  float sum = __expf(static_cast<float>(a - m)) + __expf(static_cast<float>(b - m));
  if (use_c) sum += __expf(static_cast<float>(c - m));
  return m + static_cast<double>(__logf(sum));
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_lse_kernel(const float* __restrict__ logits, int64_t ld, int64_t rows, int V, float* __restrict__ lse) {
  __shared__ float red[CTC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* x = logits + row * ld;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < V; c += CTC_THREADS) m = fmaxf(m, x[c]);
    m = warp_max(m);
    __syncthreads();
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < CTC_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    float s = 0.f;
    for (int c = threadIdx.x; c < V; c += CTC_THREADS) s += __expf(x[c] - m);
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < CTC_THREADS / 32; ++w) t += red[w];
      lse[row] = m + logf(t);
    }
  }
}

// grid (B, 2): y = 0 runs alpha (t ascending), y = 1 runs beta (t descending).  ab: (2, B, T, S_max) log-domain values.
__global__ void ctc_alpha_beta_kernel(const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
                                      const int64_t* __restrict__ targets, int64_t ld_tgt, const int64_t* __restrict__ in_len,
                                      const int64_t* __restrict__ tgt_len, int blank, int B, int T, int V, int S_max,
                                      double* __restrict__ ab, double* __restrict__ nll_d, float* __restrict__ nll) {
  extern __shared__ double sh[];          // two ping-pong rows of S_max + 4 (two -inf pads on either side)
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int s = threadIdx.x;
  int Tb = static_cast<int>(in_len[b]);
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int L = static_cast<int>(tgt_len[b]);
  L = L < 0 ? 0 : (2 * L + 1 > S_max ? (S_max - 1) / 2 : L);
  const int S = 2 * L + 1;
  double* out = ab + ((static_cast<int64_t>(is_beta ? 1 : 0) * B + b) * T) * S_max;
  const bool live = s < S;
  // this state's label, and whether the skip transition into it (alpha: from s-2; beta: from s+2) is allowed
  int lab = blank;
  bool skip = false;
  if (live && (s & 1)) {
    lab = static_cast<int>(targets[b * ld_tgt + (s >> 1)]);
    lab = lab < 0 ? 0 : (lab >= V ? V - 1 : lab);
    const int other = is_beta ? s + 2 : s - 2;
    if (other >= 0 && other < S) {
      int lo = static_cast<int>(targets[b * ld_tgt + (other >> 1)]);
      skip = lo != lab;
    }
  }
  double* row0 = sh + 2;
  double* row1 = sh + (S_max + 4) + 2;
  if (s < 2) { row0[-2 + s] = -INFINITY; row1[-2 + s] = -INFINITY; }            // pads for s-1, s-2 (alpha)
  if (s < 2) { row0[S_max + s] = -INFINITY; row1[S_max + s] = -INFINITY; }      // pads for s+1, s+2 (beta)
  if (Tb == 0) {
    if (!is_beta && s == 0) { nll[b] = (S == 1) ? 0.f : INFINITY; nll_d[b] = nll[b]; }
    return;
  }
  const float* xb = logits + static_cast<int64_t>(b) * T * ld;
  const float* lb = lse + static_cast<int64_t>(b) * T;
  // Emissions are gathered PF frames ahead of their use.  The raw loads (logit, log-sum-exp) sit in registers until then:
  // combining them at issue time would stall the warp on the load latency every frame (in-order issue).
  auto frame = [&](int i) { return is_beta ? Tb - 1 - i : i; };   // i-th frame in recursion order
  float px[PF], pl[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const bool ok = live && i < Tb;
    px[i] = ok ? xb[static_cast<int64_t>(frame(i)) * ld + lab] : 0.f;
    pl[i] = ok ? lb[frame(i)] : 0.f;
  }
  double* prev = row0;
  double* cur = row1;
  // unrolled by PF: prefetch slot k is a fixed register pair, consumed and immediately refilled (no register shifting,
  // which would read registers whose loads are still in flight)
  for (int i0 = 0; i0 < Tb; i0 += PF) {
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int i = i0 + k;
      if (i >= Tb) break;                 // uniform across the block
      const int t = frame(i);
      const float e = live ? px[k] - pl[k] : -INFINITY;
      {
        const bool ok = live && i + PF < Tb;
        px[k] = ok ? xb[static_cast<int64_t>(frame(i + PF)) * ld + lab] : 0.f;
        pl[k] = ok ? lb[frame(i + PF)] : 0.f;
      }
      double v;
      if (i == 0) {
        if (!is_beta) v = (s < 2 && live) ? e : -INFINITY;                    // alpha_0: states 0 and 1
        else v = (live && s >= S - 2) ? e : -INFINITY;                        // beta_{T-1}: states S-1 and S-2
      } else if (live) {      // threads beyond the last state never touch the rows (they would read past the pads)
        const double a0 = prev[s];
        const double a1 = is_beta ? prev[s + 1] : prev[s - 1];
        const double a2 = is_beta ? prev[s + 2] : prev[s - 2];    // a pad (-inf) or a real state; ignored unless `skip`
        v = log_add3(a0, a1, a2, skip) + e;
      } else {
        v = -INFINITY;
      }
      if (s < S_max) { cur[s] = v; out[static_cast<int64_t>(t) * S_max + s] = v; }
      __syncthreads();
      double* tmp = prev; prev = cur; cur = tmp;
    }
  }
  if (!is_beta && s == 0) {
    const double a_last = prev[S - 1];
    const double a_prev = S > 1 ? prev[S - 2] : static_cast<double>(-INFINITY);
    const double r = -log_add(a_last, a_prev);
    nll_d[b] = r;
    nll[b] = static_cast<float>(r);
  }
}

// grad[b,t,v] = scale_b * (softmax - label corrections) for t < in_len[b], 0 beyond.  scale: per-utterance factor
// (upstream gradient x reduction weight), may be nullptr (= 1).
__global__ void __launch_bounds__(CTC_THREADS)
ctc_grad_kernel(const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
                const int64_t* __restrict__ targets, int64_t ld_tgt, const int64_t* __restrict__ in_len,
                const int64_t* __restrict__ tgt_len, int blank, int B, int T, int V, int S_max,
                const double* __restrict__ ab, const double* __restrict__ nll_d, const float* __restrict__ scale,
                float* __restrict__ grad, int64_t ldg) {
  for (int64_t row = blockIdx.x; row < static_cast<int64_t>(B) * T; row += gridDim.x) {
    const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<int64_t>(b) * T);
    float* g = grad + row * ldg;
    const int Tb = static_cast<int>(in_len[b]);
    const double loss = nll_d[b];
    const float sc = scale ? scale[b] : 1.f;
    if (t >= Tb || !(loss < static_cast<double>(INFINITY))) {   // padded frame, or an infeasible alignment (torch zero_infinity-style zero gradient)
      for (int c = threadIdx.x; c < V; c += CTC_THREADS) g[c] = 0.f;
      __syncthreads();
      continue;
    }
    const float* x = logits + row * ld;
    const float l = lse[row];
    for (int c = threadIdx.x; c < V; c += CTC_THREADS) g[c] = sc * __expf(x[c] - l);
    __syncthreads();   // the corrections below touch columns written by other threads of this block
    int L = static_cast<int>(tgt_len[b]);
    L = 2 * L + 1 > S_max ? (S_max - 1) / 2 : L;
    const int S = 2 * L + 1;
    const double* al = ab + ((static_cast<int64_t>(b)) * T + t) * S_max;
    const double* be = ab + ((static_cast<int64_t>(B) + b) * T + t) * S_max;
    for (int s = threadIdx.x; s < S; s += CTC_THREADS) {
      int lab = blank;
      if (s & 1) {
        lab = static_cast<int>(targets[b * ld_tgt + (s >> 1)]);
        lab = lab < 0 ? 0 : (lab >= V ? V - 1 : lab);
      }
      const float lp = x[lab] - l;
      const float term = __expf(static_cast<float>(al[s] + be[s] - static_cast<double>(lp) + loss));     // alpha beta / (y P)
      atomicAdd(g + lab, -sc * term);
    }
    __syncthreads();
  }
}

}  // namespace

// lse (B*T floats) | alpha, beta (2*B*T*S_max doubles) | nll in double (B)
int64_t ctc_ws_floats(int B, int T, int S_max) {
  return ((static_cast<int64_t>(B) * T + 63) & ~int64_t(63)) + 2 * (2ll * B * T * S_max) + 2ll * B + 64;
}

int ctc_fwd_bwd(cudaStream_t stream, const float* logits, int64_t ld, const int64_t* targets, int64_t ld_tgt,
                const int64_t* in_len, const int64_t* tgt_len, int blank, int B, int T, int V, int L_max, float* nll,
                const float* scale, float* grad, int64_t ldg, float* ws, int64_t ws_floats, int grad_only) {
  ST_REQUIRE(B > 0 && T > 0 && V > 1 && L_max >= 0, "st_ctc: empty problem (B=%d T=%d V=%d L_max=%d)", B, T, V, L_max);
  ST_REQUIRE(blank >= 0 && blank < V, "st_ctc: blank index %d outside the vocabulary", blank);
  const int S_max = 2 * L_max + 1;
  ST_REQUIRE(S_max <= 1024, "st_ctc: target length %d exceeds the 511 labels one CTA can carry", L_max);
  ST_REQUIRE(ld >= V && (!grad || ldg >= V), "st_ctc: leading dimensions smaller than V");
  ST_REQUIRE(ws != nullptr && ws_floats >= ctc_ws_floats(B, T, S_max), "st_ctc: workspace too small");
  ST_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "st_ctc: workspace must be 8-byte aligned");
  float* lse = ws;
  double* ab = reinterpret_cast<double*>(ws + ((static_cast<int64_t>(B) * T + 63) & ~int64_t(63)));
  double* nll_d = ab + 2ll * B * T * S_max;
  const int64_t rows = static_cast<int64_t>(B) * T;
  const int cap = num_sms() * 8;
  ST_REQUIRE(!grad_only || grad, "st_ctc_grad: grad is required");
  if (!grad_only) {
    ProfScope prof(stream, PROF_CTC, 1.0 * rows * V * 4);
    ctc_lse_kernel<<<static_cast<int>(rows < cap ? rows : cap), CTC_THREADS, 0, stream>>>(logits, ld, rows, V, lse);
    ST_CHECK_LAUNCH();
  }
  if (!grad_only) {
    const int threads = ((S_max + 31) / 32) * 32;
    const int smem = 2 * (S_max + 4) * static_cast<int>(sizeof(double));
    ProfScope prof(stream, PROF_CTC, 2.0 * rows * S_max * 4);
    ctc_alpha_beta_kernel<<<dim3(B, 2), threads, smem, stream>>>(logits, ld, lse, targets, ld_tgt, in_len, tgt_len, blank, B, T, V,
                                                                  S_max, ab, nll_d, nll);
    ST_CHECK_LAUNCH();
  }
  if (grad) {
    ProfScope prof(stream, PROF_CTC, 2.0 * rows * V * 4);
    ctc_grad_kernel<<<static_cast<int>(rows < cap ? rows : cap), CTC_THREADS, 0, stream>>>(logits, ld, lse, targets, ld_tgt, in_len,
                                                                                           tgt_len, blank, B, T, V, S_max, ab, nll_d,
                                                                                           scale, grad, ldg);
    ST_CHECK_LAUNCH();
  }
  return ST_OK;
}

}  // namespace st
