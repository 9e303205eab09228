// st_api.cu — the extern "C" boundary (include/st_b200.h) and the host-side orchestration of the
// composite operators: MultiHeadAttention (transformer/Attention.py:64-96) and
// PositionwiseFeedForward (transformer/SubLayers.py:24-28), forward and backward.
//
// Precision contract: every tensor-core operand is rounded to TF32 with round-to-nearest where it
// is PRODUCED (GEMM / LayerNorm / attention epilogues) or, for tensors that arrive from outside the
// library, by one explicit rounding pass; accumulation is fp32.
#include <math.h>
#include <string.h>

#include "../../include/st_b200.h"
#include "st_gemm.cuh"
#include "st_host.h"
#include "st_kernels.h"
#include "st_attn.cuh"

namespace st {

DropoutCfg make_dropout(float p, uint64_t seed) {
  DropoutCfg c;
  if (p > 0.f) {  // drop probability quantised to thresh/65536; the scale uses the quantised value (unbiased)
    long t = lround(static_cast<double>(p) * 65536.0);
    if (t < 1) t = 1;
    if (t > 65535) t = 65535;
    c.thresh = static_cast<uint32_t>(t);
    c.scale = 65536.f / static_cast<float>(65536 - t);
    c.seed = seed;
    double t32 = floor(static_cast<double>(p) * 4294967296.0);
    if (t32 < 1.0) t32 = 1.0;
    if (t32 > 4294967295.0) t32 = 4294967295.0;
    c.thresh32 = static_cast<uint32_t>(t32);
    c.scale32 = static_cast<float>(4294967296.0 / (4294967296.0 - t32));
  }
  return c;
}

int selftest(int which, double* err);
int mma_bench(int variant, int n, int iters, double* clk_per_mma);
int selftest_count();

namespace {

inline size_t esz(int dt) { return dt == ST_DTYPE_F32 ? 4 : 2; }
inline bool is16(int dt) { return dt == ST_DTYPE_F16 || dt == ST_DTYPE_BF16; }
inline bool is_mixed(int dt) { return dt == ST_DTYPE_F32_H16; }
// element type of the INTERNAL tensors of a composite operator (its boundary tensors are fp32 in mixed mode)
inline int internal_dt(int dt) { return is_mixed(dt) ? ST_DTYPE_F16 : dt; }
// floats occupied by n activation elements of type dt
inline int64_t act_floats(int dt, int64_t n) { return dt == ST_DTYPE_F32 ? n : (n + 1) / 2; }
// pointer to element `off` of an activation buffer
inline void* at(void* p, int dt, int64_t off) { return static_cast<char*>(p) + off * static_cast<int64_t>(esz(dt)); }
inline const void* at(const void* p, int dt, int64_t off) { return static_cast<const char*>(p) + off * static_cast<int64_t>(esz(dt)); }

// split-K factor for a weight-gradient GEMM: enough CTAs to fill the machine ~2x, at least 4 k-blocks each
int wgrad_splits(int m_out, int n_out, int64_t k_len, int dt) {
  const int bn = n_out > 128 ? 256 : (n_out > 64 ? 128 : 64);
  const int tiles = ((m_out + 127) / 128) * ((n_out + bn - 1) / bn);
  const int bk = dt == ST_DTYPE_F32 ? 32 : 64;
  const int64_t kblocks = (k_len + bk - 1) / bk;
  int s = (2 * num_sms() + tiles - 1) / tiles;
  const int64_t max_s = kblocks / 4 > 0 ? kblocks / 4 : 1;
  if (s > max_s) s = static_cast<int>(max_s);
  return s < 1 ? 1 : s;
}

// dW[out,in] (fp32) = dY[rows,out]^T * X[rows,in]   (overwrites dW); dY and X of type dt
// zeroed: dW already holds zeros (st_*_bwd_args.grads_zeroed) — no clear needed before the split-K reductions
int wgrad(cudaStream_t s, int dt, const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int rows, int n_out,
          int n_in, bool zeroed = false, const float* unscale_amax = nullptr) {
  if (!zeroed) ST_CHECK_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(n_out) * n_in * sizeof(float), s));
  GemmEpilogue ep;
  ep.atomic = 1;
  ep.unscale_amax = unscale_amax;
  return gemm_any(s, dt, GEMM_TN, dy, lddy, x, ldx, dw, n_in, 0, n_out, n_in, rows, ep, wgrad_splits(n_out, n_in, rows, dt));
}

// clear a gradient vector unless the caller says it is already zero
#define ST_CLEAR(ptr, n)                                                                  \
  do {                                                                                    \
    if (!zeroed) ST_CHECK_CUDA(cudaMemsetAsync((ptr), 0, (n) * sizeof(float), s));        \
  } while (0)

struct Carver {
  float* base;
  int64_t used = 0;
  int64_t cap;
  Carver(float* b, int64_t c) : base(b), cap(c) {}
  float* take(int64_t n) {
    n = (n + 63) & ~int64_t(63);  // keep every sub-buffer 256-byte aligned
    float* p = base + used;
    used += n;
    return p;
  }
  void* take_act(int dt, int64_t n) { return take(act_floats(dt, n)); }
  bool ok() const { return base != nullptr && used <= cap; }
};
int64_t pad64(int64_t n) { return (n + 63) & ~int64_t(63); }
int64_t pad64a(int dt, int64_t n) { return pad64(act_floats(dt, n)); }

// operand-precision copy of an fp32 weight matrix: TF32 rounding (fp32 path) or conversion to the 16-bit type
int weight_copy(cudaStream_t s, int dt, const float* w, void* dst, int64_t rows, int cols) {
  if (dt == ST_DTYPE_F32) return round_tf32_2d(s, w, cols, static_cast<float*>(dst), cols, rows, cols);
  return cast_2d(s, w, ST_DTYPE_F32, cols, dst, dt, cols, rows, cols);
}

// ------------------------------------------------------------------ MHA buffer plans
struct MhaPlan {
  bool same_qkv, same_kv;
  int64_t M, Mk;
  int d, dt;
  // saved
  const void *xq_r, *xk_r, *xv_r;   // GEMM-ready inputs (alias the inputs when inputs_tf32 or 16-bit)
  void *projq, *projk, *projv;
  int64_t ldpq, ldpk, ldpv;
  void* ctx;
  float *lse, *z, *mean, *rstd;
  void *w_r /*[3d,d] q,k,v*/, *wo_r;
  float* b_pack /*[3d]*/;
  bool pre_rounded;
};

bool mha_pre_rounded(const st_mha_args& a) {
  const int64_t dd = static_cast<int64_t>(a.d_model) * a.d_model;
  const int dt = internal_dt(a.dtype);
  return a.wq_tf32 && a.wk_tf32 && a.wv_tf32 && a.wo_tf32 && a.wk_tf32 == at(a.wq_tf32, dt, dd) && a.wv_tf32 == at(a.wk_tf32, dt, dd) &&
         a.bk == a.bq + a.d_model && a.bv == a.bk + a.d_model;
}

int64_t mha_saved_floats(int dt_, int B, int Lq, int Lk, int H, int d, bool same_qkv, bool same_kv, bool inputs_tf32) {
  const int64_t M = static_cast<int64_t>(B) * Lq, Mk = static_cast<int64_t>(B) * Lk;
  const int dt = internal_dt(dt_);
  int64_t n = 0;
  if (is_mixed(dt_)) {            // fp16 operand copies of the fp32 inputs, unless the caller supplies them (q_h16 ...)
    if (!inputs_tf32) {
      n += pad64a(dt, M * d);
      if (!same_qkv) { n += pad64a(dt, Mk * d); if (!same_kv) n += pad64a(dt, Mk * d); }
    }
  } else if (dt == ST_DTYPE_F32 && !inputs_tf32) {
    n += pad64(M * d);
    if (!same_qkv) { n += pad64(Mk * d); if (!same_kv) n += pad64(Mk * d); }
  }
  n += same_qkv ? pad64a(dt, M * 3 * d) : pad64a(dt, M * d) + (same_kv ? pad64a(dt, Mk * 2 * d) : 2 * pad64a(dt, Mk * d));
  n += pad64a(dt, M * d);                          // ctx
  n += pad64(static_cast<int64_t>(B) * H * Lq);    // lse
  n += pad64(M * d) + 2 * pad64(M);                // z, mean, rstd
  n += pad64a(dt, 3ll * d * d) + pad64(3 * d) + pad64a(dt, static_cast<int64_t>(d) * d);
  return n;
}

int plan_mha(const st_mha_args& a, MhaPlan& p) {
  p.same_qkv = (a.q_in == a.k_in && a.k_in == a.v_in && a.Lq == a.Lk);
  p.same_kv = (a.k_in == a.v_in);
  p.M = static_cast<int64_t>(a.B) * a.Lq;
  p.Mk = static_cast<int64_t>(a.B) * a.Lk;
  p.d = a.d_model;
  p.dt = internal_dt(a.dtype);
  const int d = a.d_model, dt = p.dt;
  Carver c(a.saved, a.saved_floats);
  if (is_mixed(a.dtype) && a.inputs_tf32) {      // the caller's fp16 copies
    ST_REQUIRE(a.q_h16 != nullptr, "st_mha: inputs_tf32 with ST_DTYPE_F32_H16 needs q_h16");
    p.xq_r = a.q_h16;
    if (p.same_qkv) { p.xk_r = p.xv_r = p.xq_r; }
    else {
      p.xk_r = (a.k_in == a.q_in && a.Lq == a.Lk) ? a.q_h16 : a.k_h16;
      p.xv_r = p.same_kv ? p.xk_r : ((a.v_in == a.q_in && a.Lq == a.Lk) ? a.q_h16 : a.v_h16);
      ST_REQUIRE(p.xk_r && p.xv_r, "st_mha: inputs_tf32 with ST_DTYPE_F32_H16 needs k_h16 / v_h16 for inputs other than q_in");
    }
  } else if (is_mixed(a.dtype)) {
    p.xq_r = c.take_act(dt, p.M * d);
    if (p.same_qkv) { p.xk_r = p.xv_r = p.xq_r; }
    else {
      p.xk_r = c.take_act(dt, p.Mk * d);
      p.xv_r = p.same_kv ? p.xk_r : c.take_act(dt, p.Mk * d);
    }
  } else if (dt != ST_DTYPE_F32 || a.inputs_tf32) {
    p.xq_r = a.q_in; p.xk_r = a.k_in; p.xv_r = a.v_in;
  } else {
    p.xq_r = c.take(p.M * d);
    if (p.same_qkv) { p.xk_r = p.xv_r = p.xq_r; }
    else {
      p.xk_r = c.take(p.Mk * d);
      p.xv_r = p.same_kv ? p.xk_r : c.take(p.Mk * d);
    }
  }
  if (p.same_qkv) {
    void* b = c.take_act(dt, p.M * 3 * d);
    p.projq = b; p.projk = at(b, dt, d); p.projv = at(b, dt, 2 * d);
    p.ldpq = p.ldpk = p.ldpv = 3 * d;
  } else {
    p.projq = c.take_act(dt, p.M * d); p.ldpq = d;
    if (p.same_kv) {
      void* b = c.take_act(dt, p.Mk * 2 * d);
      p.projk = b; p.projv = at(b, dt, d); p.ldpk = p.ldpv = 2 * d;
    } else {
      p.projk = c.take_act(dt, p.Mk * d); p.projv = c.take_act(dt, p.Mk * d); p.ldpk = p.ldpv = d;
    }
  }
  p.ctx = c.take_act(dt, p.M * d);
  p.lse = c.take(static_cast<int64_t>(a.B) * a.H * a.Lq);
  p.z = c.take(p.M * d);
  p.mean = c.take(p.M);
  p.rstd = c.take(p.M);
  p.w_r = c.take_act(dt, 3ll * d * d);
  p.b_pack = c.take(3 * d);
  p.wo_r = c.take_act(dt, static_cast<int64_t>(d) * d);
  p.pre_rounded = mha_pre_rounded(a);
  if (p.pre_rounded) {   // caller-maintained operand-precision weights, already packed: nothing to round, copy or save
    p.w_r = const_cast<void*>(a.wq_tf32);
    p.b_pack = const_cast<float*>(a.bq);
    p.wo_r = const_cast<void*>(a.wo_tf32);
  }
  if (!c.ok()) {
    set_error("st_mha: saved buffer too small (%lld floats given, %lld needed)", (long long)a.saved_floats, (long long)c.used);
    return ST_ERR_WORKSPACE;
  }
  return ST_OK;
}

int check_dtype(int dt, const char* who, bool composite = false) {
  ST_REQUIRE(dt == ST_DTYPE_F32 || dt == ST_DTYPE_F16 || dt == ST_DTYPE_BF16 || (composite && dt == ST_DTYPE_F32_H16),
             "%s: bad dtype %d", who, dt);
  return ST_OK;
}

int check_mha(const st_mha_args& a) {
  ST_TRY(check_dtype(a.dtype, "st_mha", true));
  ST_REQUIRE(a.B > 0 && a.Lq > 0 && a.Lk > 0 && a.H > 0, "st_mha: empty problem");
  ST_REQUIRE(a.d_model == a.H * a.dk, "st_mha: d_model (%d) != n_head (%d) * d_k (%d)", a.d_model, a.H, a.dk);
  if (is_mixed(a.dtype))
    ST_REQUIRE(a.residual == a.q_in || a.residual == a.k_in || a.residual == a.v_in,
               "st_mha: with ST_DTYPE_F32_H16 the residual must be one of the inputs");
  if (is16(a.dtype) || is_mixed(a.dtype)) ST_REQUIRE(a.dk == 64, "st_mha: 16-bit operands need d_k = 64 (got %d)", a.dk);
  else ST_REQUIRE(a.dk == 32 || a.dk == 64 || a.dk == 128, "st_mha: d_k must be 32, 64 or 128 (got %d)", a.dk);
  return ST_OK;
}

int64_t mha_ws_floats(int dt_, int B, int Lq, int Lk, int H, int d) {
  const int64_t M = static_cast<int64_t>(B) * Lq, Mk = static_cast<int64_t>(B) * Lk;
  const int dt = internal_dt(dt_);
  // backward: dz, dctx, dprojq, dprojk, dprojv, delta, amax
  return 2 * pad64a(dt, M * d) + pad64a(dt, M * 3 * d) + 2 * pad64a(dt, Mk * 2 * d) + pad64(static_cast<int64_t>(B) * H * Lq) + 128;
}

void fill_attn(AttnArgs& at_, const st_mha_args& a, const MhaPlan& p) {
  at_.B = a.B; at_.H = a.H; at_.Lq = a.Lq; at_.Lk = a.Lk; at_.dk = a.dk; at_.dtype = p.dt;
  at_.q = p.projq; at_.ldq = p.ldpq; at_.k = p.projk; at_.ldk = p.ldpk; at_.v = p.projv; at_.ldv = p.ldpv;
  at_.mask = a.mask; at_.ms_b = a.ms_b; at_.ms_q = a.ms_q; at_.ms_k = a.ms_k;
  at_.k_len = a.k_len; at_.causal = a.causal;
  at_.scale = 1.f / sqrtf(static_cast<float>(a.dk));
  at_.drop = make_dropout(a.dropout_p, a.seed);
  at_.ctx = p.ctx; at_.ldctx = a.d_model; at_.lse = p.lse; at_.attn = nullptr;
}

}  // namespace
}  // namespace st

using namespace st;

extern "C" {

int st_version(void) { return 100; }
const char* st_last_error(void) { return last_error(); }

int st_device_check(int device) {
  cudaDeviceProp prop;
  ST_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d (%s); this library only runs on sm_100 (B200)", device, prop.major, prop.minor, prop.name);
    return ST_ERR_DEVICE;
  }
  return ST_OK;
}

int st_set_option(const char* name, int v) { return set_option(name, v); }
int64_t st_launch_count(void) { return static_cast<int64_t>(launch_count()); }
int st_profile_enable(int on) { profile_enable(on); return ST_OK; }
int st_profile_reset(void) { profile_reset(); return ST_OK; }
int st_profile_dump(const char* path) { return profile_dump(path); }
int st_profile_classes(void) { return PROF_NUM; }
const char* st_profile_class_name(int cls) {
  static const char* names[PROF_NUM] = {"gemm_tf32", "attn_fwd", "attn_bwd_dkv", "attn_bwd_dq", "attn_bwd_delta", "add_ln_fwd",
                                        "add_ln_bwd", "round_tf32", "colsum", "lsce", "sumsq", "adam", "embed", "ctc"};
  return (cls >= 0 && cls < PROF_NUM) ? names[cls] : "?";
}
int st_profile_read(int cls, double* ms, double* work, int64_t* launches) {
  long long n = 0;
  const int st = profile_read(cls, ms, work, &n);
  *launches = n;
  return st;
}
int st_debug_read_fwd_trace(uint64_t* host_out, int n) { return attn_read_fwd_trace(reinterpret_cast<unsigned long long*>(host_out), n); }
int st_debug_read_trace(uint64_t* host_out, int n) { return attn_read_trace(reinterpret_cast<unsigned long long*>(host_out), n); }
int st_debug_mma_bench(int variant, int n, int iters, double* clk_per_mma) { return mma_bench(variant, n, iters, clk_per_mma); }
int st_selftest_count(void) { return selftest_count(); }
int st_selftest(int which, double* rel_err_out) { return selftest(which, rel_err_out); }

int st_add_ln_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out, float* z_out,
                  float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_tf32, float dropout_p,
                  uint64_t seed, cudaStream_t stream) {
  return add_ln_fwd(stream, a, b, gamma, beta, out, z_out, mean_out, rstd_out, rows, d, eps, round_tf32,
                    make_dropout(dropout_p, seed));
}

int st_add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                  float* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, int round_tf32,
                  float dropout_p, uint64_t seed, cudaStream_t stream) {
  return add_ln_bwd(stream, dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dzsum, rows, d, round_tf32,
                    make_dropout(dropout_p, seed));
}

int st_lsce_fwd_bwd(const float* logits, int64_t ld_logits, const int64_t* target, const float* one_hot,
                    const float* weight, float confidence, int64_t padding_idx, int size_average, int64_t N, int V,
                    float* row_loss, float* loss, float* grad, int64_t ld_grad, cudaStream_t stream) {
  LsceArgs a{};
  a.logits = logits; a.ldl = ld_logits; a.target = target; a.q_dense = nullptr; a.one_hot = one_hot; a.weight = weight;
  a.confidence = confidence; a.padding_idx = padding_idx;
  a.inv_z = (size_average && N > 0) ? 1.f / static_cast<float>(N) : 1.f;
  a.N = N; a.V = V; a.row_loss = row_loss; a.loss = loss; a.grad = grad; a.ldg = ld_grad;
  ST_REQUIRE(target != nullptr && one_hot != nullptr, "st_lsce_fwd_bwd: target and one_hot are required");
  return lsce_fwd_bwd(stream, a);
}

int st_softce_fwd_bwd(const float* logits, int64_t ld_logits, const float* q, const float* weight, int size_average,
                      int64_t N, int V, float* row_loss, float* loss, float* grad, int64_t ld_grad,
                      cudaStream_t stream) {
  LsceArgs a{};
  a.logits = logits; a.ldl = ld_logits; a.target = nullptr; a.q_dense = q; a.one_hot = nullptr; a.weight = weight;
  a.confidence = 0.f; a.padding_idx = -1;
  a.inv_z = (size_average && N > 0) ? 1.f / static_cast<float>(N) : 1.f;
  a.N = N; a.V = V; a.row_loss = row_loss; a.loss = loss; a.grad = grad; a.ldg = ld_grad;
  ST_REQUIRE(q != nullptr, "st_softce_fwd_bwd: q is required");
  return lsce_fwd_bwd(stream, a);
}

int st_round_tf32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, cudaStream_t stream) {
  return round_tf32_2d(stream, src, lds, dst, ldd, rows, cols);
}
int st_colsum_add(const float* x, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t stream) {
  return colsum_add(stream, x, ld, rows, cols, out);
}

int st_gemm(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
            int K, const st_gemm_epilogue* e, cudaStream_t stream) {
  ST_REQUIRE(mode >= 0 && mode <= 2, "st_gemm: bad mode %d", mode);
  GemmEpilogue ep;
  int splits = 1;
  if (e) {
    ep.bias = e->bias; ep.aux = e->aux; ep.ldaux = e->ldaux; ep.aux_mode = e->aux_mode; ep.relu = e->relu;
    ep.round_tf32 = e->round_tf32;
    const DropoutCfg dc = make_dropout(e->dropout_p, e->seed);
    ep.drop_thresh = dc.thresh; ep.drop_scale = dc.scale; ep.drop_seed = dc.seed;
    splits = e->k_splits > 1 ? e->k_splits : 1;
    ep.atomic = splits > 1;
    ST_REQUIRE(!ep.aux_mode || ep.aux, "st_gemm: aux_mode set without aux");
  }
  return gemm_tf32(stream, static_cast<GemmMode>(mode), A, lda, B, ldb, C, ldc, M, N, K, ep, splits);
}

int st_sumsq(const float* x, int64_t n, float* out, cudaStream_t stream) { return sumsq_add(stream, x, n, out); }
int st_adam_step(const st_adam_args* a, cudaStream_t stream) {
  ST_REQUIRE(a != nullptr, "st_adam_step: null args");
  return adam_step(stream, a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->n, a->lr, a->beta1, a->beta2, a->eps,
                   a->step, a->max_grad_norm, a->grad_scale, a->norm_ws, a->param_tf32, a->twin_dtype);
}

// ------------------------------------------------------------------ attention core
static AttnArgs to_attn(const st_attn_args& a) {
  AttnArgs r{};
  r.B = a.B; r.H = a.H; r.Lq = a.Lq; r.Lk = a.Lk; r.dk = a.dk; r.dtype = a.dtype;
  r.q = a.q; r.ldq = a.ldq; r.k = a.k; r.ldk = a.ldk; r.v = a.v; r.ldv = a.ldv;
  r.mask = a.mask; r.ms_b = a.ms_b; r.ms_q = a.ms_q; r.ms_k = a.ms_k;
  r.k_len = a.k_len; r.causal = a.causal;
  r.scale = 1.f / sqrtf(static_cast<float>(a.dk));
  r.drop = make_dropout(a.dropout_p, a.seed);
  r.ctx = a.ctx; r.ldctx = a.ldctx; r.lse = a.lse; r.attn = a.attn;
  return r;
}

int st_attn_fwd(const st_attn_args* a, cudaStream_t stream) {
  ST_REQUIRE(a != nullptr, "st_attn_fwd: null args");
  ST_TRY(check_dtype(a->dtype, "st_attn_fwd"));
  return attn_fwd(stream, to_attn(*a));
}

int st_attn_bwd(const st_attn_bwd_args* a, cudaStream_t stream) {
  ST_REQUIRE(a != nullptr, "st_attn_bwd: null args");
  ST_TRY(check_dtype(a->f.dtype, "st_attn_bwd"));
  AttnBwdArgs b{};
  b.f = to_attn(a->f);
  b.dctx = a->dctx; b.lddctx = a->lddctx; b.delta = a->delta;
  b.dq = a->dq; b.lddq = a->lddq; b.dk_ = a->dk; b.lddk = a->lddk; b.dv = a->dv; b.lddv = a->lddv;
  return attn_bwd(stream, b);
}

int st_cast(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd, int64_t rows, int cols, float scale,
            cudaStream_t stream) {
  return cast_2d(stream, src, src_dtype, lds, dst, dst_dtype, ldd, rows, cols, scale);
}

int st_gemm_dt(int dtype, int mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int c_lp, int M,
               int N, int K, const st_gemm_epilogue* e, cudaStream_t stream) {
  ST_REQUIRE(mode >= 0 && mode <= 2, "st_gemm_dt: bad mode %d", mode);
  ST_TRY(check_dtype(dtype, "st_gemm_dt"));
  GemmEpilogue ep;
  int splits = 1;
  if (e) {
    ep.bias = e->bias; ep.aux = e->aux; ep.ldaux = e->ldaux; ep.aux_mode = e->aux_mode; ep.relu = e->relu;
    ep.round_tf32 = e->round_tf32;
    const DropoutCfg dc = make_dropout(e->dropout_p, e->seed);
    ep.drop_thresh = dc.thresh; ep.drop_scale = dc.scale; ep.drop_seed = dc.seed;
    splits = e->k_splits > 1 ? e->k_splits : 1;
    ep.atomic = splits > 1;
    ST_REQUIRE(!ep.aux_mode || ep.aux, "st_gemm_dt: aux_mode set without aux");
  }
  return gemm_any(stream, dtype, static_cast<GemmMode>(mode), A, lda, B, ldb, C, ldc, c_lp, M, N, K, ep, splits);
}

// ------------------------------------------------------------------ MultiHeadAttention
int64_t st_mha_saved_floats(int B, int Lq, int Lk, int H, int d_model, int same_qkv, int same_kv, int inputs_tf32) {
  return mha_saved_floats(ST_DTYPE_F32, B, Lq, Lk, H, d_model, same_qkv != 0, same_kv != 0 || same_qkv != 0, inputs_tf32 != 0);
}
int64_t st_mha_ws_floats(int B, int Lq, int Lk, int H, int d_model) { return mha_ws_floats(ST_DTYPE_F32, B, Lq, Lk, H, d_model); }
int64_t st_mha_saved_floats_dt(int dtype, int B, int Lq, int Lk, int H, int d_model, int same_qkv, int same_kv, int inputs_tf32) {
  return mha_saved_floats(dtype, B, Lq, Lk, H, d_model, same_qkv != 0, same_kv != 0 || same_qkv != 0, inputs_tf32 != 0);
}
int64_t st_mha_ws_floats_dt(int dtype, int B, int Lq, int Lk, int H, int d_model) { return mha_ws_floats(dtype, B, Lq, Lk, H, d_model); }

int st_mha_fwd(const st_mha_args* ap, cudaStream_t s) {
  ST_REQUIRE(ap != nullptr, "st_mha_fwd: null args");
  const st_mha_args& a = *ap;
  ST_TRY(check_mha(a));
  MhaPlan p;
  ST_TRY(plan_mha(a, p));
  const int d = a.d_model, dt = p.dt;
  const bool mixed = is_mixed(a.dtype);
  const int M = static_cast<int>(p.M), Mk = static_cast<int>(p.Mk);
  const int64_t dd = static_cast<int64_t>(d) * d;

  // 1. operand-precision copies of the weights, packed [wq; wk; wv] so that shared inputs need one GEMM
  if (!p.pre_rounded) {
    ST_TRY(weight_copy(s, dt, a.wq, p.w_r, d, d));
    ST_TRY(weight_copy(s, dt, a.wk, at(p.w_r, dt, dd), d, d));
    ST_TRY(weight_copy(s, dt, a.wv, at(p.w_r, dt, 2 * dd), d, d));
    ST_TRY(weight_copy(s, dt, a.wo, p.wo_r, d, d));
    ST_CHECK_CUDA(cudaMemcpyAsync(p.b_pack, a.bq, d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ST_CHECK_CUDA(cudaMemcpyAsync(p.b_pack + d, a.bk, d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ST_CHECK_CUDA(cudaMemcpyAsync(p.b_pack + 2 * d, a.bv, d * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  // 2. operand copies of the inputs: TF32 rounding (fp32 path), fp16 conversion (mixed: exact for TF32-representable
  //    inputs); 16-bit activations are GEMM operands as they are
  if (mixed && a.inputs_tf32) {
    // the caller's fp16 copies are the operands
  } else if (mixed) {
    ST_TRY(cast_2d(s, a.q_in, ST_DTYPE_F32, d, const_cast<void*>(p.xq_r), dt, d, M, d));
    if (!p.same_qkv) {
      ST_TRY(cast_2d(s, a.k_in, ST_DTYPE_F32, d, const_cast<void*>(p.xk_r), dt, d, Mk, d));
      if (!p.same_kv) ST_TRY(cast_2d(s, a.v_in, ST_DTYPE_F32, d, const_cast<void*>(p.xv_r), dt, d, Mk, d));
    }
  } else if (dt == ST_DTYPE_F32 && !a.inputs_tf32) {
    ST_TRY(round_tf32_2d(s, static_cast<const float*>(a.q_in), d, const_cast<float*>(static_cast<const float*>(p.xq_r)), d, M, d));
    if (!p.same_qkv) {
      ST_TRY(round_tf32_2d(s, static_cast<const float*>(a.k_in), d, const_cast<float*>(static_cast<const float*>(p.xk_r)), d, Mk, d));
      if (!p.same_kv)
        ST_TRY(round_tf32_2d(s, static_cast<const float*>(a.v_in), d, const_cast<float*>(static_cast<const float*>(p.xv_r)), d, Mk, d));
    }
  }
  // 3. projections (Attention.py:74-76), outputs in operand precision for the attention MMAs
  GemmEpilogue ep;
  ep.round_tf32 = 1;
  if (p.same_qkv) {
    ep.bias = p.b_pack;
    ST_TRY(gemm_any(s, dt, GEMM_NT, p.xq_r, d, p.w_r, d, p.projq, p.ldpq, 1, M, 3 * d, d, ep));
  } else {
    Fork fk(s);                           // the query projection and the key / value projections read different inputs
    cudaStream_t skv = fk.branch(0);
    ep.bias = p.b_pack;
    ST_TRY(gemm_any(s, dt, GEMM_NT, p.xq_r, d, p.w_r, d, p.projq, p.ldpq, 1, M, d, d, ep));
    if (p.same_kv) {
      ep.bias = p.b_pack + d;
      ST_TRY(gemm_any(skv, dt, GEMM_NT, p.xk_r, d, at(p.w_r, dt, dd), d, p.projk, p.ldpk, 1, Mk, 2 * d, d, ep));
    } else {
      ep.bias = p.b_pack + d;
      ST_TRY(gemm_any(skv, dt, GEMM_NT, p.xk_r, d, at(p.w_r, dt, dd), d, p.projk, p.ldpk, 1, Mk, d, d, ep));
      ep.bias = p.b_pack + 2 * d;
      ST_TRY(gemm_any(skv, dt, GEMM_NT, p.xv_r, d, at(p.w_r, dt, 2 * dd), d, p.projv, p.ldpv, 1, Mk, d, d, ep));
    }
    ST_TRY(fk.join());
  }
  // 4. attention core (Attention.py:78-90)
  AttnArgs at_{};
  fill_attn(at_, a, p);
  at_.attn = a.attn;
  ST_TRY(attn_fwd(s, at_));
  // 5. output projection + bias + residual (Attention.py:92,94): the pre-LayerNorm sum stays fp32
  GemmEpilogue eo;
  eo.bias = a.bo; eo.aux = a.residual; eo.ldaux = d; eo.aux_mode = 1;
  if (mixed) eo.aux = (a.residual == a.q_in) ? p.xq_r : (a.residual == a.k_in ? p.xk_r : p.xv_r);   // the fp16 copy of the residual
  ST_TRY(gemm_any(s, dt, GEMM_NT, p.ctx, d, p.wo_r, d, p.z, d, 0, M, d, d, eo));
  // 6. LayerNorm (Attention.py:94)
  // (mixed: the fp32 output is not rounded — its consumers convert it themselves, a TF32 rounding here would only be a second one)
  return add_ln_fwd_any(s, ST_DTYPE_F32, mixed ? ST_DTYPE_F32 : dt, p.z, nullptr, a.ln_g, a.ln_b, a.out, nullptr, p.mean, p.rstd, M, d,
                        a.eps, mixed ? 0 : a.round_out, DropoutCfg{}, nullptr, 0, mixed ? a.out_h16 : nullptr);
}

int st_mha_bwd(const st_mha_bwd_args* bp, cudaStream_t s) {
  ST_REQUIRE(bp != nullptr, "st_mha_bwd: null args");
  const st_mha_bwd_args& b = *bp;
  const bool zeroed = b.grads_zeroed != 0;
  const st_mha_args& a = b.f;
  ST_TRY(check_mha(a));
  MhaPlan p;
  ST_TRY(plan_mha(a, p));
  const int d = a.d_model, dt = p.dt;
  const bool mixed = is_mixed(a.dtype);
  const int M = static_cast<int>(p.M), Mk = static_cast<int>(p.Mk);
  const int64_t dd = static_cast<int64_t>(d) * d;
  ST_REQUIRE(a.ws != nullptr && a.ws_floats >= mha_ws_floats(a.dtype, a.B, a.Lq, a.Lk, a.H, d), "st_mha_bwd: workspace too small");
  Carver w(a.ws, a.ws_floats);
  void* dz = w.take_act(dt, p.M * d);
  void* dctx = w.take_act(dt, p.M * d);
  void *dpq, *dpk, *dpv;
  if (p.same_qkv) { void* t = w.take_act(dt, p.M * 3 * d); dpq = t; dpk = at(t, dt, d); dpv = at(t, dt, 2 * d); }
  else {
    dpq = w.take_act(dt, p.M * d);
    if (p.same_kv) { void* t = w.take_act(dt, p.Mk * 2 * d); dpk = t; dpv = at(t, dt, d); }
    else { dpk = w.take_act(dt, p.Mk * d); dpv = w.take_act(dt, p.Mk * d); }
  }
  float* delta = w.take(static_cast<int64_t>(a.B) * a.H * a.Lq);
  // mixed mode: the gradient scale of this call, a power of two derived on the device from max|dout| (st_common.cuh)
  float* ws_amax = mixed ? w.take(64) : nullptr;
  const float* amax = ws_amax;
  const bool fused_bias = dt == ST_DTYPE_F32 && attn_bwd_fuses_bias(a.dk);

  // LayerNorm backward; dbo = column sums of dz
  ST_CLEAR(b.dln_g, d);
  ST_CLEAR(b.dln_b, d);
  ST_CLEAR(b.dbo, d);
  if (mixed) {
    if (b.dout_amax) amax = b.dout_amax;     // measured by the operator that produced dout
    else ST_TRY(amax_abs(s, static_cast<const float*>(b.dout), p.M * d, ws_amax));
    ST_TRY(add_ln_bwd_mixed(s, static_cast<const float*>(b.dout), p.z, p.mean, p.rstd, a.ln_g, dz, b.dln_g, b.dln_b, b.dbo, M, d,
                            DropoutCfg{}, amax, b.dq_amax));
  } else {
    ST_TRY(add_ln_bwd_any(s, dt, b.dout, p.z, p.mean, p.rstd, a.ln_g, dz, b.dln_g, b.dln_b, b.dbo, M, d, 1, DropoutCfg{}));
  }
  // output projection backward.  Weight gradients and bias column sums never feed another kernel of this call: they run on
  // a side stream (st_host.h Fork) next to the chain dz -> dctx -> attention backward -> input gradients.
  Fork fk(s);
  {
    cudaStream_t sw = fk.branch(0);
    GemmEpilogue e; e.round_tf32 = 1;
    ST_TRY(gemm_any(s, dt, GEMM_NN, dz, d, p.wo_r, d, dctx, d, 1, M, d, d, e));
    ST_TRY(wgrad(sw, dt, dz, d, p.ctx, d, b.dwo, M, d, d, zeroed, amax));
  }
  // attention core backward
  {
    AttnBwdArgs ab{};
    fill_attn(ab.f, a, p);
    ab.dctx = dctx; ab.lddctx = d; ab.delta = delta;
    ab.dq = dpq; ab.lddq = p.ldpq; ab.dk_ = dpk; ab.lddk = p.ldpk; ab.dv = dpv; ab.lddv = p.ldpv;
    if (fused_bias) {   // dbq / dbk / dbv = column sums of dq / dk / dv, accumulated by the kernels' epilogues
      ST_CLEAR(b.dbq, d);
      ST_CLEAR(b.dbk, d);
      ST_CLEAR(b.dbv, d);
      ab.dbq = b.dbq; ab.dbk = b.dbk; ab.dbv = b.dbv;
    }
    ST_TRY(attn_bwd(s, ab));
  }
  // projection bias / weight gradients.  When the caller hands out dW / db as slices of one packed [wq; wk; wv]
  // buffer (functional.py does) and the projections share their input, one GEMM / column sum covers all of them.
  const bool pack_kv = p.same_kv && b.dwv == b.dwk + dd && b.dbv == b.dbk + d;
  const bool pack_qkv = p.same_qkv && pack_kv && b.dwk == b.dwq + dd && b.dbk == b.dbq + d;
  {
    cudaStream_t s_main = s;
    cudaStream_t s = fk.branch(0);     // (shadows the caller's stream inside this block: ST_CLEAR and the kernels below)
    (void)s_main;
    if (pack_qkv) {
      if (!fused_bias) {
        ST_CLEAR(b.dbq, 3 * d);
        ST_TRY(colsum_add_any(s, dt, dpq, p.ldpq, M, 3 * d, b.dbq, amax));
      }
      ST_TRY(wgrad(s, dt, dpq, p.ldpq, p.xq_r, d, b.dwq, M, 3 * d, d, zeroed, amax));
    } else {
      if (!fused_bias) {
        ST_CLEAR(b.dbq, d);
        ST_TRY(colsum_add_any(s, dt, dpq, p.ldpq, M, d, b.dbq, amax));
      }
      ST_TRY(wgrad(s, dt, dpq, p.ldpq, p.xq_r, d, b.dwq, M, d, d, zeroed, amax));
      if (pack_kv) {
        if (!fused_bias) {
          ST_CLEAR(b.dbk, 2 * d);
          ST_TRY(colsum_add_any(s, dt, dpk, p.ldpk, Mk, 2 * d, b.dbk, amax));
        }
        ST_TRY(wgrad(s, dt, dpk, p.ldpk, p.xk_r, d, b.dwk, Mk, 2 * d, d, zeroed, amax));
      } else {
        if (!fused_bias) {
          ST_CLEAR(b.dbk, d);
          ST_CLEAR(b.dbv, d);
          ST_TRY(colsum_add_any(s, dt, dpk, p.ldpk, Mk, d, b.dbk, amax));
          ST_TRY(colsum_add_any(s, dt, dpv, p.ldpv, Mk, d, b.dbv, amax));
        }
        ST_TRY(wgrad(s, dt, dpk, p.ldpk, p.xk_r, d, b.dwk, Mk, d, d, zeroed, amax));
        ST_TRY(wgrad(s, dt, dpv, p.ldpv, p.xv_r, d, b.dwv, Mk, d, d, zeroed, amax));
      }
    }
  }
  // input gradients; the residual branch contributes dz to EXACTLY ONE input buffer: the first of q, k, v that aliases the
  // residual tensor (aliased inputs receive the sum of their buffers from the caller)
  const bool res_q = (a.residual == a.q_in), res_k = !res_q && (a.residual == a.k_in),
             res_v = !res_q && !res_k && (a.residual == a.v_in);
  // mixed mode: the input gradients leave the operator as fp32 and shed the gradient scale in the epilogue
  const int out_lp = mixed ? 0 : 1;
  auto with_res = [&](bool on, float* amax_out = nullptr) {
    GemmEpilogue e;
    if (on) { e.aux = dz; e.ldaux = d; e.aux_mode = 1; }
    e.unscale_amax = amax;
    e.amax_out = mixed ? amax_out : nullptr;
    return e;
  };
  if (p.same_qkv) {
    ST_TRY(gemm_any(s, dt, GEMM_NN, dpq, p.ldpq, p.w_r, d, b.dq_in, d, out_lp, M, d, 3 * d, with_res(res_q, b.dq_amax)));
  } else {
    ST_TRY(gemm_any(s, dt, GEMM_NN, dpq, p.ldpq, p.w_r, d, b.dq_in, d, out_lp, M, d, d, with_res(res_q, b.dq_amax)));
    if (p.same_kv) {
      ST_TRY(gemm_any(s, dt, GEMM_NN, dpk, p.ldpk, at(p.w_r, dt, dd), d, b.dk_in, d, out_lp, Mk, d, 2 * d, with_res(res_k || res_v)));
    } else {
      ST_TRY(gemm_any(s, dt, GEMM_NN, dpk, p.ldpk, at(p.w_r, dt, dd), d, b.dk_in, d, out_lp, Mk, d, d, with_res(res_k)));
      ST_TRY(gemm_any(s, dt, GEMM_NN, dpv, p.ldpv, at(p.w_r, dt, 2 * dd), d, b.dv_in, d, out_lp, Mk, d, d, with_res(res_v)));
    }
  }
  if (!(res_q || res_k || res_v)) {
    ST_REQUIRE(b.dresidual != nullptr, "st_mha_bwd: dresidual is required when residual is a separate tensor");
    ST_CHECK_CUDA(cudaMemcpyAsync(b.dresidual, dz, p.M * d * esz(dt), cudaMemcpyDeviceToDevice, s));
  }
  return fk.join();
}

// ------------------------------------------------------------------ PositionwiseFeedForward
namespace {
struct FfnPlan { const void* x_r; void* h; float *z, *mean, *rstd; void *w1_r, *w2_r; };
int plan_ffn(const st_ffn_args& a, FfnPlan& p) {
  const int dt = internal_dt(a.dtype);
  Carver c(a.saved, a.saved_floats);
  // NOTE: the hidden activation must stay the FIRST 16-bit / fp32 tensor after the optional input copy (st_ffn_hidden_offset)
  if (is_mixed(a.dtype) && a.x_is_tf32) {
    ST_REQUIRE(a.x_h16 != nullptr, "st_ffn: x_is_tf32 with ST_DTYPE_F32_H16 needs x_h16");
    p.x_r = a.x_h16;
  } else if (is_mixed(a.dtype)) p.x_r = c.take_act(dt, a.rows * a.d_model);
  else p.x_r = (dt != ST_DTYPE_F32 || a.x_is_tf32) ? a.x : c.take(a.rows * a.d_model);
  p.h = c.take_act(dt, a.rows * a.d_ff);
  p.z = c.take(a.rows * a.d_model);
  p.mean = c.take(a.rows);
  p.rstd = c.take(a.rows);
  p.w1_r = c.take_act(dt, static_cast<int64_t>(a.d_ff) * a.d_model);
  p.w2_r = c.take_act(dt, static_cast<int64_t>(a.d_ff) * a.d_model);
  if (a.w1_tf32 && a.w2_tf32) { p.w1_r = const_cast<void*>(a.w1_tf32); p.w2_r = const_cast<void*>(a.w2_tf32); }
  if (!c.ok()) {
    set_error("st_ffn: saved buffer too small (%lld floats given, %lld needed)", (long long)a.saved_floats, (long long)c.used);
    return ST_ERR_WORKSPACE;
  }
  return ST_OK;
}
constexpr uint64_t kSeedMix1 = 0x5DEECE66Dull, kSeedMix2 = 0xB5297A4D3F84D5B5ull;
int64_t ffn_input_copy_floats(int dt_, int64_t rows, int d_model, int x_is_tf32) {
  if (is_mixed(dt_)) return x_is_tf32 ? 0 : pad64a(ST_DTYPE_F16, rows * d_model);   // x_is_tf32: the caller supplies x_h16
  return (dt_ != ST_DTYPE_F32 || x_is_tf32) ? 0 : pad64(rows * d_model);
}
int64_t ffn_saved_floats(int dt_, int64_t rows, int d_model, int d_ff, int x_is_tf32) {
  const int dt = internal_dt(dt_);
  return ffn_input_copy_floats(dt_, rows, d_model, x_is_tf32) + pad64a(dt, rows * d_ff) + pad64(rows * d_model) +
         2 * pad64(rows) + 2 * pad64a(dt, static_cast<int64_t>(d_ff) * d_model);
}
int64_t ffn_ws_floats(int dt_, int64_t rows, int d_model, int d_ff) {
  const int dt = internal_dt(dt_);
  return pad64a(dt, rows * d_model) + pad64a(dt, rows * d_ff) + 128;
}
}  // namespace

int64_t st_ffn_saved_floats(int64_t rows, int d_model, int d_ff, int x_is_tf32) {
  return ffn_saved_floats(ST_DTYPE_F32, rows, d_model, d_ff, x_is_tf32);
}
int64_t st_ffn_saved_floats_dt(int dtype, int64_t rows, int d_model, int d_ff, int x_is_tf32) {
  return ffn_saved_floats(dtype, rows, d_model, d_ff, x_is_tf32);
}
int64_t st_ffn_hidden_offset(int64_t rows, int d_model, int d_ff, int x_is_tf32) {
  (void)d_ff;
  return x_is_tf32 ? 0 : pad64(rows * d_model);
}
int64_t st_ffn_hidden_offset_dt(int dtype, int64_t rows, int d_model, int d_ff, int x_is_tf32) {
  (void)d_ff;
  return ffn_input_copy_floats(dtype, rows, d_model, x_is_tf32);
}
int64_t st_ffn_ws_floats(int64_t rows, int d_model, int d_ff) { return ffn_ws_floats(ST_DTYPE_F32, rows, d_model, d_ff); }
int64_t st_ffn_ws_floats_dt(int dtype, int64_t rows, int d_model, int d_ff) { return ffn_ws_floats(dtype, rows, d_model, d_ff); }

int st_ffn_fwd(const st_ffn_args* ap, cudaStream_t s) {
  ST_REQUIRE(ap != nullptr, "st_ffn_fwd: null args");
  const st_ffn_args& a = *ap;
  ST_TRY(check_dtype(a.dtype, "st_ffn", true));
  ST_REQUIRE(a.rows > 0 && a.rows < (1ll << 31) && a.d_model > 0 && a.d_ff > 0, "st_ffn: bad shape");
  FfnPlan p;
  ST_TRY(plan_ffn(a, p));
  const int M = static_cast<int>(a.rows), d = a.d_model, f = a.d_ff, dt = internal_dt(a.dtype);
  const bool mixed = is_mixed(a.dtype);
  if (!(a.w1_tf32 && a.w2_tf32)) {
    ST_TRY(weight_copy(s, dt, a.w1, p.w1_r, f, d));
    ST_TRY(weight_copy(s, dt, a.w2, p.w2_r, d, f));
  }
  if (mixed && a.x_is_tf32) {
    // the caller's fp16 copy is the operand
  } else if (mixed)
    ST_TRY(cast_2d(s, a.x, ST_DTYPE_F32, d, const_cast<void*>(p.x_r), dt, d, M, d));
  else if (dt == ST_DTYPE_F32 && !a.x_is_tf32)
    ST_TRY(round_tf32_2d(s, static_cast<const float*>(a.x), d, const_cast<float*>(static_cast<const float*>(p.x_r)), d, M, d));
  // h = dropout1(relu(fc1(x)))                                           SubLayers.py:25
  GemmEpilogue e1;
  e1.bias = a.b1; e1.relu = 1; e1.round_tf32 = 1;
  const DropoutCfg d1 = make_dropout(a.dropout_p, a.seed ^ kSeedMix1);
  e1.drop_thresh = d1.thresh; e1.drop_scale = d1.scale; e1.drop_seed = d1.seed;
  ST_TRY(gemm_any(s, dt, GEMM_NT, p.x_r, d, p.w1_r, d, p.h, f, 1, M, f, d, e1));
  // z = x + fc2(h)  (fp32)                                               SubLayers.py:26-27
  GemmEpilogue e2;
  e2.bias = a.b2; e2.aux = mixed ? p.x_r : a.x; e2.ldaux = d; e2.aux_mode = 1;
  ST_TRY(gemm_any(s, dt, GEMM_NT, p.h, f, p.w2_r, f, p.z, d, 0, M, d, f, e2));
  // out = dropout2(LN(z))                                                SubLayers.py:27
  return add_ln_fwd_any(s, ST_DTYPE_F32, mixed ? ST_DTYPE_F32 : dt, p.z, nullptr, a.ln_g, a.ln_b, a.out, nullptr, p.mean, p.rstd, M, d,
                        a.eps, mixed ? 0 : a.round_out, make_dropout(a.dropout_p, a.seed ^ kSeedMix2), nullptr, 0,
                        mixed ? a.out_h16 : nullptr);
}

int st_ffn_bwd(const st_ffn_bwd_args* bp, cudaStream_t s) {
  ST_REQUIRE(bp != nullptr, "st_ffn_bwd: null args");
  const st_ffn_bwd_args& b = *bp;
  const bool zeroed = b.grads_zeroed != 0;
  const st_ffn_args& a = b.f;
  ST_TRY(check_dtype(a.dtype, "st_ffn", true));
  FfnPlan p;
  ST_TRY(plan_ffn(a, p));
  const int M = static_cast<int>(a.rows), d = a.d_model, f = a.d_ff, dt = internal_dt(a.dtype);
  const bool mixed = is_mixed(a.dtype);
  ST_REQUIRE(a.ws != nullptr && a.ws_floats >= ffn_ws_floats(a.dtype, a.rows, d, f), "st_ffn_bwd: workspace too small");
  Carver w(a.ws, a.ws_floats);
  void* dz = w.take_act(dt, a.rows * d);
  void* dh = w.take_act(dt, a.rows * f);
  float* ws_amax = mixed ? w.take(64) : nullptr;     // gradient scale of this call (see st_mha_bwd)
  const float* amax = ws_amax;
  ST_CLEAR(b.dln_g, d);
  ST_CLEAR(b.dln_b, d);
  ST_CLEAR(b.db2, d);
  ST_CLEAR(b.db1, f);
  if (mixed) {
    if (b.dout_amax) amax = b.dout_amax;
    else ST_TRY(amax_abs(s, static_cast<const float*>(b.dout), a.rows * d, ws_amax));
    ST_TRY(add_ln_bwd_mixed(s, static_cast<const float*>(b.dout), p.z, p.mean, p.rstd, a.ln_g, dz, b.dln_g, b.dln_b, b.db2, M, d,
                            make_dropout(a.dropout_p, a.seed ^ kSeedMix2), amax, b.dx_amax));
  } else {
    ST_TRY(add_ln_bwd_any(s, dt, b.dout, p.z, p.mean, p.rstd, a.ln_g, dz, b.dln_g, b.dln_b, b.db2, M, d, 1,
                          make_dropout(a.dropout_p, a.seed ^ kSeedMix2)));
  }
  // dh = (dz W2) * [h > 0] * dropout1 scale
  GemmEpilogue e;
  e.aux = p.h; e.ldaux = f; e.aux_mode = 2; e.round_tf32 = 1;
  e.aux_scale = make_dropout(a.dropout_p, 0).scale;
  e.colsum = b.db1;   // db1 = column sums of dh, accumulated by the epilogue that produces dh
  e.unscale_amax = amax;
  Fork fk(s);                                   // the two weight gradients run next to the chain dz -> dh -> dx (st_host.h)
  cudaStream_t sw2 = fk.branch(0);
  ST_TRY(gemm_any(s, dt, GEMM_NN, dz, d, p.w2_r, f, dh, f, 1, M, f, d, e));
  cudaStream_t sw1 = fk.branch(1);
  ST_TRY(wgrad(sw2, dt, dz, d, p.h, f, b.dw2, M, d, f, zeroed, amax));
  ST_TRY(wgrad(sw1, dt, dh, f, p.x_r, d, b.dw1, M, f, d, zeroed, amax));
  GemmEpilogue ex;
  ex.aux = dz; ex.ldaux = d; ex.aux_mode = 1;
  ex.unscale_amax = amax;
  ex.amax_out = mixed ? b.dx_amax : nullptr;
  ST_TRY(gemm_any(s, dt, GEMM_NN, dh, f, p.w1_r, d, b.dx, d, mixed ? 0 : 1, M, d, f, ex));
  return fk.join();
}


// ------------------------------------------------------------------ decoder input: embedding + positional encoding
int st_embed_fwd(const int64_t* idx, const float* table, const float* pe, int64_t pe_rows, void* out, int64_t n, int d,
                 int vocab, int round_tf32, int dtype, cudaStream_t stream) {
  ST_TRY(check_dtype(dtype, "st_embed_fwd"));
  return embed_fwd(stream, idx, table, pe, pe_rows, out, n, d, vocab, round_tf32, dtype);
}
int st_embed_bwd(const int64_t* idx, const void* dout, float* dtable, int64_t n, int d, int vocab, int64_t padding_idx,
                 int zero_first, int dtype, cudaStream_t stream) {
  ST_TRY(check_dtype(dtype, "st_embed_bwd"));
  return embed_bwd(stream, idx, dout, dtable, n, d, vocab, padding_idx, zero_first, dtype);
}

int st_decode_self_attn(const float* qkv, float* k_cache, float* v_cache, int t, int n, int H, int dk, float* ctx,
                        int round_tf32, int32_t* slot_of, cudaStream_t stream) {
  return decode_self_attn(stream, qkv, k_cache, v_cache, t, n, H, dk, ctx, round_tf32, slot_of);
}

int st_beam_step(const float* logits, int64_t ld_logits, int B, int beam, int V, int first, int eos, int pad, float* scores,
                 uint8_t* done, int64_t* prev_k, int64_t* next_y, int64_t* parent, int64_t* tokens, cudaStream_t stream) {
  return beam_step(stream, logits, ld_logits, B, beam, V, first, eos, pad, scores, done, prev_k, next_y, parent, tokens);
}

// ------------------------------------------------------------------ encoder input front-end (Models.py:28-33,42-44)
namespace {
struct FrontPlan { float *x_r, *w_r, *h, *mean, *rstd; };
int plan_front(const st_frontend_args& a, FrontPlan& p) {
  ST_REQUIRE(a.rows > 0 && a.rows < (1ll << 31) && a.in_dim > 0 && (a.in_dim & 3) == 0 && a.d_model > 0 && a.T > 0,
             "st_frontend: bad shape (rows=%lld T=%d in_dim=%d must be a multiple of 4, d_model=%d)", (long long)a.rows, a.T,
             a.in_dim, a.d_model);
  Carver c(a.saved, a.saved_floats);
  p.x_r = c.take(a.rows * a.in_dim);
  p.w_r = c.take(static_cast<int64_t>(a.d_model) * a.in_dim);
  p.h = c.take(a.rows * a.d_model);
  p.mean = c.take(a.rows);
  p.rstd = c.take(a.rows);
  if (!c.ok()) {
    set_error("st_frontend: saved buffer too small (%lld floats given, %lld needed)", (long long)a.saved_floats, (long long)c.used);
    return ST_ERR_WORKSPACE;
  }
  return ST_OK;
}
constexpr uint64_t kSeedMixFront = 0x2545F4914F6CDD1Dull;
}  // namespace

int64_t st_frontend_saved_floats(int64_t rows, int in_dim, int d_model) {
  return pad64(rows * in_dim) + pad64(static_cast<int64_t>(d_model) * in_dim) + pad64(rows * d_model) + 2 * pad64(rows);
}
int64_t st_frontend_hidden_offset(int64_t rows, int in_dim, int d_model) {
  return pad64(rows * in_dim) + pad64(static_cast<int64_t>(d_model) * in_dim);
}
// backward scratch: dz, plus an fp32 copy of a 16-bit dout (the 80-wide input GEMMs of this layer stay on the TF32 path)
int64_t st_frontend_ws_floats(int64_t rows, int in_dim, int d_model) { (void)in_dim; return 2 * pad64(rows * d_model) + 64; }

int st_frontend_fwd(const st_frontend_args* ap, cudaStream_t s) {
  ST_REQUIRE(ap != nullptr, "st_frontend_fwd: null args");
  const st_frontend_args& a = *ap;
  FrontPlan p;
  ST_TRY(plan_front(a, p));
  const int M = static_cast<int>(a.rows), d = a.d_model, k = a.in_dim;
  ST_TRY(check_dtype(a.dtype, "st_frontend"));
  ST_TRY(round_tf32_2d(s, a.x, k, p.x_r, k, M, k));
  ST_TRY(round_tf32_2d(s, a.w, k, p.w_r, k, d, k));
  // h = Dropout(ReLU(Linear(x)))                                          Models.py:28-31
  GemmEpilogue e;
  e.bias = a.b; e.relu = 1;
  const DropoutCfg dc = make_dropout(a.dropout_p, a.seed ^ kSeedMixFront);
  e.drop_thresh = dc.thresh; e.drop_scale = dc.scale; e.drop_seed = dc.seed;
  ST_TRY(gemm_tf32(s, GEMM_NT, p.x_r, k, p.w_r, k, p.h, d, M, d, k, e));
  // out = LayerNorm(h) + positional encoding of the frame index          Models.py:32,42-44
  return add_ln_fwd_any(s, ST_DTYPE_F32, a.dtype, p.h, nullptr, a.ln_g, a.ln_b, a.out, nullptr, p.mean, p.rstd, M, d, a.eps,
                        a.round_out, DropoutCfg{}, a.pe, a.T);
}

int st_frontend_bwd(const st_frontend_bwd_args* bp, cudaStream_t s) {
  ST_REQUIRE(bp != nullptr, "st_frontend_bwd: null args");
  const st_frontend_bwd_args& b = *bp;
  const bool zeroed = b.grads_zeroed != 0;
  const st_frontend_args& a = b.f;
  FrontPlan p;
  ST_TRY(plan_front(a, p));
  const int M = static_cast<int>(a.rows), d = a.d_model, k = a.in_dim;
  ST_REQUIRE(a.ws != nullptr && a.ws_floats >= st_frontend_ws_floats(a.rows, k, d), "st_frontend_bwd: workspace too small");
  Carver w(a.ws, a.ws_floats);
  float* dz = w.take(a.rows * d);
  const float* dout = static_cast<const float*>(b.dout);
  if (a.dtype != ST_DTYPE_F32) {
    float* dout32 = w.take(a.rows * d);
    ST_TRY(cast_2d(s, b.dout, a.dtype, d, dout32, ST_DTYPE_F32, d, M, d));
    dout = dout32;
  }
  ST_CLEAR(b.dln_g, d);
  ST_CLEAR(b.dln_b, d);
  ST_CLEAR(b.db, d);
  // LayerNorm backward, then the gate of dropout(relu(.)) (h > 0 <=> kept and positive); db = column sums of the result
  ST_TRY(add_ln_bwd(s, dout, p.h, p.mean, p.rstd, a.ln_g, dz, b.dln_g, b.dln_b, b.db, M, d, 1, DropoutCfg{}, p.h,
                    make_dropout(a.dropout_p, 0).scale));
  ST_TRY(wgrad(s, ST_DTYPE_F32, dz, d, p.x_r, k, b.dw, M, d, k, zeroed));
  if (b.dx) {
    GemmEpilogue e;
    ST_TRY(gemm_tf32(s, GEMM_NN, dz, d, p.w_r, k, b.dx, k, M, k, d, e));
  }
  return ST_OK;
}

// ------------------------------------------------------------------ plain linear layer (vocabulary projection, Models.py:145,151)
namespace {
int64_t pad4(int64_t n) { return (n + 3) & ~int64_t(3); }
int64_t pad8(int64_t n) { return (n + 7) & ~int64_t(7); }
int64_t lin_ldr(int dt, int n) { return dt == ST_DTYPE_F32 ? pad4(n) : pad8(n); }   // row stride of the operand-precision dy copy
struct LinPlan { const void* x_r; void* w_r; };
int plan_linear(const st_linear_args& a, LinPlan& p) {
  ST_TRY(check_dtype(a.dtype, "st_linear"));
  const int al = a.dtype == ST_DTYPE_F32 ? 4 : 8;
  ST_REQUIRE(a.rows > 0 && a.rows < (1ll << 31) && a.in_dim > 0 && (a.in_dim % al) == 0 && a.out_dim > 0,
             "st_linear: bad shape (rows=%lld in_dim=%d must be a multiple of %d, out_dim=%d)", (long long)a.rows, a.in_dim, al, a.out_dim);
  Carver c(a.saved, a.saved_floats);
  p.x_r = (a.dtype != ST_DTYPE_F32 || a.x_is_tf32) ? a.x : c.take(a.rows * a.in_dim);
  p.w_r = c.take_act(a.dtype, pad4(a.out_dim) * a.in_dim);   // rows up to a multiple of 4 (zero-filled): see st_linear_fwd
  if (!c.ok()) {
    set_error("st_linear: saved buffer too small (%lld floats given, %lld needed)", (long long)a.saved_floats, (long long)c.used);
    return ST_ERR_WORKSPACE;
  }
  return ST_OK;
}
}  // namespace

int64_t st_linear_saved_floats(int64_t rows, int in_dim, int out_dim, int x_is_tf32) {
  return (x_is_tf32 ? 0 : pad64(rows * in_dim)) + pad64(pad4(out_dim) * in_dim);
}
int64_t st_linear_ws_floats(int64_t rows, int in_dim, int out_dim) { (void)in_dim; return pad64(rows * pad4(out_dim)) + 64; }
int64_t st_linear_saved_floats_dt(int dtype, int64_t rows, int in_dim, int out_dim, int x_is_tf32) {
  return ((dtype != ST_DTYPE_F32 || x_is_tf32) ? 0 : pad64(rows * in_dim)) + pad64a(dtype, pad4(out_dim) * in_dim);
}
int64_t st_linear_ws_floats_dt(int dtype, int64_t rows, int in_dim, int out_dim) {
  (void)in_dim;
  return pad64a(dtype, rows * lin_ldr(dtype, out_dim)) + 64;
}

int st_linear_fwd(const st_linear_args* ap, cudaStream_t s) {
  ST_REQUIRE(ap != nullptr, "st_linear_fwd: null args");
  const st_linear_args& a = *ap;
  LinPlan p;
  ST_TRY(plan_linear(a, p));
  ST_REQUIRE(a.ldy >= a.out_dim, "st_linear_fwd: ldy (%lld) < out_dim (%d)", (long long)a.ldy, a.out_dim);
  const int M = static_cast<int>(a.rows), k = a.in_dim, n = a.out_dim, dt = a.dtype;
  if (dt == ST_DTYPE_F32 && !a.x_is_tf32)
    ST_TRY(round_tf32_2d(s, static_cast<const float*>(a.x), k, const_cast<float*>(static_cast<const float*>(p.x_r)), k, M, k));
  ST_TRY(weight_copy(s, dt, a.w, p.w_r, n, k));
  GemmEpilogue e;
  e.bias = a.b;
  // A width that is not a multiple of 4 (V = 4337) would force the scalar epilogue.  Without a bias the GEMM can run on
  // the width rounded up instead: the extra weight rows are zeros and land in the padding columns of the row-padded y.
  int n_eff = n;
  if (!a.b && pad4(n) != n && a.ldy >= pad4(n)) {
    n_eff = static_cast<int>(pad4(n));
    ST_CHECK_CUDA(cudaMemsetAsync(at(p.w_r, dt, static_cast<int64_t>(n) * k), 0, static_cast<size_t>(n_eff - n) * k * esz(dt), s));
  }
  return gemm_any(s, dt, GEMM_NT, p.x_r, k, p.w_r, k, a.y, a.ldy, 0, M, n_eff, k, e);   // logits stay fp32
}

int st_linear_bwd(const st_linear_bwd_args* bp, cudaStream_t s) {
  ST_REQUIRE(bp != nullptr, "st_linear_bwd: null args");
  const st_linear_bwd_args& b = *bp;
  const bool zeroed = b.grads_zeroed != 0;
  const st_linear_args& a = b.f;
  LinPlan p;
  ST_TRY(plan_linear(a, p));
  const int M = static_cast<int>(a.rows), k = a.in_dim, n = a.out_dim, dt = a.dtype;
  ST_REQUIRE(a.ws != nullptr && a.ws_floats >= st_linear_ws_floats_dt(dt, a.rows, k, n), "st_linear_bwd: workspace too small");
  ST_REQUIRE(b.lddy >= n, "st_linear_bwd: lddy (%lld) < out_dim (%d)", (long long)b.lddy, n);
  Carver w(a.ws, a.ws_floats);
  const int64_t ldr = lin_ldr(dt, n);
  void* dy_r = w.take_act(dt, a.rows * ldr);
  if (ldr != n)   // the ragged tail columns are read by vector column sums / TMA boxes: keep them finite
    ST_CHECK_CUDA(cudaMemsetAsync(dy_r, 0, static_cast<size_t>(a.rows) * ldr * esz(dt), s));
  if (dt == ST_DTYPE_F32) ST_TRY(round_tf32_2d(s, b.dy, b.lddy, static_cast<float*>(dy_r), ldr, M, n));
  else ST_TRY(cast_2d(s, b.dy, ST_DTYPE_F32, b.lddy, dy_r, dt, ldr, M, n));
  if (b.dx) {
    GemmEpilogue e;
    ST_TRY(gemm_any(s, dt, GEMM_NN, dy_r, ldr, p.w_r, k, b.dx, k, 1, M, k, n, e));
  }
  if (b.dw) ST_TRY(wgrad(s, dt, dy_r, ldr, p.x_r, k, b.dw, M, n, k, zeroed));
  if (b.db && a.b) {
    ST_CLEAR(b.db, n);
    ST_TRY(colsum_add_any(s, dt, dy_r, ldr, M, n, b.db));
  }
  return ST_OK;
}

// gradient alone, from the workspace (log-sum-exp, alpha, beta) and nll a previous st_ctc_fwd_bwd call left behind
int st_ctc_grad(const float* logits, int64_t ld_logits, const int64_t* targets, int64_t ld_targets,
                const int64_t* input_lengths, const int64_t* target_lengths, int blank, int B, int T, int V, int L_max,
                const float* nll, const float* scale, float* grad, int64_t ld_grad, float* ws, int64_t ws_floats,
                cudaStream_t stream) {
  return ctc_fwd_bwd(stream, logits, ld_logits, targets, ld_targets, input_lengths, target_lengths, blank, B, T, V, L_max,
                     const_cast<float*>(nll), scale, grad, ld_grad, ws, ws_floats, 1);
}

// ------------------------------------------------------------------ CTC head of the joint CTC / attention loss
int64_t st_ctc_ws_floats(int B, int T, int L_max) { return ctc_ws_floats(B, T, 2 * L_max + 1); }
int st_ctc_fwd_bwd(const float* logits, int64_t ld_logits, const int64_t* targets, int64_t ld_targets,
                   const int64_t* input_lengths, const int64_t* target_lengths, int blank, int B, int T, int V, int L_max,
                   float* nll, const float* scale, float* grad, int64_t ld_grad, float* ws, int64_t ws_floats,
                   cudaStream_t stream) {
  return ctc_fwd_bwd(stream, logits, ld_logits, targets, ld_targets, input_lengths, target_lengths, blank, B, T, V, L_max, nll,
                     scale, grad, ld_grad, ws, ws_floats);
}

}  // extern "C"
