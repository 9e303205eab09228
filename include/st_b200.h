/* st_b200.h — C ABI of libst_b200.so, the sm_100a (B200) implementation of the Speech-Transformer
 * hot path.  This is the drop-in boundary: plain pointers, sizes and a cudaStream_t; no torch or
 * C++ types.  Every pointer is a DEVICE pointer unless stated otherwise; every tensor is fp32,
 * row-major and contiguous unless a leading dimension / stride argument says otherwise.  Calls
 * enqueue work on `stream` and return immediately.
 *
 * Return value: 0 on success, negative on error (ST_ERR_*); st_last_error() describes the failure.
 *
 * The reference is pure PyTorch (no FFI of its own); each entry point below replaces the eager
 * ATen call sequence at the cited reference lines (paths relative to the reference repo root).
 */
#ifndef ST_B200_H_
#define ST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define ST_OK 0
#define ST_ERR_INVALID (-1)
#define ST_ERR_CUDA (-2)
#define ST_ERR_DEVICE (-3)
#define ST_ERR_WORKSPACE (-4)

/* Element type of ACTIVATION tensors (module inputs / outputs and their gradients).  F32 is the fp32 / TF32 path every
 * entry point defaults to.  F16 / BF16 select 16-bit operands on the tensor cores (tcgen05 kind::f16, fp32 accumulate;
 * LayerNorm statistics, softmax statistics, parameters, parameter gradients and optimizer state stay fp32):
 * BASELINE.json configs[2].  Arguments declared `void*` below are of this type; `float*` ones are always fp32.   */
#define ST_DTYPE_F32 0
#define ST_DTYPE_F16 1
#define ST_DTYPE_BF16 2
/* Composite operators only (st_mha_*, st_ffn_*): fp32 activations at the boundary exactly as with ST_DTYPE_F32 — inputs,
 * outputs and all gradients are fp32 tensors (inputs_tf32 / x_is_tf32 / round_out are ignored) — but every INTERNAL
 * tensor-core operand (the inputs' operand copies, Q/K/V, context, FFN hidden, all internal gradients, the weight copies)
 * is fp16, which carries the same 10-bit mantissa as TF32 at twice the MMA rate and half the bytes.  Internal gradients are
 * scaled by a power of two derived on the device from max|dout| of each backward call, so gradients of any magnitude are
 * safe; forward activations must lie in fp16's range.  d_k must be 64, the residual must alias q_in, k_in or v_in, and the
 * optional *_tf32 weight copies are then fp16.  Sizes: st_*_saved_floats_dt / st_*_ws_floats_dt with this code.          */
#define ST_DTYPE_F32_H16 3

/* ---- library ------------------------------------------------------------------------------- */
int st_version(void);                       /* 10000*major + 100*minor + patch */
const char* st_last_error(void);            /* host string, valid until the next failing call */
int st_device_check(int device);            /* ST_OK iff `device` is compute capability 10.x */
int st_set_option(const char* name, int v); /* debug/tuning switches, see st_host.cu */
int64_t st_launch_count(void);              /* kernels this library has launched in this process so far */
/* Optional per-kernel-class timing with CUDA events on the launching stream (used by bench.py's roofline).
 * `work` is the algorithmic FLOP count (tensor-core classes) or byte count (HBM classes) of the launches.  */
int st_profile_enable(int on);
int st_profile_reset(void);
int st_profile_dump(const char* path /* host */);   /* CSV of every recorded launch: index,class,work,ms */
int st_profile_classes(void);
const char* st_profile_class_name(int cls);
int st_profile_read(int cls, double* ms /* host */, double* work /* host */, int64_t* launches /* host */);
/* debug: clock64() timeline of one CTA of the attention dK/dV kernel (option "attn_trace"); returns the slot count */
int st_debug_read_trace(uint64_t* host_out /* host */, int n);
int st_debug_read_fwd_trace(uint64_t* host_out /* host */, int n);   /* same for thread 0 of the forward kernel: [tile][8 events] */
/* debug: cycles per tcgen05.mma (M=128, N=n, K=8, kind::tf32) issued back to back by one CTA.
 * variant bit 0: A operand from TMEM (.ts) instead of shared memory; bit 1: B operand MN-major instead of K-major. */
int st_debug_mma_bench(int variant, int n, int iters, double* clk_per_mma /* host */);
int st_selftest_count(void);
int st_selftest(int which, double* rel_err_out /* host */); /* tcgen05 building-block self tests */

/* ---- (C) residual + LayerNorm ---------------------------------------------------------------
 * out = dropout(LayerNorm(a + b) * gamma + beta)           Attention.py:62,94  SubLayers.py:18,27
 * b, z_out, mean_out, rstd_out may be NULL.  z_out receives a + b (needed by the backward).
 * dropout_p == 0 disables dropout; round_tf32 != 0 rounds `out` to TF32 (it feeds a tensor-core op).
 */
int st_add_ln_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out, float* z_out,
                  float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_tf32, float dropout_p,
                  uint64_t seed, cudaStream_t stream);
/* dz = d(loss)/d(a+b); dgamma/dbeta/dzsum (each [d], may be NULL) are ACCUMULATED into (caller zeroes).
 * dzsum = column sums of dz = bias gradient of the linear layer that produced the LN input. */
int st_add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                  float* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, int round_tf32,
                  float dropout_p, uint64_t seed, cudaStream_t stream);

/* ---- (D) label-smoothed / soft-target cross entropy -------------------------------------------
 * LabelSmoothingLoss.forward (Loss.py:28-39): q_i = one_hot with q_i[target_i] = confidence, rows with
 * target_i == padding_idx zeroed when padding_idx >= 0; then CrossEntropyLoss.forward (Loss.py:50-73):
 * loss = sum_i sum_c weight_c q_ic (-log_softmax(logits_i)_c), divided by N when size_average.
 * grad (may be NULL) receives d(loss)/d(logits).  row_loss is an N-float workspace.            */
int st_lsce_fwd_bwd(const float* logits, int64_t ld_logits, const int64_t* target, const float* one_hot,
                    const float* weight, float confidence, int64_t padding_idx, int size_average, int64_t N, int V,
                    float* row_loss, float* loss, float* grad, int64_t ld_grad, cudaStream_t stream);
/* CrossEntropyLoss.forward with a dense soft target q (N, V)                       Loss.py:50-73 */
int st_softce_fwd_bwd(const float* logits, int64_t ld_logits, const float* q, const float* weight, int size_average,
                      int64_t N, int V, float* row_loss, float* loss, float* grad, int64_t ld_grad,
                      cudaStream_t stream);

/* ---- TF32 tensor-core linear algebra ------------------------------------------------------------
 * dst = round_to_tf32(src), 2-D strided copy (vectorised when cols, lds, ldd are multiples of 4). */
int st_round_tf32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, cudaStream_t stream);
/* out[c] += sum_r x[r, c]                                                                        */
int st_colsum_add(const float* x, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t stream);

/* General TF32 GEMM on tcgen05 (operands must already be TF32-representable for exact TF32 semantics):
 *   mode 0 (NT): C[M,N] = A[M,K] * B[N,K]^T     nn.Linear forward      Attention.py:74-76,92 SubLayers.py:25-26
 *   mode 1 (NN): C[M,N] = A[M,K] * B[K,N]       its input gradient
 *   mode 2 (TN): C[M,N] = A[K,M]^T * B[K,N]     its weight gradient (k_splits > 1: C must be zeroed, atomics)
 * epilogue: (+bias[N]) -> (relu) -> (dropout) -> aux_mode 1: += aux[M,ldaux] | 2: zero where aux <= 0
 *           -> (round to TF32).                                                                   */
typedef struct {
  const float* bias;
  const float* aux;
  int64_t ldaux;
  int aux_mode;
  int relu;
  int round_tf32;
  int k_splits;
  float dropout_p;
  uint64_t seed;
} st_gemm_epilogue;
int st_gemm(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
            int K, const st_gemm_epilogue* ep /* host, may be NULL */, cudaStream_t stream);
/* The same GEMM for any operand type: A, B and ep->aux are of `dtype` (ST_DTYPE_F32 = st_gemm; F16 / BF16 = tcgen05
 * kind::f16, fp32 accumulate, leading dimensions in elements and multiples of 8); C is fp32, or `dtype` when c_lp != 0. */
int st_gemm_dt(int dtype, int mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int c_lp,
               int M, int N, int K, const st_gemm_epilogue* ep /* host, may be NULL */, cudaStream_t stream);
/* dst = (dst_dtype)(src * scale): 2-D strided element-type conversion; one of the two types is fp32.              */
int st_cast(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd, int64_t rows, int cols,
            float scale, cudaStream_t stream);

/* ---- (A) attention core: softmax(mask(Q K^T / sqrt(d_k))) V for all heads ------------------------
 * Attention.py:78-90 (split heads, scores, masked_fill_(-inf), softmax, dropout, P·V, merge heads)
 * and Attention.py:27-35 (ScaledDotProductAttention = the H == 1 case).
 * q: (B*Lq, >= H*dk) with leading dimension ldq; head h occupies columns [h*dk, (h+1)*dk).  Same for
 * k, v (B*Lk rows), ctx (B*Lq rows, merged heads).  Operands must be TF32-representable.
 * mask: NULL or bytes, nonzero = masked, element (b,i,j) at mask[b*ms_b + i*ms_q + j*ms_k] (stride 0
 * broadcasts, as in the expanded view built by Utils.py:53-54).  A fully masked row yields NaN, like
 * the reference.  lse: (B,H,Lq) log-sum-exp of the scaled masked scores (saved for backward).
 * attn: NULL, or (B,H,Lq,Lk) to receive the post-dropout probabilities the module returns.        */
typedef struct {
  int B, H, Lq, Lk, dk;
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const uint8_t* mask; int64_t ms_b, ms_q, ms_k;
  float dropout_p; uint64_t seed;
  void* ctx; int64_t ldctx;
  float* lse;
  float* attn;
  int dtype;              /* ST_DTYPE_*: element type of q, k, v, ctx (and dctx, dq, dk, dv); 16-bit types need dk = 64 */
  /* Length-aware masking — no mask tensor is built or scanned (Utils.py:41-70, Models.py:46,89-97): keys j >= k_len[b]
   * are masked (k_len: NULL or (B,) int64) and, when causal != 0, keys j > i.  ORed with `mask` when both are given. */
  const int64_t* k_len;
  int causal;
} st_attn_args;
int st_attn_fwd(const st_attn_args* a /* host */, cudaStream_t stream);

typedef struct {
  st_attn_args f;       /* the forward problem: q,k,v,mask,dropout,ctx,lse exactly as given to / produced by st_attn_fwd */
  const void* dctx; int64_t lddctx;    /* gradient w.r.t. ctx (TF32-representable, or f.dtype) */
  float* delta;                        /* (B,H,Lq) workspace */
  void* dq; int64_t lddq;              /* gradients, same indexing as q/k/v; rounded to TF32 (or f.dtype) */
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
} st_attn_bwd_args;
int st_attn_bwd(const st_attn_bwd_args* a /* host */, cudaStream_t stream);

/* ---- composite: MultiHeadAttention.forward / backward -------------------------------------------
 * Attention.py:64-96:  LN(Linear_o(attention(Linear_q(q), Linear_k(k), Linear_v(v), mask)) + residual)
 * `residual` is v in the reference (Attention.py:94); the caller passes the tensor to add.
 * q_in/k_in/v_in: (B*Lq, d), (B*Lk, d), (B*Lk, d).  Weights are [out, in] like nn.Linear.
 * `saved` is a caller-allocated float buffer of st_mha_saved_floats() elements that forward fills and
 * backward reads; `ws` is scratch of st_mha_ws_floats() elements (contents undefined afterwards).  */
typedef struct {
  int B, Lq, Lk, H, d_model, dk;
  const void* q_in; const void* k_in; const void* v_in;
  const void* residual;
  const float* wq; const float* bq; const float* wk; const float* bk;
  const float* wv; const float* bv; const float* wo; const float* bo;
  const float* ln_g; const float* ln_b;
  const uint8_t* mask; int64_t ms_b, ms_q, ms_k;
  float eps; float dropout_p; uint64_t seed;
  int inputs_tf32;      /* q_in/k_in/v_in are already TF32-representable: skip the rounding copies */
  int round_out;        /* round the module output to TF32 (it feeds the next layer's GEMM) */
  void* out;            /* (B*Lq, d) */
  float* attn;          /* NULL or (B,H,Lq,Lk) */
  float* saved; int64_t saved_floats;
  float* ws; int64_t ws_floats;
  /* Optional operand-precision copies of the weights (TF32-rounded fp32, or `dtype` when that is 16-bit), maintained by
   * the caller (st_adam_step writes them): when all four are given, [wq; wk; wv] are adjacent in memory
   * (wk_tf32 == wq_tf32 + d*d elements, ...) and so are the biases (bk == bq + d, bv == bk + d), the per-call rounding /
   * packing passes are skipped.  NULL = round / convert internally. */
  const void* wq_tf32; const void* wk_tf32; const void* wv_tf32; const void* wo_tf32;
  int dtype;            /* ST_DTYPE_*: element type of q_in, k_in, v_in, residual, out (and of dout, dq_in, dk_in, dv_in,
                           dresidual in the backward); parameters and their gradients are always fp32 */
  const int64_t* k_len; /* see st_attn_args: key lengths (B,) or NULL */
  int causal;
  /* ST_DTYPE_F32_H16 only — chaining operators without conversion passes.  When inputs_tf32 != 0 the caller supplies the
   * fp16 copies of the inputs (the producer's out_h16, or st_cast): q_h16, plus k_h16 / v_h16 where k_in / v_in are other
   * tensors; they must stay valid until the backward call, and `saved` does not hold copies of its own.  out_h16: NULL or
   * (B*Lq, d) fp16, receives a copy of `out` for the next operator.                                                     */
  const void* q_h16; const void* k_h16; const void* v_h16;
  void* out_h16;
} st_mha_args;
int64_t st_mha_saved_floats(int B, int Lq, int Lk, int H, int d_model, int same_qkv, int same_kv, int inputs_tf32);
int64_t st_mha_ws_floats(int B, int Lq, int Lk, int H, int d_model);
int64_t st_mha_saved_floats_dt(int dtype, int B, int Lq, int Lk, int H, int d_model, int same_qkv, int same_kv, int inputs_tf32);
int64_t st_mha_ws_floats_dt(int dtype, int B, int Lq, int Lk, int H, int d_model);
int st_mha_fwd(const st_mha_args* a /* host */, cudaStream_t stream);

typedef struct {
  st_mha_args f;        /* same problem description as forward (out/attn unused) */
  const void* dout;     /* (B*Lq, d) */
  void* dq_in; void* dk_in; void* dv_in;      /* (rows, d); when inputs alias (self-attention) pass the same
                                                 pointer and the sum is written once */
  void* dresidual;      /* (B*Lq, d) gradient of the residual input; needed only when the residual is none of the inputs
                           (otherwise it is added to the FIRST of dq_in, dk_in, dv_in whose input is the residual) */
  float* dwq; float* dbq; float* dwk; float* dbk; float* dwv; float* dbv; float* dwo; float* dbo;
  float* dln_g; float* dln_b;                 /* all parameter gradients are OVERWRITTEN */
  int grads_zeroed;     /* the parameter-gradient buffers already hold zeros (slices of a gradient buffer the caller cleared in
                           one pass, parallel.FlatParams.zero_grad): skip the per-tensor clears (~200 memsets per step) */
  /* ST_DTYPE_F32_H16 only: dout_amax = NULL or a device scalar holding max|dout| (the dq_amax / dx_amax a later operator's
   * backward produced for exactly this tensor) — saves the pass that measures it; dq_amax = NULL or a device scalar that
   * receives max|dq_in|.                                                                                                */
  const float* dout_amax;
  float* dq_amax;
} st_mha_bwd_args;
int st_mha_bwd(const st_mha_bwd_args* a /* host */, cudaStream_t stream);

/* ---- composite: PositionwiseFeedForward.forward / backward ---------------------------------------
 * SubLayers.py:24-28:  dropout2(LN(x + fc2(dropout1(relu(fc1(x))))))                               */
typedef struct {
  int64_t rows; int d_model, d_ff;
  const void* x;
  const float* w1; const float* b1; const float* w2; const float* b2;
  const float* ln_g; const float* ln_b;
  float eps; float dropout_p; uint64_t seed;
  int x_is_tf32;        /* x is already TF32-representable (produced by this library with round_out) */
  int round_out;
  void* out;
  float* saved; int64_t saved_floats;
  float* ws; int64_t ws_floats;
  const void* w1_tf32; const void* w2_tf32;   /* optional operand-precision weights (see st_mha_args); NULL = round internally */
  int dtype;            /* ST_DTYPE_*: element type of x, out (and dout, dx) */
  const void* x_h16;    /* ST_DTYPE_F32_H16 with x_is_tf32 != 0: the caller's fp16 copy of x (see st_mha_args.q_h16) */
  void* out_h16;        /* ST_DTYPE_F32_H16: NULL or (rows, d_model) fp16 copy of `out` for the next operator */
} st_ffn_args;
int64_t st_ffn_saved_floats(int64_t rows, int d_model, int d_ff, int x_is_tf32);
int64_t st_ffn_ws_floats(int64_t rows, int d_model, int d_ff);
int64_t st_ffn_saved_floats_dt(int dtype, int64_t rows, int d_model, int d_ff, int x_is_tf32);
int64_t st_ffn_hidden_offset_dt(int dtype, int64_t rows, int d_model, int d_ff, int x_is_tf32);
int64_t st_ffn_ws_floats_dt(int dtype, int64_t rows, int d_model, int d_ff);
/* float offset inside `saved` of the hidden activation h = dropout1(relu(fc1(x))), shape (rows, d_ff) — test hook */
int64_t st_ffn_hidden_offset(int64_t rows, int d_model, int d_ff, int x_is_tf32);
int st_ffn_fwd(const st_ffn_args* a /* host */, cudaStream_t stream);

typedef struct {
  st_ffn_args f;
  const void* dout;
  void* dx;
  float* dw1; float* db1; float* dw2; float* db2; float* dln_g; float* dln_b;   /* OVERWRITTEN */
  int grads_zeroed;     /* see st_mha_bwd_args */
  const float* dout_amax;   /* see st_mha_bwd_args */
  float* dx_amax;
} st_ffn_bwd_args;
int st_ffn_bwd(const st_ffn_bwd_args* a /* host */, cudaStream_t stream);

/* ---- callers either side of the path ------------------------------------------------
 * Decoder input (Models.py:84-87, Embedding.py:21-29): out[i] = table[idx[i]] + pe[i mod pe_rows]
 * (pe may be NULL).  idx are int64 token ids in [0, vocab).  round_tf32 rounds the result (it feeds a GEMM). */
int st_embed_fwd(const int64_t* idx, const float* table, const float* pe, int64_t pe_rows, void* out, int64_t n, int d,
                 int vocab, int round_tf32, int dtype /* of out */, cudaStream_t stream);
/* dtable[idx[i]] += dout[i] for idx[i] != padding_idx (nn.Embedding(padding_idx=PAD) semantics, Models.py:73);
 * zero_first clears the (vocab, d) gradient table before accumulating.                                     */
int st_embed_bwd(const int64_t* idx, const void* dout, float* dtable, int64_t n, int d, int vocab, int64_t padding_idx,
                 int zero_first, int dtype /* of dout */, cudaStream_t stream);

/* Incremental-decode self-attention (Decode.py:48-179 decodes the full prefix every step; this is the K/V-reuse form):
 * qkv (n, 3*H*dk) = [q | k | v] projections of the ONE new position of each of the n hypotheses.  Appends k, v as
 * row t of the time-major caches (L_max, n, H*dk) and writes ctx (n, H*dk) = softmax(q K[0..t]^T / sqrt(dk)) V[0..t]
 * per head (Attention.py:78-90 with Lq = 1; no mask: every cached position is in the past).
 * slot_of: NULL, or a time-major (L_max, n) int32 table: position j of hypothesis i's history lives in cache slot
 * slot_of[j*n + i].  The call records slot_of[t*n + i] = i for the rows it appends; beam search re-parents hypotheses by
 * permuting the table (slot_of[:t] <- slot_of[:t][:, parent]) instead of the caches themselves.                    */
int st_decode_self_attn(const float* qkv, float* k_cache, float* v_cache, int t, int n, int H, int dk, float* ctx,
                        int round_tf32, int32_t* slot_of, cudaStream_t stream);

/* One position of beam-search bookkeeping on the device (Beam.advance, Beam.py:43-74, as driven by Decode.py:120-160) for B
 * utterances x `beam` live hypotheses.  logits: (B*beam, >= V) scores of the next symbol, row stride ld_logits.  Per
 * utterance: cand[k][v] = scores[k] + log_softmax(logits[k])[v] (first != 0: only hypothesis 0 is expanded — all beams are
 * identical at the first position, Beam.py:49-52); the `beam` best candidates, best first (ties: lower k*V + v first), give
 * scores (B, beam) in/out, prev_k (B, beam) integer back-pointers k (Beam.py:66), next_y (B, beam) symbols v, parent
 * (B*beam) = b*beam + prev_k (cache re-parenting for st_decode_self_attn's caches), tokens (B*beam) = next_y.  done (B)
 * bytes in/out: a finished utterance is frozen (scores kept, prev_k = identity, next_y = pad); an utterance becomes
 * finished when its best hypothesis emits eos (Beam.py:70-72).  beam <= 32.                                      */
int st_beam_step(const float* logits, int64_t ld_logits, int B, int beam, int V, int first, int eos, int pad, float* scores,
                 uint8_t* done, int64_t* prev_k, int64_t* next_y, int64_t* parent, int64_t* tokens, cudaStream_t stream);

/* Encoder input front-end (Models.py:28-33,42-44):
 *   out = LayerNorm(Dropout(ReLU(x W^T + b))) * gamma + beta + pe[frame index]
 * x: (rows, in_dim) with rows = B*T and frame index = row mod T; in_dim a multiple of 4 (80 for fbank).
 * pe: (>= T, d_model) sinusoid table or NULL.  saved / ws as for the other composites.                     */
typedef struct {
  int64_t rows; int T; int in_dim, d_model;
  const float* x;
  const float* w; const float* b;
  const float* ln_g; const float* ln_b;
  const float* pe;
  float eps; float dropout_p; uint64_t seed;
  int round_out;
  void* out;
  float* saved; int64_t saved_floats;
  float* ws; int64_t ws_floats;
  int dtype;                       /* ST_DTYPE_*: element type of out (and dout); x, dx and the 80-wide GEMMs stay fp32 / TF32 */
} st_frontend_args;
int64_t st_frontend_saved_floats(int64_t rows, int in_dim, int d_model);
int64_t st_frontend_ws_floats(int64_t rows, int in_dim, int d_model);
/* float offset inside `saved` of h = Dropout(ReLU(Linear(x))), shape (rows, d_model) — test hook (ReLU gate pattern) */
int64_t st_frontend_hidden_offset(int64_t rows, int in_dim, int d_model);
int st_frontend_fwd(const st_frontend_args* a /* host */, cudaStream_t stream);
typedef struct {
  st_frontend_args f;
  const void* dout;
  float* dx;                       /* may be NULL: acoustic features need no gradient */
  float* dw; float* db; float* dln_g; float* dln_b;   /* OVERWRITTEN */
  int grads_zeroed;                /* see st_mha_bwd_args */
} st_frontend_bwd_args;
int st_frontend_bwd(const st_frontend_bwd_args* a /* host */, cudaStream_t stream);

/* Plain linear layer y = x W^T (+ b) on the TF32 tensor cores — the vocabulary projection tgt_word_proj
 * (Models.py:145,151).  out_dim may be any size (V = 4337); y has leading dimension ldy >= out_dim.
 * Backward accepts dy with any leading dimension lddy >= out_dim; dx / dw / db may each be NULL.           */
typedef struct {
  int64_t rows; int in_dim, out_dim;
  const void* x; int x_is_tf32;
  const float* w; const float* b;
  float* y; int64_t ldy;
  float* saved; int64_t saved_floats;
  float* ws; int64_t ws_floats;
  int dtype;                       /* ST_DTYPE_*: element type of x (and dx); y and dy (logits) are always fp32 */
} st_linear_args;
int64_t st_linear_saved_floats(int64_t rows, int in_dim, int out_dim, int x_is_tf32);
int64_t st_linear_ws_floats(int64_t rows, int in_dim, int out_dim);
int64_t st_linear_saved_floats_dt(int dtype, int64_t rows, int in_dim, int out_dim, int x_is_tf32);
int64_t st_linear_ws_floats_dt(int dtype, int64_t rows, int in_dim, int out_dim);
int st_linear_fwd(const st_linear_args* a /* host */, cudaStream_t stream);
typedef struct {
  st_linear_args f;
  const float* dy; int64_t lddy;
  void* dx; float* dw; float* db;                     /* OVERWRITTEN */
  int grads_zeroed;                                   /* see st_mha_bwd_args */
} st_linear_bwd_args;
int st_linear_bwd(const st_linear_bwd_args* a /* host */, cudaStream_t stream);

/* ---- CTC head of the joint CTC / attention objective (train_attn_and_ctc.py is empty in the
 * reference, semantics are those of torch.nn.functional.ctc_loss) ------------------------------------------------
 * logits: (B, T, V) unnormalised scores, row (b, t) at logits + (b*T + t) * ld_logits; the log-softmax is taken
 * inside.  targets: (B, >= L_max) int64 labels (no blanks), row stride ld_targets; input_lengths / target_lengths:
 * (B,) int64.  nll: (B,) receives -log p(targets_b | logits_b) (+inf when no alignment exists).  grad: NULL, or
 * (B, T, ld_grad) receiving scale[b] * d nll[b] / d logits (scale NULL = 1; frames beyond input_lengths[b] and
 * infeasible utterances get 0).  ws: st_ctc_ws_floats(B, T, L_max) floats of scratch.  L_max <= 511.            */
int64_t st_ctc_ws_floats(int B, int T, int L_max);
int st_ctc_fwd_bwd(const float* logits, int64_t ld_logits, const int64_t* targets, int64_t ld_targets,
                   const int64_t* input_lengths, const int64_t* target_lengths, int blank, int B, int T, int V, int L_max,
                   float* nll, const float* scale, float* grad, int64_t ld_grad, float* ws, int64_t ws_floats,
                   cudaStream_t stream);
/* The gradient alone, from the `ws` contents and `nll` a previous st_ctc_fwd_bwd call (same arguments) left behind:
 * lets a caller run the forward without knowing the upstream gradient and apply `scale` in its backward.      */
int st_ctc_grad(const float* logits, int64_t ld_logits, const int64_t* targets, int64_t ld_targets,
                const int64_t* input_lengths, const int64_t* target_lengths, int blank, int B, int T, int V, int L_max,
                const float* nll, const float* scale, float* grad, int64_t ld_grad, float* ws, int64_t ws_floats,
                cudaStream_t stream);

/* ---- flat-buffer optimizer step (train.py:45-46, Optim.py:9-14,36-45) --------------------------------
 * st_sumsq: *out += sum(x[i]^2) (caller zeroes `out`, a device float).
 * st_adam_step: g' = grad * grad_scale * min(1, max_grad_norm / (sqrt(*norm_ws) * grad_scale + 1e-6))
 * (clip_grad_norm_ semantics; skipped when norm_ws is NULL or max_grad_norm <= 0), then torch.optim.Adam's
 * update with bias correction for `step` (1-based).  n must be a multiple of 4.  A non-finite *norm_ws (an fp16
 * activation gradient overflowed under loss scaling) skips the update: nothing is written.             */
int st_sumsq(const float* x, int64_t n, float* out, cudaStream_t stream);
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t n;
  float lr, beta1, beta2, eps;
  int step;
  float max_grad_norm, grad_scale;
  const float* norm_ws;
  void* param_tf32;     /* optional (n elements): receives the operand-precision copy of the updated parameters for the
                           next step's GEMMs: round_to_tf32 (fp32) or, with twin_dtype F16 / BF16, the 16-bit conversion */
  int twin_dtype;       /* ST_DTYPE_* of param_tf32 */
} st_adam_args;
int st_adam_step(const st_adam_args* a /* host */, cudaStream_t stream);

/* ---- data-parallel gradient exchange (train_multi.py:20,128,161-163,176-177: Horovod all-reduce of every gradient,
 * broadcast of the initial state) -------------------------------------------------------------------------------
 * One process per GPU; all gradients live in ONE flat fp32 buffer (what st_*_bwd write into and st_adam_step reads), so
 * the exchange is one in-place ncclAllReduce(sum) over NVLink / NVSwitch; the average is st_adam_args.grad_scale = 1/world.
 * NCCL is resolved with dlopen at the first call (no link-time dependency).
 *   st_allreduce_unique_id : rank 0 fills `id_out` (st_allreduce_id_bytes() = 128 host bytes) and hands it to every rank out
 *                            of band (MPI, a file, torch.distributed ...)
 *   st_allreduce_init      : every rank, with its CUDA device current; collective
 *   st_allreduce_run       : buf[0..n) <- sum over ranks, enqueued on `stream`; st_allreduce_broadcast: buf <- root's buf
 *   st_allreduce_destroy   : releases the communicator                                                              */
int st_allreduce_id_bytes(void);
int st_allreduce_unique_id(void* id_out /* host */);
int st_allreduce_init(const void* unique_id /* host */, int world, int rank, void** comm_out /* host */);
int st_allreduce_run(void* comm, float* buf, int64_t n, cudaStream_t stream);
int st_allreduce_broadcast(void* comm, float* buf, int64_t n, int root, cudaStream_t stream);
int st_allreduce_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* ST_B200_H_ */
