// st_kernels.h — internal C++ interface of the kernel launchers (everything below the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/st_b200.h"

namespace st {

// Inverted dropout with a stateless counter RNG: element idx is kept iff hash(seed, idx) >= thresh
// and then scaled by `scale` = 1/(1-p).  thresh == 0 disables dropout.
struct DropoutCfg {
  uint32_t thresh = 0;     // 16-bit threshold (drop probability thresh/65536) for the pair-hash scheme
  float scale = 1.f;
  uint64_t seed = 0;
  uint32_t thresh32 = 0;   // 32-bit threshold (drop probability thresh32/2^32) for the attention xor-key scheme
  float scale32 = 1.f;
};
DropoutCfg make_dropout(float p, uint64_t seed);

// ---- st_ln.cu
int add_ln_fwd(cudaStream_t stream, const float* a, const float* b, const float* gamma, const float* beta, float* out,
               float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
               const DropoutCfg& drop, const float* post = nullptr, int64_t post_rows = 0);
int add_ln_bwd(cudaStream_t stream, const float* dy, const float* z, const float* mean, const float* rstd,
               const float* gamma, float* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d,
               int round_out, const DropoutCfg& drop, const float* gate = nullptr, float gate_scale = 1.f);
// the same with 16-bit activations (ST_DTYPE_*): see st_ln.cu
int add_ln_fwd_any(cudaStream_t stream, int in_dt, int out_dt, const void* a, const void* b, const float* gamma, const float* beta,
                   void* out, float* z_out, float* mean_out, float* rstd_out, int64_t rows, int d, float eps, int round_out,
                   const DropoutCfg& drop, const float* post = nullptr, int64_t post_rows = 0, void* out_h16 = nullptr);
int add_ln_bwd_any(cudaStream_t stream, int dt, const void* dy, const float* z, const float* mean, const float* rstd,
                   const float* gamma, void* dz, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, int round_out,
                   const DropoutCfg& drop, const float* gate = nullptr, float gate_scale = 1.f);
// amax (optional device scalar): the sums are divided by grad_scale_from_amax(*amax) (mixed mode, st_common.cuh)
int colsum_add_any(cudaStream_t stream, int dt, const void* x, int64_t ld, int64_t rows, int cols, float* out,
                   const float* amax = nullptr);
int amax_abs(cudaStream_t stream, const float* x, int64_t n, float* out);
int add_ln_bwd_mixed(cudaStream_t stream, const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                     void* dz16, float* dgamma, float* dbeta, float* dzsum, int64_t rows, int d, const DropoutCfg& drop,
                     const float* amax, float* clear_scalar = nullptr);
int cast_2d(cudaStream_t stream, const void* src, int src_dt, int64_t lds, void* dst, int dst_dt, int64_t ldd, int64_t rows, int cols,
            float scale = 1.f);
int round_tf32_2d(cudaStream_t stream, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols);
int colsum_add(cudaStream_t stream, const float* x, int64_t ld, int64_t rows, int cols, float* out);

// ---- st_embed.cu
// dt: element type of out / dout (ST_DTYPE_*); the table, its gradient and pe are fp32
int embed_fwd(cudaStream_t stream, const int64_t* idx, const float* table, const float* pe, int64_t pe_rows, void* out,
              int64_t n, int d, int vocab, int round_out, int dt = 0);
int embed_bwd(cudaStream_t stream, const int64_t* idx, const void* dout, float* dtable, int64_t n, int d, int vocab,
              int64_t padding_idx, int zero_first, int dt = 0);

int decode_self_attn(cudaStream_t stream, const float* qkv, float* k_cache, float* v_cache, int t, int n, int H, int dk,
                     float* ctx, int round_out, int* slot_of);
// st_beam.cu: one position of beam-search bookkeeping (log-softmax, score update, top-`beam` of beam x V, back-pointers)
int beam_step(cudaStream_t stream, const float* logits, int64_t ld, int B, int beam, int V, int first, int eos, int pad,
              float* scores, uint8_t* done, int64_t* prev_k, int64_t* next_y, int64_t* parent, int64_t* tokens);

// ---- st_ctc.cu
int64_t ctc_ws_floats(int B, int T, int S_max);
int ctc_fwd_bwd(cudaStream_t stream, const float* logits, int64_t ld, const int64_t* targets, int64_t ld_tgt,
                const int64_t* in_len, const int64_t* tgt_len, int blank, int B, int T, int V, int L_max, float* nll,
                const float* scale, float* grad, int64_t ldg, float* ws, int64_t ws_floats, int grad_only = 0);

// ---- st_optim.cu
int sumsq_add(cudaStream_t s, const float* x, int64_t n, float* out);
int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
              float eps, int step, float max_norm, float gscale, const float* sumsq, void* p_twin = nullptr, int twin_dt = 0);

// ---- st_lsce.cu
struct LsceArgs {
  const float* logits;   // (N, ldl)
  int64_t ldl;
  const int64_t* target; // (N,)  sparse mode, or nullptr
  const float* q_dense;  // (N, V) dense soft target (CrossEntropyLoss mode), or nullptr
  const float* one_hot;  // (V,)  base smoothed row (sparse mode)
  const float* weight;   // (V,)
  float confidence;
  int64_t padding_idx;   // rows with target == padding_idx are zeroed when padding_idx >= 0
  float inv_z;           // 1/N if size_average else 1
  int64_t N;
  int V;
  float* row_loss;       // (N,) workspace
  float* loss;           // scalar out
  float* grad;           // (N, ldg) out, may be nullptr
  int64_t ldg;
};
int lsce_fwd_bwd(cudaStream_t stream, const LsceArgs& a);

// ---- st_attn.cu (TF32 operands), st_attn16.cu (fp16 / bf16 operands)
struct AttnArgs {
  int B, H, Lq, Lk, dk;
  int dtype = 0;                 // ST_DTYPE_*: element type of q, k, v, ctx (and dctx, dq, dk, dv in the backward)
  const void* q; int64_t ldq;    // row (b*Lq + i) at q + row*ldq, head h at column h*dk
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const uint8_t* mask;           // nullable; element (b,i,j) at mask[b*ms_b + i*ms_q + j*ms_k]; nonzero = masked
  int64_t ms_b, ms_q, ms_k;
  // length-aware masking (no mask tensor built or scanned): keys j >= k_len[b] are masked (Utils.padding_info_mask,
  // Utils.py:41-57) and, with causal != 0, keys j > i (Utils.feature_info_mask, Utils.py:60-70).  Combine with `mask` by OR.
  const int64_t* k_len = nullptr;
  int causal = 0;
  float scale;                   // 1/sqrt(dk)
  DropoutCfg drop;               // dropout on the attention probabilities (Attention.py:89)
  void* ctx; int64_t ldctx;      // (B*Lq, H*dk) merged heads
  float* lse;                    // (B, H, Lq) natural-log sum-exp of the scaled, masked scores
  float* attn;                   // nullable (B, H, Lq, Lk): post-dropout probabilities (return value of the module)
};
int attn_fwd(cudaStream_t stream, const AttnArgs& a);

struct AttnBwdArgs {
  AttnArgs f;                    // forward problem (q,k,v,mask,lse,ctx as produced by attn_fwd)
  const void* dctx; int64_t lddctx;
  float* delta;                  // (B, H, Lq) workspace: rowsum(dctx * ctx)
  void* dq; int64_t lddq;        // same indexing as q/k/v
  void* dk_; int64_t lddk;
  void* dv; int64_t lddv;
  // optional [H*dk] each: column sums of dq / dk / dv are ACCUMULATED here (bias gradients of the Q/K/V projections,
  // Attention.py:74-76); honoured by the pipelined TF32 kernels (d_k <= 64) — attn_bwd_fuses_bias() tells the caller
  float* dbq = nullptr; float* dbk = nullptr; float* dbv = nullptr;
};
bool attn_bwd_fuses_bias(int dk);
int attn_bwd(cudaStream_t stream, const AttnBwdArgs& a);
// st_attn16.cu: the 16-bit kernels (d_k = 64 or 128); called by attn_fwd / attn_bwd when dtype != ST_DTYPE_F32
int attn16_fwd(cudaStream_t stream, const AttnArgs& a);
int attn16_bwd(cudaStream_t stream, const AttnBwdArgs& a);

}  // namespace st
