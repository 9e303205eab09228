// st_attn.cuh — declarations shared by the attention kernels (st_attn.cu: forward + simple backward,
// st_attn_bwd.cu: warp-specialised pipelined backward).
#pragma once
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

constexpr float kLog2e = 1.4426950408889634f;

struct AttnDev {
  int B, H, Lq, Lk;
  const uint8_t* mask;
  int64_t ms_b, ms_q, ms_k;
  const int64_t* k_len;   // nullable (B,): keys >= k_len[b] are masked
  int causal;             // keys > query index are masked
  float scale;        // 1/sqrt(dk)
  float scale_log2;   // scale * log2(e)
  uint32_t drop_thresh;   // 32-bit threshold of the xor-key dropout scheme (0 = off)
  float drop_scale;
  uint64_t drop_seed;
  void* ctx; int64_t ldctx;
  float* lse2;        // (B,H,Lq) log2-domain log-sum-exp of the scaled masked scores
  float* attn;
  // backward
  const float* delta;
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
  float ds_boost;     // 16-bit kernels: power-of-two factor applied to dS before it is rounded to fp16 (undone in the epilogue)
  float* dbq; float* dbk; float* dbv;   // optional [H*dk]: column sums of dq / dk / dv ACCUMULATED here (projection bias gradients)
  int trace;          // debug: record the pipeline timeline of CTA (0,0,0) (option "attn_trace")
};

// number of keys of batch b that are not excluded by the length vector
__device__ __forceinline__ int key_limit(const AttnDev& p, int b) {
  if (p.k_len == nullptr) return p.Lk;
  const long long l = __ldg(p.k_len + b);
  return l < 0 ? 0 : (l < p.Lk ? static_cast<int>(l) : p.Lk);
}

// bit i set <=> (query row, key k0+i) is masked.  Keys beyond Lk (and beyond k_len[b]) are always masked.
__device__ __forceinline__ uint32_t mask_bits_row(const AttnDev& p, int b, int row, bool row_ok, int k0) {
  uint32_t bits = 0;
  const int valid = key_limit(p, b) - k0;  // number of in-range keys in this 32-chunk (may be <= 0 or > 32)
  if (valid < 32) bits = valid <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << valid);
  if (p.causal) {                          // keys > row
    const int keep = row - k0 + 1;         // keys k0 .. row stay
    if (keep < 32) bits |= keep <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << keep);
  }
  if (p.mask != nullptr && row_ok && valid > 0) {
    const uint8_t* m = p.mask + b * p.ms_b + static_cast<int64_t>(row) * p.ms_q + static_cast<int64_t>(k0) * p.ms_k;
    if (p.ms_k == 1 && valid >= 32 && ((reinterpret_cast<uintptr_t>(m) & 3) == 0)) {
      const uint32_t* m4 = reinterpret_cast<const uint32_t*>(m);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const uint32_t v = m4[w];
        bits |= ((v & 0x000000FFu) ? 1u : 0u) << (w * 4 + 0);
        bits |= ((v & 0x0000FF00u) ? 1u : 0u) << (w * 4 + 1);
        bits |= ((v & 0x00FF0000u) ? 1u : 0u) << (w * 4 + 2);
        bits |= ((v & 0xFF000000u) ? 1u : 0u) << (w * 4 + 3);
      }
    } else {
      const int n = valid < 32 ? valid : 32;
      for (int i = 0; i < n; ++i) bits |= (m[static_cast<int64_t>(i) * p.ms_k] ? 1u : 0u) << i;
    }
  }
  return bits;
}

// every query row of a tile sees the same mask bits (no mask, lengths only, or a key-padding mask with stride 0 over queries)
__device__ __forceinline__ bool mask_is_row_invariant(const AttnDev& p) {
  return !p.causal && (p.mask == nullptr || p.ms_q == 0);
}

// Key extent of batch `b`: 1 + index of the last key that some query may attend to, 0 if every key is masked.  With a length
// vector it is read, not scanned; under a key-padding mask (one mask row shared by all queries, ms_q == 0 — what
// Utils.padding_info_mask builds) the mask row is scanned; Lk when the mask depends on the query.  Keys at or beyond the extent
// have probability exactly 0 for every query, so whole key tiles beyond it can be skipped without changing any result
// (ragged batches: utterances shorter than T_max).
// Called by ALL threads of the block (two __syncthreads inside); `slot` is a shared int.
__device__ __forceinline__ int block_key_extent(const AttnDev& p, int b, int* slot) {
  const int lim = key_limit(p, b);
  if (p.mask == nullptr || p.ms_q != 0) return p.k_len != nullptr ? (lim > 0 ? lim : 0) : p.Lk;
  if (threadIdx.x == 0) *slot = 0;
  __syncthreads();
  const uint8_t* m = p.mask + b * p.ms_b;
  int last = 0;
  for (int j = threadIdx.x; j < lim; j += blockDim.x)
    if (m[static_cast<int64_t>(j) * p.ms_k] == 0) last = j + 1;
  if (last) atomicMax(slot, last);
  __syncthreads();
  return *slot;
}

// host helpers (st_attn.cu)
int make_act_tmap(CUtensorMap* m, const void* base, int64_t ld, int cols, int L, int B, int box_rows, int atom32, int dk);
// 16-bit activations: {64 columns, L rows, cols/64 column groups, B}, box {64, box_rows, dk/64, 1}, plain 128-byte swizzle
int make_act_tmap16(CUtensorMap* m, int dtype, const void* base, int64_t ld, int cols, int L, int B, int box_rows, int dk);
AttnDev attn_to_dev(const AttnArgs& a);

// st_attn_bwd.cu: pipelined dQ and dK/dV kernels for d_k in {32, 64}; p already carries the backward pointers
int attn_bwd_pipelined(cudaStream_t s, const AttnBwdArgs& a, const AttnDev& p);
int attn_read_trace(unsigned long long* host_out, int n);   // returns the number of slots
int attn_read_fwd_trace(unsigned long long* host_out, int n);

}  // namespace st
