// st_embed.cu — the decoder's input step either side of the hot path:
//   dec_input = tgt_word_emb(outputs_data) + position_enc        transformer/Models.py:84-87, Embedding.py:21-29
// forward: gather one table row per token, add the sinusoid row of the token's position, round to TF32 (the result
// feeds the first decoder layer's projection GEMMs); backward: scatter-add of the output gradient into the table
// gradient, skipping the padding row (nn.Embedding(padding_idx=PAD) never receives a gradient for it).
// Both are tiny HBM/latency-bound kernels: one warp per token row, float4 lanes.
#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table, const float* __restrict__ pe,
                 int64_t pe_rows, T* __restrict__ out, int64_t n, int d, int vocab, int round_out) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < n; row += static_cast<int64_t>(gridDim.x) * 8) {
    int64_t t = idx[row];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);   // indices are validated on the host side of the ABI; never read out of bounds
    const float* src = table + t * d;
    const float* pr = pe ? pe + (row % pe_rows) * d : nullptr;
    for (int c = lane * 4; c < d; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(src + c);
      if (pr) {
        const float4 p = *reinterpret_cast<const float4*>(pr + c);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      if (sizeof(T) == 4 && round_out) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
      stv4(out + row * d + c, v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int64_t* __restrict__ idx, const T* __restrict__ dout, float* __restrict__ dtable, int64_t n,
                 int d, int vocab, int64_t padding_idx) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < n; row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t t = idx[row];
    if (t == padding_idx || t < 0 || t >= vocab) continue;   // warp-uniform
    float* dst = dtable + t * d;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = ldv4(dout + row * d + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

}  // namespace

int embed_fwd(cudaStream_t stream, const int64_t* idx, const float* table, const float* pe, int64_t pe_rows, void* out,
              int64_t n, int d, int vocab, int round_out, int dt) {
  if (n == 0) return ST_OK;
  ST_REQUIRE(d > 0 && (d & 3) == 0 && vocab > 0, "embed_fwd: d=%d must be a positive multiple of 4", d);
  ST_REQUIRE(!pe || pe_rows > 0, "embed_fwd: pe_rows must be positive");
  const int64_t blocks = (n + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(stream, PROF_EMBED, (pe ? 3.0 : 2.0) * n * d * 4);
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  if (dt == ST_DTYPE_F16) embed_fwd_kernel<__half><<<grid, 256, 0, stream>>>(idx, table, pe, pe_rows, static_cast<__half*>(out), n, d, vocab, round_out);
  else if (dt == ST_DTYPE_BF16) embed_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(idx, table, pe, pe_rows, static_cast<__nv_bfloat16*>(out), n, d, vocab, round_out);
  else embed_fwd_kernel<float><<<grid, 256, 0, stream>>>(idx, table, pe, pe_rows, static_cast<float*>(out), n, d, vocab, round_out);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int embed_bwd(cudaStream_t stream, const int64_t* idx, const void* dout, float* dtable, int64_t n, int d, int vocab,
              int64_t padding_idx, int zero_first, int dt) {
  ST_REQUIRE(d > 0 && (d & 3) == 0 && vocab > 0, "embed_bwd: d=%d must be a positive multiple of 4", d);
  if (zero_first) ST_CHECK_CUDA(cudaMemsetAsync(dtable, 0, static_cast<size_t>(vocab) * d * sizeof(float), stream));
  if (n == 0) return ST_OK;
  const int64_t blocks = (n + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(stream, PROF_EMBED, 2.0 * n * d * 4 + (zero_first ? 1.0 * vocab * d * 4 : 0.0));
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  if (dt == ST_DTYPE_F16) embed_bwd_kernel<__half><<<grid, 256, 0, stream>>>(idx, static_cast<const __half*>(dout), dtable, n, d, vocab, padding_idx);
  else if (dt == ST_DTYPE_BF16) embed_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(idx, static_cast<const __nv_bfloat16*>(dout), dtable, n, d, vocab, padding_idx);
  else embed_bwd_kernel<float><<<grid, 256, 0, stream>>>(idx, static_cast<const float*>(dout), dtable, n, d, vocab, padding_idx);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st

// ------------------------------------------------------------------------------------------------------------------
// Incremental-decode self-attention: ONE new query position per hypothesis attends to the cached
// keys / values of positions 0..t (Attention.py:78-90 with Lq = 1; no mask is needed — everything cached is in the
// past).  The tensor-core kernel would run a 128-row tile for a single valid row; here one warp owns one
// (hypothesis, head): it appends the new K / V rows to the time-major caches (L_max, n, d), then streams the t + 1
// cached rows with an online softmax in fp32.  qkv is the packed projection output (n, 3d) = [q | k | v].
namespace st {
namespace {

// slot_of (optional, time-major (L_max, n) int32): the cache slot that holds position j of hypothesis `hyp`'s history.
// Beam search re-parents hypotheses every step; with the table only its 4-byte entries are permuted (by the caller:
// slot_of[:t] <- slot_of[:t][:, parent]) instead of the 12 K/V caches (≈ 0.4 GB per position at width 10, 50 positions).
// The kernel records slot_of[t][hyp] = hyp for the row it appends.  NULL = every hypothesis owns slot `hyp` throughout.
template <int DK>
__global__ void __launch_bounds__(256)
decode_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ k_cache, float* __restrict__ v_cache, int t, int n,
                        int H, float scale_log2, float* __restrict__ ctx, int round_out, int* __restrict__ slot_of) {
  pdl_wait();
  pdl_trigger();
  constexpr int E = DK / 32;                       // elements per lane
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n * H) return;                          // warp-uniform
  const int hyp = w / H, h = w - hyp * H;
  const int d = H * DK;
  const int64_t row_stride = static_cast<int64_t>(n) * d;          // cache row (one position, all hypotheses)
  const int64_t col = static_cast<int64_t>(h) * DK + lane * E;
  const int64_t off = static_cast<int64_t>(hyp) * d + col;
  float q[E], kn[E], vn[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    q[e] = qkv[static_cast<int64_t>(hyp) * 3 * d + h * DK + lane * E + e];
    kn[e] = qkv[static_cast<int64_t>(hyp) * 3 * d + d + h * DK + lane * E + e];
    vn[e] = qkv[static_cast<int64_t>(hyp) * 3 * d + 2 * d + h * DK + lane * E + e];
    k_cache[t * row_stride + off + e] = kn[e];
    v_cache[t * row_stride + off + e] = vn[e];
  }
  if (slot_of != nullptr && h == 0 && lane == 0) slot_of[static_cast<int64_t>(t) * n + hyp] = hyp;
  float m = -INFINITY, l = 0.f, acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  int slots = hyp;                                 // lane i: slot of position (j & ~31) + i, refreshed every 32 positions
  for (int j = 0; j <= t; ++j) {
    if (slot_of != nullptr && (j & 31) == 0) slots = (j + lane < t) ? slot_of[static_cast<int64_t>(j + lane) * n + hyp] : hyp;
    const int slot = __shfl_sync(0xffffffffu, slots, j & 31);
    const int64_t src = j * row_stride + static_cast<int64_t>(slot) * d + col;
    float kk[E], vv[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      kk[e] = (j == t) ? kn[e] : k_cache[src + e];
      vv[e] = (j == t) ? vn[e] : v_cache[src + e];
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(q[e], kk[e], s);
    s = warp_sum(s) * scale_log2;
    const float m_new = fmaxf(m, s);
    const float alpha = exp2f(m - m_new);          // m == -inf -> 0
    const float p = exp2f(s - m_new);
    l = l * alpha + p;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = fmaf(acc[e], alpha, p * vv[e]);
    m = m_new;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float o = acc[e] * inv;
    ctx[off + e] = round_out ? tf32_rna(o) : o;
  }
}

}  // namespace

int decode_self_attn(cudaStream_t stream, const float* qkv, float* k_cache, float* v_cache, int t, int n, int H, int dk,
                     float* ctx, int round_out, int* slot_of) {
  ST_REQUIRE(n > 0 && H > 0 && t >= 0, "decode_self_attn: bad shape (n=%d H=%d t=%d)", n, H, t);
  ST_REQUIRE(dk == 32 || dk == 64 || dk == 128, "decode_self_attn: d_k must be 32, 64 or 128 (got %d)", dk);
  const int warps = n * H;
  const dim3 grid((warps + 7) / 8), block(256);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  ProfScope prof(stream, PROF_ATTN_FWD, 4.0 * n * H * static_cast<double>(t + 1) * dk);
  if (dk == 32) ST_CHECK_CUDA(launch_pdl(decode_self_attn_kernel<32>, grid, block, 0, stream, qkv, k_cache, v_cache, t, n, H, scale_log2, ctx, round_out, slot_of));
  else if (dk == 64) ST_CHECK_CUDA(launch_pdl(decode_self_attn_kernel<64>, grid, block, 0, stream, qkv, k_cache, v_cache, t, n, H, scale_log2, ctx, round_out, slot_of));
  else ST_CHECK_CUDA(launch_pdl(decode_self_attn_kernel<128>, grid, block, 0, stream, qkv, k_cache, v_cache, t, n, H, scale_log2, ctx, round_out, slot_of));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
