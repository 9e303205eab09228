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

__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table, const float* __restrict__ pe,
                 int64_t pe_rows, float* __restrict__ out, int64_t n, int d, int vocab, int round_out) {
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
      if (round_out) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
      *reinterpret_cast<float4*>(out + row * d + c) = v;
    }
  }
}

__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dout, float* __restrict__ dtable, int64_t n,
                 int d, int vocab, int64_t padding_idx) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < n; row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t t = idx[row];
    if (t == padding_idx || t < 0 || t >= vocab) continue;   // warp-uniform
    float* dst = dtable + t * d;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(dout + row * d + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

}  // namespace

int embed_fwd(cudaStream_t stream, const int64_t* idx, const float* table, const float* pe, int64_t pe_rows, float* out,
              int64_t n, int d, int vocab, int round_out) {
  if (n == 0) return ST_OK;
  ST_REQUIRE(d > 0 && (d & 3) == 0 && vocab > 0, "embed_fwd: d=%d must be a positive multiple of 4", d);
  ST_REQUIRE(!pe || pe_rows > 0, "embed_fwd: pe_rows must be positive");
  const int64_t blocks = (n + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(stream, PROF_EMBED, (pe ? 3.0 : 2.0) * n * d * 4);
  embed_fwd_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, stream>>>(idx, table, pe, pe_rows, out, n, d, vocab, round_out);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int embed_bwd(cudaStream_t stream, const int64_t* idx, const float* dout, float* dtable, int64_t n, int d, int vocab,
              int64_t padding_idx, int zero_first) {
  ST_REQUIRE(d > 0 && (d & 3) == 0 && vocab > 0, "embed_bwd: d=%d must be a positive multiple of 4", d);
  if (zero_first) ST_CHECK_CUDA(cudaMemsetAsync(dtable, 0, static_cast<size_t>(vocab) * d * sizeof(float), stream));
  if (n == 0) return ST_OK;
  const int64_t blocks = (n + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  ProfScope prof(stream, PROF_EMBED, 2.0 * n * d * 4 + (zero_first ? 1.0 * vocab * d * 4 : 0.0));
  embed_bwd_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, stream>>>(idx, dout, dtable, n, d, vocab, padding_idx);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
