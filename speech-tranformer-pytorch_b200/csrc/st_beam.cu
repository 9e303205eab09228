// st_beam.cu — one position of beam-search bookkeeping on the device (Beam.advance, transformer/Beam.py:43-74, as driven by
// Decode.py:120-160): log-softmax of the hypotheses' logits, running-score update, the `beam` best of beam x vocab
// candidates per utterance with integer back-pointers, freezing of finished utterances, and the re-parenting / next-token
// vectors the incremental decoder consumes.  One block per utterance; the candidate scan is `beam` argmax passes over
// beam x V values that sit in L2 (43 370 floats per utterance at width 10, V = 4337).
#include <math_constants.h>

#include "st_common.cuh"
#include "st_host.h"
#include "st_kernels.h"

namespace st {

namespace {

constexpr int BEAM_THREADS = 512;
constexpr int BEAM_MAX = 32;

struct Cand { float v; int c; };

// total order of the selection: higher value first, lower flat index first among equals (deterministic ties)
__device__ __forceinline__ bool better(float v, int c, float bv, int bc) { return v > bv || (v == bv && c < bc); }

__device__ __forceinline__ Cand warp_best(Cand x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, x.v, off);
    const int oc = __shfl_xor_sync(0xffffffffu, x.c, off);
    if (better(ov, oc, x.v, x.c)) { x.v = ov; x.c = oc; }
  }
  return x;
}

__global__ void __launch_bounds__(BEAM_THREADS)
beam_step_kernel(const float* __restrict__ logits, int64_t ld, int beam, int V, int first, int eos, int pad,
                 float* __restrict__ scores, uint8_t* __restrict__ done, int64_t* __restrict__ prev_k,
                 int64_t* __restrict__ next_y, int64_t* __restrict__ parent, int64_t* __restrict__ tokens) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float s_off[BEAM_MAX];            // running score - logsumexp of each expanded hypothesis
  __shared__ float s_redf[BEAM_THREADS / 32];
  __shared__ Cand s_redc[BEAM_THREADS / 32];
  __shared__ Cand s_pick[BEAM_MAX];

  if (done[b]) {   // frozen: the beams keep their scores and repeat themselves (decode.beam_search)
    for (int k = tid; k < beam; k += BEAM_THREADS) {
      prev_k[b * beam + k] = k;
      next_y[b * beam + k] = pad;
      parent[b * beam + k] = static_cast<int64_t>(b) * beam + k;
      tokens[b * beam + k] = pad;
    }
    return;
  }

  const int rows = first ? 1 : beam;           // first position: all beams are identical, expand one (Beam.py:49-52)
  const float* lg = logits + static_cast<int64_t>(b) * beam * ld;

  // ---- log-softmax normaliser of every expanded row
  for (int k = 0; k < rows; ++k) {
    const float* x = lg + static_cast<int64_t>(k) * ld;
    float m = -CUDART_INF_F;
    for (int v = tid; v < V; v += BEAM_THREADS) m = fmaxf(m, x[v]);
    m = warp_max(m);
    if (lane == 0) s_redf[warp] = m;
    __syncthreads();
    m = s_redf[0];
#pragma unroll
    for (int w = 1; w < BEAM_THREADS / 32; ++w) m = fmaxf(m, s_redf[w]);
    __syncthreads();
    float s = 0.f;
    for (int v = tid; v < V; v += BEAM_THREADS) s += expf(x[v] - m);
    s = warp_sum(s);
    if (lane == 0) s_redf[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < BEAM_THREADS / 32; ++w) tot += s_redf[w];
      s_off[k] = (first ? 0.f : scores[b * beam + k]) - (m + logf(tot));
    }
    __syncthreads();
  }

  // ---- the `beam` best candidates, best first: pass j takes the best candidate that comes after pick j-1 in the order
  float last_v = CUDART_INF_F;
  int last_c = -1;
  for (int j = 0; j < beam; ++j) {
    Cand best{-CUDART_INF_F, 0x7fffffff};
    for (int k = 0; k < rows; ++k) {
      const float* x = lg + static_cast<int64_t>(k) * ld;
      const float off = s_off[k];
      const int c0 = k * V;
      for (int v = tid; v < V; v += BEAM_THREADS) {
        const float val = x[v] + off;
        const int c = c0 + v;
        const bool after = val < last_v || (val == last_v && c > last_c);   // strictly after the previous pick
        if (after && better(val, c, best.v, best.c)) { best.v = val; best.c = c; }
      }
    }
    best = warp_best(best);
    if (lane == 0) s_redc[warp] = best;
    __syncthreads();
    if (warp == 0) {
      Cand x = lane < BEAM_THREADS / 32 ? s_redc[lane] : Cand{-CUDART_INF_F, 0x7fffffff};
      x = warp_best(x);
      if (lane == 0) s_pick[j] = x;
    }
    __syncthreads();
    last_v = s_pick[j].v;
    last_c = s_pick[j].c;
  }

  // ---- outputs
  for (int j = tid; j < beam; j += BEAM_THREADS) {
    const Cand p = s_pick[j];
    const int pk = p.c / V, y = p.c - pk * V;                       // integer back-pointer (Beam.py:66)
    scores[b * beam + j] = p.v;
    prev_k[b * beam + j] = pk;
    next_y[b * beam + j] = y;
    parent[b * beam + j] = static_cast<int64_t>(b) * beam + pk;
    tokens[b * beam + j] = y;
    if (j == 0 && y == eos) done[b] = 1;                             // finished when the best hypothesis ends (Beam.py:70-72)
  }
}

}  // namespace

int beam_step(cudaStream_t stream, const float* logits, int64_t ld, int B, int beam, int V, int first, int eos, int pad,
              float* scores, uint8_t* done, int64_t* prev_k, int64_t* next_y, int64_t* parent, int64_t* tokens) {
  ST_REQUIRE(B > 0 && beam > 0 && beam <= BEAM_MAX && V > 0 && ld >= V, "beam_step: bad shape (B=%d beam=%d <= %d, V=%d, ld=%lld)", B,
             beam, BEAM_MAX, V, (long long)ld);
  ST_REQUIRE(static_cast<int64_t>(beam) * V < (1ll << 31), "beam_step: beam * V must fit in 31 bits");
  ST_REQUIRE(logits && scores && done && prev_k && next_y && parent && tokens, "beam_step: null pointer");
  ST_CHECK_CUDA(launch_pdl(beam_step_kernel, dim3(B), dim3(BEAM_THREADS), 0, stream, logits, ld, beam, V, first, eos, pad, scores,
                           done, prev_k, next_y, parent, tokens));
  ST_CHECK_LAUNCH();
  return ST_OK;
}

}  // namespace st
