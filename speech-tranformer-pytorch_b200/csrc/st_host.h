// st_host.h — host-side plumbing shared by the launchers: error reporting across the C ABI and
// TMA tensor-map construction (driver entry point resolved at run time; no libcuda link).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/st_b200.h"

namespace st {

// status codes (ST_OK, ST_ERR_*) come from the public header

void count_launch();
long long launch_count();
void set_error(const char* fmt, ...);
const char* last_error();

#define ST_CHECK_CUDA(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      st::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return ST_ERR_CUDA;                                                                        \
    }                                                                                                \
  } while (0)

// after every kernel launch: count it (st_launch_count over the C ABI) and surface launch errors
#define ST_CHECK_LAUNCH()                  \
  do {                                     \
    st::count_launch();                    \
    ST_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define ST_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      st::set_error(__VA_ARGS__);          \
      return ST_ERR_INVALID;           \
    }                                      \
  } while (0)

#define ST_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != ST_OK) return _s; \
  } while (0)

// Build a tiled tensor map over fp32 data with 128-byte swizzle.  dims/strides are innermost
// first; strides_bytes[i] is the byte stride of dim i+1 (dim 0 is contiguous).  Out-of-bounds
// elements are zero-filled.
// atom32: 0 = SWIZZLE_128B (16-byte chunks; K-major operands), 1 = SWIZZLE_128B_ATOM_32B (32-byte chunks; the only
// layout tcgen05 accepts for MN-major TF32 operands, UMMA layout type SWIZZLE_128B_BASE32B).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int atom32 = 0);
// Same for any element type: dtype 0 = fp32, 1 = fp16, 2 = bf16 (ST_DTYPE_*).  16-bit operands always use the plain
// 128-byte swizzle (atom32 = 0): box[0] = 64 elements.
int make_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int atom32 = 0);

int num_sms();

// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl may start while its predecessor in
// the stream is still draining: its CTAs run their prologue (mbarrier init, TMEM allocation, descriptor prefetch) and
// then block in pdl_wait() (griddepcontrol.wait, st_common.cuh) until the predecessor has completed and its writes are
// visible.  EVERY thread of such a kernel must execute pdl_wait() before its first global-memory access and before any
// early exit (a kernel that finished without waiting would let its successor overtake the predecessor's writes).
// pdl_allowed: option "pdl" (default 1, environment ST_PDL=0 disables); under stream capture option "pdl_graphs" (default 1).
int pdl_allowed(cudaStream_t s);
template <typename... P, typename... A>
cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_allowed(s);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

// ---- optional per-kernel-class timing (CUDA events on the launching stream), used by bench.py for the
// roofline numbers.  Disabled by default; when disabled a ProfScope costs one relaxed atomic load.
enum ProfClass : int {
  PROF_GEMM = 0, PROF_ATTN_FWD, PROF_ATTN_DKV, PROF_ATTN_DQ, PROF_ATTN_DELTA, PROF_LN_FWD, PROF_LN_BWD, PROF_ROUND,
  PROF_COLSUM, PROF_LSCE, PROF_SUMSQ, PROF_ADAM, PROF_EMBED, PROF_CTC, PROF_NUM
};
struct ProfScope {
  ProfScope(cudaStream_t s, ProfClass cls, double work, long long tag = 0);   // tag: free-form id kept in the dump  // work: algorithmic FLOPs (tensor kernels) or bytes (HBM kernels)
  ~ProfScope();
  cudaStream_t stream;
  int slot;
};
void profile_enable(int on);
int profile_read(int cls, double* ms, double* work, long long* launches);  // synchronises; sums since the last reset
void profile_reset();
int profile_dump(const char* path);  // CSV: one line per recorded launch

// ---- side streams.  The backward operators launch kernels that do not depend on each other (the weight-gradient GEMMs
// and bias column sums against the input-gradient chain; attention's dQ against dK/dV).  Each is a full-device kernel at
// the encoder's shapes, but at the decoder's (1600 rows) and in every kernel's last wave most SMs idle: branches run on
// library-owned non-blocking streams (two per device) and rejoin the caller's stream before the operator returns, so the
// caller sees ordinary stream semantics.  Off under stream capture and with option "side_streams" = 0.
class Fork {
 public:
  explicit Fork(cudaStream_t main);
  // the stream of branch i (0 or 1), ordered after everything enqueued on the main stream so far; the main stream itself
  // when side streams are off
  cudaStream_t branch(int i);
  // the main stream waits for the branches used so far (also run by the destructor; call it to see the error code)
  int join();
  ~Fork() { join(); }
  Fork(const Fork&) = delete;
  Fork& operator=(const Fork&) = delete;

 private:
  cudaStream_t main_;
  void* dev_;          // per-device state, null = disabled
  bool used_[2] = {false, false};
};

// debug / tuning options (st_set_option over the C ABI); unknown names are rejected.
int set_option(const char* name, int value);
int get_option(const char* name);

}  // namespace st
