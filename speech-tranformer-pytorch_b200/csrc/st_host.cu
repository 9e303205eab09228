// st_host.cu — error string, device query and TMA tensor-map construction.
#include "st_host.h"

#include <cudaTypedefs.h>
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace st {

namespace {
std::mutex g_err_mutex;
char g_err[1024] = "";
}  // namespace

void set_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lock(g_err_mutex);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

namespace {
std::atomic<long long> g_launches{0};
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct Option { const char* name; int value; };
Option g_options[] = {
    {"gemm_max_ctas", 0},   // 0 = one CTA per SM; >0 caps the persistent grid (tests multi-tile paths)
    {"attn_bwd_simple", 0}, // 1 = use the non-pipelined attention backward kernels for every d_k (A/B testing)
    {"gemm_cluster", 0},    // 0 = heuristic; 1 = independent CTAs, 2 = multicast B tile, 3 = CTA-pair MMA (cta_group::2)
    {"gemm_generic_epilogue", 0},   // 1 = force the generic (runtime-flag) GEMM epilogue (tests)
    {"gemm_rows16", 2},   // 16-bit-output epilogues without aux: 0 = fp32 transpose through smem, 1 = row layout with direct stores,
                          // 2 = row layout, packed result staged through smem (st_gemm_impl.cuh epilogue_rows16; default:
                          // K = 512 GEMMs 8-10 % faster, bit-identical results — tools/check_rows16.py)
    {"gemm_clc", 0},      // CTA-pair GEMM: 1 = cluster-launch-control tile scheduling (one cluster per tile, resident pairs cancel
                          // pending ones and take their tiles) instead of a persistent grid with a static stride
    {"gemm_bn", 0},       // 0 = heuristic; 64/128/256 forces the GEMM tile width (tuning / tests)
    {"attn_trace", 0},         // 1 = dK/dV kernel CTA (0,0,0) records its pipeline timeline (st_debug_read_trace)
    {"attn_dkv_small_split", 0},   // CTAs per (batch, head) of the single-query-tile dK/dV kernel (0 = default 1)
    {"attn_dkv_no_small", 0},  // 1 = never use the single-query-tile dK/dV kernel (A/B testing)
    {"attn_fuse_bias", 0},     // 1 = dq/dk/dv bias column sums from the attention-backward epilogues (measured slower)
    {"attn_dq_res_smem", 0},   // 1 = dQ kernel keeps its resident Q/dO tiles in shared memory (.ss MMAs) instead of TMEM
    {"ln_bwd_registers", 0},   // 1 = LayerNorm backward with register prefetch (one row per warp in flight) instead of the cp.async ring
    {"ln_fwd_registers", 0},   // same for the forward
    {"side_streams", -1},      // independent kernels of a backward operator on library-owned side streams (st_host.h Fork):
                               // 1 = on, 0 = off, -1 = unset (on unless ST_SIDE_STREAMS=0 in the environment)
    {"pdl", -1},
    {"pdl_graphs", 1},         // keep programmatic dependent launch while the stream is being captured into a CUDA graph (decode:
                               // 73.8 -> 71.0 ms per beam search); 0 / ST_PDL_GRAPHS=0 = plain launches under capture               // programmatic dependent launch: 1 = on, 0 = off, -1 = unset (on unless ST_PDL=0 in the environment)
    {"attn_dkv_res_smem", 0},  // resident K/V tiles of the dK/dV kernel: 0 = heuristic (smem when Lq <= 128), 1 = smem, 2 = TMEM
};
}  // namespace

int set_option(const char* name, int value) {
  for (auto& o : g_options)
    if (strcmp(o.name, name) == 0) { o.value = value; return ST_OK; }
  set_error("unknown option '%s'", name);
  return ST_ERR_INVALID;
}
int get_option(const char* name) {
  for (auto& o : g_options)
    if (strcmp(o.name, name) == 0) return o.value;
  return 0;
}

int pdl_allowed(cudaStream_t s) {
  static Option* opt = [] {
    Option* o = nullptr;
    for (auto& x : g_options)
      if (strcmp(x.name, "pdl") == 0) o = &x;
    if (o->value < 0) {
      const char* e = getenv("ST_PDL");
      o->value = (e && e[0] == '0') ? 0 : 1;
    }
    return o;
  }();
  if (opt->value <= 0) return 0;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (st == cudaStreamCaptureStatusNone) return 1;
  static Option* gopt = [] {
    Option* o = nullptr;
    for (auto& x : g_options)
      if (strcmp(x.name, "pdl_graphs") == 0) o = &x;
    const char* e = getenv("ST_PDL_GRAPHS");
    if (e) o->value = (e[0] == '1') ? 1 : 0;
    return o;
  }();
  return gopt->value > 0 ? 1 : 0;
}

// ---- per-kernel-class event timing
namespace {
struct ProfRec { cudaEvent_t a, b; int cls; double work; long long tag; };
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace

void profile_enable(int on) { g_prof_on.store(on ? 1 : 0, std::memory_order_relaxed); }

ProfScope::ProfScope(cudaStream_t s, ProfClass cls, double work, long long tag) : stream(s), slot(-1) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{};
  r.cls = cls;
  r.work = work;
  r.tag = tag;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back(r);
  slot = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (slot < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[slot].b, stream);
}

int profile_read(int cls, double* ms, double* work, long long* launches) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  double t = 0, w = 0;
  long long n = 0;
  for (auto& r : g_prof) {
    if (r.cls != cls) continue;
    float e = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) {
      set_error("profile_read: event query failed");
      return ST_ERR_CUDA;
    }
    t += e; w += r.work; ++n;
  }
  *ms = t; *work = w; *launches = n;
  return ST_OK;
}

int profile_dump(const char* path) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  FILE* f = fopen(path, "w");
  if (!f) { set_error("profile_dump: cannot open %s", path); return ST_ERR_INVALID; }
  fprintf(f, "index,class,work,ms,tag\n");
  int i = 0;
  for (auto& r : g_prof) {
    float e = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&e, r.a, r.b);
    fprintf(f, "%d,%d,%.0f,%.6f,%lld\n", i++, r.cls, r.work, e, r.tag);
  }
  fclose(f);
  return ST_OK;
}

void profile_reset() {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
}

// ---- side streams
namespace {
struct SideDev {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ok = false;
  std::recursive_mutex mu;   // held from a Fork's first branch() to its join(): the streams and events are shared by every host
                             // thread that drives this device (nested Forks of one operator run on the same thread)
};
SideDev g_side[64];
std::mutex g_side_mu;

SideDev* side_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideDev& d = g_side[dev];
  if (!d.ok) {
    std::lock_guard<std::mutex> lk(g_side_mu);
    if (!d.ok) {
      bool good = cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < 2 && good; ++i)
        good = cudaStreamCreateWithFlags(&d.s[i], cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&d.join[i], cudaEventDisableTiming) == cudaSuccess;
      if (!good) { cudaGetLastError(); return nullptr; }
      d.ok = true;
    }
  }
  return &d;
}

bool side_streams_on(cudaStream_t s) {
  static Option* opt = [] {
    Option* o = nullptr;
    for (auto& x : g_options)
      if (strcmp(x.name, "side_streams") == 0) o = &x;
    if (o->value < 0) {
      const char* e = getenv("ST_SIDE_STREAMS");
      o->value = (e && e[0] == '0') ? 0 : 1;
    }
    return o;
  }();
  if (opt->value <= 0) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st == cudaStreamCaptureStatusNone;
}
}  // namespace

Fork::Fork(cudaStream_t main) : main_(main), dev_(side_streams_on(main) ? side_dev() : nullptr) {}

cudaStream_t Fork::branch(int i) {
  SideDev* d = static_cast<SideDev*>(dev_);
  if (!d || i < 0 || i > 1) return main_;
  const bool first = !used_[0] && !used_[1];
  if (first) d->mu.lock();
  // an event wait refers to the most recent record at the time of the call: one fork event per device is enough
  if (cudaEventRecord(d->fork, main_) != cudaSuccess || cudaStreamWaitEvent(d->s[i], d->fork, 0) != cudaSuccess) {
    cudaGetLastError();
    if (first) d->mu.unlock();
    return main_;
  }
  used_[i] = true;
  return d->s[i];
}

int Fork::join() {
  SideDev* d = static_cast<SideDev*>(dev_);
  if (!d || (!used_[0] && !used_[1])) return ST_OK;
  int status = ST_OK;
  for (int i = 0; i < 2; ++i) {
    if (!used_[i]) continue;
    used_[i] = false;
    if (cudaEventRecord(d->join[i], d->s[i]) != cudaSuccess || cudaStreamWaitEvent(main_, d->join[i], 0) != cudaSuccess) {
      set_error("side streams: join failed (%s)", cudaGetErrorString(cudaGetLastError()));
      status = ST_ERR_CUDA;
    }
  }
  d->mu.unlock();
  return status;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int atom32) {
  return make_tmap(out, 0, base, rank, dims, strides_bytes, box, atom32);
}

int make_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int atom32) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  const CUtensorMapDataType cu_dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                  : (dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = encode(out, cu_dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                      gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] stride0 %llu box [%u,%u,%u] base %p",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, base);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

}  // namespace st
