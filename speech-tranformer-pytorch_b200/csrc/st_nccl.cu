// st_nccl.cu — the data-parallel gradient exchange of the C ABI: st_allreduce_{unique_id,init,run,destroy}.
//
// Reference: train_multi.py:20,128 (one process per GPU), :161-163 (hvd.DistributedOptimizer: all-reduce of every gradient),
// :176-177 (broadcast of the initial state).  Here a host that is not PyTorch — or PyTorch itself, parallel.py
// `collective="library"` — reduces the ONE flat fp32 gradient buffer with ncclAllReduce(sum) over NVLink / NVSwitch through
// these four calls; averaging is folded into st_adam_step's grad_scale.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded),
// so libst_b200.so has no link-time dependency on it and loads on a box without NCCL; the calls then fail with a message.
#include <dlfcn.h>
#include <mutex>
#include <string.h>

#include "st_host.h"

namespace st {
namespace {

struct NcclUniqueId { char internal[128]; };            // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
using NcclComm = void*;
constexpr int kNcclFloat = 7, kNcclSum = 0;             // ncclFloat32, ncclSum (nccl.h enums)

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.handle, "ncclAllReduce"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(a.handle, "ncclBroadcast"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.Broadcast && a.CommDestroy;
  });
  return a;
}

int need_api() {
  if (!api().ok) {
    set_error("NCCL is not available (dlopen libnccl.so.2: %s)", api().handle ? "missing symbols" : dlerror());
    return ST_ERR_DEVICE;
  }
  return ST_OK;
}

#define ST_CHECK_NCCL(expr)                                                                                     \
  do {                                                                                                          \
    const int _r = (expr);                                                                                      \
    if (_r != 0) {                                                                                              \
      st::set_error("%s failed: %s", #expr, api().GetErrorString ? api().GetErrorString(_r) : "NCCL error");    \
      return ST_ERR_CUDA;                                                                                       \
    }                                                                                                           \
  } while (0)

}  // namespace
}  // namespace st

using namespace st;

extern "C" {

int st_allreduce_id_bytes(void) { return static_cast<int>(sizeof(NcclUniqueId)); }

int st_allreduce_unique_id(void* id_out) {
  ST_REQUIRE(id_out != nullptr, "st_allreduce_unique_id: null output");
  ST_TRY(need_api());
  NcclUniqueId id;
  ST_CHECK_NCCL(api().GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return ST_OK;
}

int st_allreduce_init(const void* unique_id, int world, int rank, void** comm_out) {
  ST_REQUIRE(unique_id && comm_out && world >= 1 && rank >= 0 && rank < world, "st_allreduce_init: bad arguments (world %d rank %d)",
             world, rank);
  ST_TRY(need_api());
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  NcclComm comm = nullptr;
  ST_CHECK_NCCL(api().CommInitRank(&comm, world, id, rank));   // binds to the calling thread's current CUDA device
  *comm_out = comm;
  return ST_OK;
}

int st_allreduce_run(void* comm, float* buf, int64_t n, cudaStream_t stream) {
  ST_REQUIRE(comm != nullptr && (buf != nullptr || n == 0) && n >= 0, "st_allreduce_run: bad arguments");
  ST_TRY(need_api());
  if (n == 0) return ST_OK;
  ST_CHECK_NCCL(api().AllReduce(buf, buf, static_cast<size_t>(n), kNcclFloat, kNcclSum, comm, stream));
  return ST_OK;
}

int st_allreduce_broadcast(void* comm, float* buf, int64_t n, int root, cudaStream_t stream) {
  ST_REQUIRE(comm != nullptr && (buf != nullptr || n == 0) && n >= 0 && root >= 0, "st_allreduce_broadcast: bad arguments");
  ST_TRY(need_api());
  if (n == 0) return ST_OK;
  ST_CHECK_NCCL(api().Broadcast(buf, buf, static_cast<size_t>(n), kNcclFloat, root, comm, stream));
  return ST_OK;
}

int st_allreduce_destroy(void* comm) {
  if (comm == nullptr) return ST_OK;
  ST_TRY(need_api());
  ST_CHECK_NCCL(api().CommDestroy(comm));
  return ST_OK;
}

}  // extern "C"
