// st_attn.cu — placeholder until the tcgen05 attention kernels land (next commit).
#include "st_host.h"
#include "st_kernels.h"

namespace st {
int attn_fwd(cudaStream_t, const AttnArgs&) {
  set_error("attn_fwd: not built yet");
  return ST_ERR_INVALID;
}
int attn_bwd(cudaStream_t, const AttnBwdArgs&) {
  set_error("attn_bwd: not built yet");
  return ST_ERR_INVALID;
}
}  // namespace st
