// Emulation counterpart of clonealign_b200/csrc/nccl_dlopen.inl (TEST INFRASTRUCTURE ONLY): the ranks of a sharded fit are
// threads of one process, see nccl_emul.h.  Included from core_support.inl inside its anonymous namespace.
}  // namespace
#include "nccl_emul.h"
namespace {
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  api.GetUniqueId = [](void* p) { return ca_emul_nccl::GetUniqueId(p); };
  api.CommInitRank = [](void** c, int w, Uid id, int r) { return ca_emul_nccl::CommInitRank(c, w, id, r); };
  api.AllReduce = [](const void* s, void* d, size_t n, int t, int o, void* c, cudaStream_t st) {
    return ca_emul_nccl::AllReduce(s, d, n, t, o, c, (void*)st);
  };
  api.CommDestroy = [](void* c) { return ca_emul_nccl::CommDestroy(c); };
  api.GetErrorString = [](int e) { return ca_emul_nccl::GetErrorString(e); };
  api.lib = (void*)&api;
  return api;
}
