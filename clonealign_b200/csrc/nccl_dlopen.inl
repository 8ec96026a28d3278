// NCCL binding of the product: libnccl resolved with dlopen on first multi-GPU use (no link-time dependency; a single-GPU
// fit never touches it).  Included from core_support.inl inside its anonymous namespace (see platform.cuh: CA_NCCL_PROVIDER).
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) fail("NCCL is required for world > 1 but libnccl.so.2 could not be loaded: %s", dlerror());
  auto sym = [&](const char* s) {
    void* p = dlsym(api.lib, s);
    if (!p) fail("NCCL symbol %s not found", s);
    return p;
  };
  api.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, Uid, int))sym("ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
  api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  return api;
}
