// Emulation counterpart of clonealign_b200/csrc/platform.cuh (TEST INFRASTRUCTURE ONLY): functional stand-ins for the
// asynchronous-copy / mbarrier primitives, no CUDA graphs, no tensor-core kernels, an in-process stand-in for NCCL.
// Copies are synchronous: a cp.async / bulk copy is a memcpy by the issuing thread, commit / wait / mbarrier waits are
// no-ops, and CA_SYNC_AFTER_SYNCHRONOUS_COPY() is the block barrier that orders a one-thread copy before its readers.
#pragma once
#include <cstring>

#define CA_TC_HEADER "kernels_tc_stub.h"
#define CA_NCCL_PROVIDER "nccl_emul_provider.inl"
#define CA_SYNC_AFTER_SYNCHRONOUS_COPY() __syncthreads()

namespace ca {
constexpr bool kGraphsAvailable = false;
inline void cp_async16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
inline void cp_async8(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 8); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline void bar_init(uint64_t*, int) {}
inline void bar_arm(uint64_t*, uint32_t) {}
inline void bar_wait(uint64_t*, uint32_t) {}
inline void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t*) { std::memcpy(dst, src, bytes); }
inline void fence_bar_init() {}
inline void fence_proxy_async() {}
}  // namespace ca
