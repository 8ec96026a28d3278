// Emulation counterpart of clonealign_b200/csrc/platform.cuh (TEST INFRASTRUCTURE ONLY): functional stand-ins for the
// asynchronous-copy / mbarrier primitives, no CUDA graphs, no tensor-core kernels, an in-process stand-in for NCCL.
// Copies are synchronous: a cp.async / bulk copy is a memcpy by the issuing thread, commit / wait / mbarrier waits are
// no-ops, and CA_SYNC_AFTER_SYNCHRONOUS_COPY() is the block barrier that orders a one-thread copy before its readers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CA_TC_HEADER "kernels_tc_stub.h"
#define CA_Y7_HEADER "kernels_ypass_tma_stub.h"
#define CA_NCCL_PROVIDER "nccl_emul_provider.inl"
#define CA_SYNC_AFTER_SYNCHRONOUS_COPY() __syncthreads()

namespace ca {
constexpr bool kGraphsAvailable = false;
// the hardware faults on a copy whose addresses are not multiples of its size (cudaErrorMisalignedAddress): so does the emulation
inline void ca_emul_check_aligned(const void* a, const void* b, unsigned n, const char* what) {
  if (((uintptr_t)a | (uintptr_t)b) % n) { fprintf(stderr, "cuda_emul: misaligned %s (%p <- %p)\n", what, a, b); abort(); }
}
inline void cp_async16(void* smem_dst, const void* gmem_src) { ca_emul_check_aligned(smem_dst, gmem_src, 16, "cp.async 16"); std::memcpy(smem_dst, gmem_src, 16); }
inline void cp_async8(void* smem_dst, const void* gmem_src) { ca_emul_check_aligned(smem_dst, gmem_src, 8, "cp.async 8"); std::memcpy(smem_dst, gmem_src, 8); }
inline float rcp_approx(float x) { return 1.0f / x; }
// the integer Y pass (ldmatrix + mma.sync) is not emulated: the host code never selects it when kImmaAvailable is false
constexpr bool kImmaAvailable = false;
inline void ca_emul_no_imma() { fprintf(stderr, "cuda_emul: ldmatrix / mma.sync are not emulated\n"); abort(); }
inline void ldmatrix_x4(uint32_t (&)[4], const void*) { ca_emul_no_imma(); }
inline void ldmatrix_x4_trans(uint32_t (&)[4], const void*) { ca_emul_no_imma(); }
inline void mma_u8s8(int (&)[4], const uint32_t (&)[4], uint32_t, uint32_t) { ca_emul_no_imma(); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline void bar_init(uint64_t*, int) {}
inline void bar_arm(uint64_t*, uint32_t) {}
inline void bar_wait(uint64_t*, uint32_t) {}
inline void bar_arrive(uint64_t*) {}
inline void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t*) {
  ca_emul_check_aligned(dst, src, 16, "bulk copy");
  if (bytes % 16) { fprintf(stderr, "cuda_emul: bulk copy of %u bytes\n", bytes); abort(); }
  std::memcpy(dst, src, bytes);
}
inline void fence_bar_init() {}
inline void fence_proxy_async() {}
}  // namespace ca
