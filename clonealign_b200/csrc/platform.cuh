// The hardware-specific primitives of the library in one place: asynchronous copies (cp.async, cp.async.bulk / TMA engine),
// mbarriers, proxy fences, CUDA-graph availability, which tensor-core header and which NCCL binding are compiled in.
// This is the ONLY file of the product with a build switch: the test suite compiles the same sources for the host against
// a functional CPU emulation of the CUDA execution model (tests/cuda_emul/, test infrastructure, never shipped or loaded by
// the product), and that build substitutes tests/cuda_emul/platform_emul.h for the definitions below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#ifdef CA_EMULATE
#include "platform_emul.h"
#else

#define CA_TC_HEADER "kernels_tc.cuh"                 // tcgen05 / TMEM / TMA contraction kernels
#define CA_Y7_HEADER "kernels_ypass_tma.cuh"          // integer Y pass on 2-D tensor copies (needs the tensor-map types of the header above)
#define CA_NCCL_PROVIDER "nccl_dlopen.inl"            // NCCL resolved with dlopen at first multi-GPU use
#define CA_SYNC_AFTER_SYNCHRONOUS_COPY() ((void)0)     // bulk copies are asynchronous here: completion is an mbarrier phase

namespace ca {

constexpr bool kGraphsAvailable = true;

// ---- cp.async (LDGSTS): 8 / 16 bytes per request, completion by commit / wait groups -------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier + 1-D bulk copies (cp.async.bulk: the TMA engine), proxy fences -----------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (addresses and size multiples of 16)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
}  // namespace ptx

// hardware reciprocal (MUFU.RCP, 1 ulp; flushes denormals): for positive normal arguments
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- warp-level matrix primitives of the integer Y pass (k_ypass_k1_v5): ldmatrix + mma.sync m16n8k32 u8 x s8 -> s32 ---------
constexpr bool kImmaAvailable = true;
// four 8 x 8 tiles of 16-bit elements; lane l supplies the address of row (l % 8) of tile (l / 8) (16 bytes, 16-byte aligned)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ptx::smem_u32(smem_row)) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ptx::smem_u32(smem_row)) : "memory");
}
// C (16 x 8, s32) += A (16 x 32, u8, row-major fragment) * B (32 x 8, s8, column-major fragment)
__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// the same, on generic pointers (what the Y pass uses)
__device__ __forceinline__ void bar_init(uint64_t* bar, int count) { ptx::mbar_init(ptx::smem_u32(bar), (uint32_t)count); }
__device__ __forceinline__ void bar_arm(uint64_t* bar, uint32_t bytes) { ptx::mbar_expect_tx(ptx::smem_u32(bar), bytes); }
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) { ptx::mbar_wait(ptx::smem_u32(bar), parity); }
__device__ __forceinline__ void bar_arrive(uint64_t* bar) { ptx::mbar_arrive(ptx::smem_u32(bar)); }
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  ptx::bulk_load_1d(ptx::smem_u32(dst), src, bytes, ptx::smem_u32(bar));
}
__device__ __forceinline__ void fence_bar_init() { ptx::fence_barrier_init(); }
__device__ __forceinline__ void fence_proxy_async() { ptx::fence_proxy_async_smem(); }

}  // namespace ca
#endif
