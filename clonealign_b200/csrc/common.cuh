// Shared device helpers for the clonealign_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>
#include "platform.cuh"

#define CA_WARP 32
#define CA_FULL 0xffffffffu
// dynamic shared memory declaration (the CPU emulation in tests/cuda_emul/ overrides it)
#ifndef CA_DYNAMIC_SMEM
#define CA_DYNAMIC_SMEM(T, name) extern __shared__ T name[]
#endif
// every kernel launch of the host code goes through this macro: CA_LAUNCH(kernel, grid, block, smem, stream)(args...)
// (the CPU emulation of tests/cuda_emul/ overrides it to run the same launch sequence without a GPU)
#ifndef CA_LAUNCH
#define CA_LAUNCH(kernel, grid, block, smem, stream) kernel<<<(grid), (block), (smem), (stream)>>>
#endif

namespace ca {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr int kMaxKP = 8;     // K + P supported by the kernels (reference default: K = 1, P = 0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator: eps(s, g) depends only on (seed, draw, s, g), so every
// rank of a cell-sharded run draws the same gene-level noise without communication.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// one N(0,1) draw for (draw, s, g) via Box-Muller on two 24-bit uniforms
__device__ __forceinline__ float normal_draw(uint64_t seed, uint64_t draw, uint32_t s, uint32_t g) {
  uint4 r = philox4x32_10(make_uint4(g, s, (uint32_t)draw, (uint32_t)(draw >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return rad * cs;
}

// ---------------------------------------------------------------------------------------------
// fixed-order reductions (run-to-run deterministic: tests/testthat/test_clonealign.R:42-66
// requires identical results for identical seeds, so no floating-point atomics anywhere)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CA_FULL, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(CA_FULL, v, o);
    v = w > v ? w : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(CA_FULL, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// block-wide sum; result valid in thread 0.  `scratch` holds >= 32 T.  Block size multiple of 32.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();   // scratch may still be in use by a previous call
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = T(0);
  if (threadIdx.x == 0)
    for (int i = 0; i < nw; ++i) r += scratch[i];
  return r;
}

__device__ __forceinline__ float softplusf(float x) {
  return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace ca
