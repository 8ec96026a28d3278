// Minimal CPU emulation of the CUDA execution model for FUNCTIONAL tests of simple kernels (no tensor cores,
// no TMA): one OS thread per CUDA thread, blocks run one after another, __syncthreads / warp shuffles via
// std::barrier.  Test infrastructure only (tests/test_cuda_emul.py); it lets the index math, reductions and panel
// logic of clonealign_b200/csrc/kernels_interp.cuh be checked without a GPU.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

namespace ca_emul {
struct BlockCtx {
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;   // one per warp
  std::vector<uint64_t> wbuf;                          // 32 slots per warp
  std::vector<unsigned char> dyn;
  BlockCtx(int nthreads, size_t dyn_bytes) : bar(nthreads), wbuf((size_t)((nthreads + 31) / 32) * 32), dyn(dyn_bytes + 64) {
    for (int w = 0; w < (nthreads + 31) / 32; ++w) {
      int lanes = std::min(32, nthreads - w * 32);
      wbar.emplace_back(new std::barrier<>(lanes));
    }
  }
};
inline thread_local BlockCtx* ctx = nullptr;
inline thread_local int linear_tid = 0;
inline void* dyn_smem() { return ctx->dyn.data(); }
}  // namespace ca_emul

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { ca_emul::ctx->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ca_emul::ctx->wbar[ca_emul::linear_tid / 32]->arrive_and_wait(); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  auto* c = ca_emul::ctx;
  const int w = ca_emul::linear_tid / 32, lane = ca_emul::linear_tid % 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  c->wbuf[(size_t)w * 32 + lane] = bits;
  c->wbar[w]->arrive_and_wait();
  uint64_t got = c->wbuf[(size_t)w * 32 + (src & 31)];
  c->wbar[w]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int mask) { return __shfl_sync(m, v, (ca_emul::linear_tid % 32) ^ mask); }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979f * x); *c = cosf(3.14159265358979f * x); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline float __fdividef(float a, float b) { return a / b; }
struct __half { unsigned short v; };
struct __nv_bfloat16 { unsigned short v; };
inline __half __float2half_rn(float) { return {0}; }
inline __nv_bfloat16 __float2bfloat16_rn(float f) { unsigned u; std::memcpy(&u, &f, 4); return {(unsigned short)((u + 0x7fffu + ((u >> 16) & 1)) >> 16)}; }
inline float __bfloat162float(__nv_bfloat16 b) { unsigned u = (unsigned)b.v << 16; float f; std::memcpy(&f, &u, 4); return f; }

#define CA_DYNAMIC_SMEM(T, name) T* name = reinterpret_cast<T*>(ca_emul::dyn_smem())

namespace ca_emul {
// launch<<<grid, block, dyn_smem>>>: blocks sequentially, threads of a block concurrently
template <typename K, typename... A>
void launch(K kernel, dim3 grid, dim3 block, size_t dyn_bytes, A... args) {
  const int nthreads = (int)(block.x * block.y * block.z);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        BlockCtx c(nthreads, dyn_bytes);
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
          ts.emplace_back([&, t]() {
            ctx = &c;
            linear_tid = t;
            threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = {bx, by, bz};
            blockDim = block;
            gridDim = grid;
            kernel(args...);
            // a thread that returns early must not dead-lock the others: drop out of the block barrier
            c.bar.arrive_and_drop();
            c.wbar[t / 32]->arrive_and_drop();
          });
        for (auto& th : ts) th.join();
      }
}
}  // namespace ca_emul
