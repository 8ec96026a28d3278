// CPU emulation of the CUDA execution model and of the small part of the CUDA runtime that
// clonealign_b200/csrc/core.cu uses.  TEST INFRASTRUCTURE ONLY (tests/test_cuda_emul.py, tests/test_emul_*.py):
// the product (clonealign_b200/) never includes, links or loads anything from this directory; it exists so that the
// kernels that need no tensor core / TMA (everything except kernels_tc.cuh) can be checked FUNCTIONALLY against the
// oracle in a container without a GPU: index math, reductions, barrier placement, reads of uninitialised memory.
// Hardware hazards (memory model, occupancy, performance) are of course not covered.
//
// Model: every CUDA thread of a block is a fiber (ucontext) on ONE OS thread; fibers run until they reach a
// synchronisation point (__syncthreads, __syncwarp, warp shuffles) and are resumed once the barrier generation
// advances.  Threads that return early count as arrived (as on hardware).  A block in which no fiber can make
// progress aborts with a message (mismatched barriers).  Blocks of a grid are distributed over a few OS threads;
// `__shared__` maps to `static thread_local`, which fibers of a block share because they never migrate.
// Device memory is host memory: cudaMalloc fills it with 0xFF bytes (NaN floats) so that a kernel consuming memory
// nobody wrote shows up as NaN in the parity tests.
#pragma once
#include <ucontext.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------------------------
// vector types, qualifiers
// ------------------------------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
inline double2 make_double2(double a, double b) { return {a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline float2 make_float2(float a, float b) { return {a, b}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static thread_local

// Context switch between fibers.  swapcontext() saves/restores the signal mask with a system call on every switch
// (most of the emulation's run time); on x86-64 a ten-instruction switch of the callee-saved registers is used instead.
#if defined(__x86_64__) && !defined(CA_EMUL_UCONTEXT)
#define CA_EMUL_ASM_SWITCH 1
extern "C" void ca_emul_switch(void** save_sp, void* const* load_sp);
asm(R"(
.text
.weak ca_emul_switch
.type ca_emul_switch,@function
ca_emul_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq (%rsi), %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size ca_emul_switch,.-ca_emul_switch
)");
#endif

namespace ca_emul {

constexpr size_t kStackBytes = 256 * 1024;

struct Warp {
  int live = 0, count = 0;
  unsigned gen = 0;
  uint64_t slot[2][32];
};
struct Fiber {
#ifdef CA_EMUL_ASM_SWITCH
  void* sp = nullptr;
#else
  ucontext_t ctx;
#endif
  char* stack = nullptr;
  int tid = 0;
  bool done = false;
  int wait_kind = 0;        // 0 runnable, 1 block barrier, 2 warp barrier
  unsigned wait_gen = 0;
};
struct Block {
  int nthreads = 0, live = 0, bar_count = 0;
  unsigned bar_gen = 0;
  std::vector<Warp> warps;
  std::vector<unsigned char> dyn;
  uint64_t progress = 0;
};

inline thread_local Block* blk = nullptr;
inline thread_local Fiber* fib = nullptr;
#ifdef CA_EMUL_ASM_SWITCH
inline thread_local void* sched_sp = nullptr;
#else
inline thread_local ucontext_t sched;
#endif
inline thread_local void (*entry)(void*) = nullptr;
inline thread_local void* entry_arg = nullptr;

#ifdef CA_EMUL_ASM_SWITCH
inline void yield_to_scheduler() { ca_emul_switch(&fib->sp, &sched_sp); }
#else
inline void yield_to_scheduler() { swapcontext(&fib->ctx, &sched); }
#endif
inline void* dyn_smem() { return blk->dyn.data(); }

inline void block_barrier() {
  Block* b = blk;
  Fiber* f = fib;
  const unsigned g = b->bar_gen;
  if (++b->bar_count >= b->live) {
    b->bar_count = 0;
    b->bar_gen++;
    b->progress++;
    return;
  }
  f->wait_kind = 1;
  f->wait_gen = g;
  yield_to_scheduler();
}
// returns the generation the caller took part in
inline unsigned warp_barrier() {
  Block* b = blk;
  Fiber* f = fib;
  Warp& w = b->warps[f->tid >> 5];
  const unsigned g = w.gen;
  if (++w.count >= w.live) {
    w.count = 0;
    w.gen++;
    b->progress++;
    return g;
  }
  f->wait_kind = 2;
  f->wait_gen = g;
  yield_to_scheduler();
  return g;
}
inline void thread_exit() {   // an exited thread counts as arrived at every later barrier
  Block* b = blk;
  Fiber* f = fib;
  Warp& w = b->warps[f->tid >> 5];
  f->done = true;
  b->progress++;
  b->live--;
  w.live--;
  if (b->live > 0 && b->bar_count >= b->live) { b->bar_count = 0; b->bar_gen++; }
  if (w.live > 0 && w.count >= w.live) { w.count = 0; w.gen++; }
}

}  // namespace ca_emul

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { ca_emul::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ca_emul::warp_barrier(); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  ca_emul::Warp& w = ca_emul::blk->warps[ca_emul::fib->tid >> 5];
  const int lane = ca_emul::fib->tid & 31;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  // double-buffered by barrier generation: nobody can write generation g+2 before everyone has read generation g
  const unsigned g = w.gen;
  w.slot[g & 1][lane] = bits;
  ca_emul::warp_barrier();
  uint64_t got = w.slot[g & 1][src & 31];
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int mask) { return __shfl_sync(m, v, (ca_emul::fib->tid & 31) ^ mask); }
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, int d) {
  int lane = ca_emul::fib->tid & 31;
  return __shfl_sync(m, v, lane + d < 32 ? lane + d : lane);
}
inline unsigned __ballot_sync(unsigned m, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (unsigned)(__shfl_sync(m, pred ? 1 : 0, l) != 0) << l;
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979f * x); *c = cosf(3.14159265358979f * x); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline int __double2hiint(double d) { long long v; std::memcpy(&v, &d, 8); return (int)(v >> 32); }
inline int __double2loint(double d) { long long v; std::memcpy(&v, &d, 8); return (int)(v & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) { long long v = ((long long)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &v, 8); return d; }
inline float __fdividef(float a, float b) { return a / b; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float __expf(float a) { return expf(a); }
inline float __logf(float a) { return logf(a); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> inline T __ldcv(const T* p) { T v; std::memcpy(&v, (const void*)p, sizeof(T)); return v; }
// a spin-wait on memory written by another block / rank: let the sibling fibers of this block run (on hardware the
// other lanes of the warp make progress independently) and give the OS thread away
inline void ca_emul_spin_pause() {
  ca_emul::blk->progress++;            // waiting on another OS thread is not a dead-lock of this block
  ca_emul::fib->wait_kind = 0;
  ca_emul::yield_to_scheduler();
  std::this_thread::yield();
}
#define CA_SPIN_PAUSE() ca_emul_spin_pause()
inline long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() * 2; }
struct __half { unsigned short v; };
struct __nv_bfloat16 { unsigned short v; };
inline __half __float2half_rn(float) { return {0}; }   // only the tensor path (not emulated) consumes halves
inline __nv_bfloat16 __float2bfloat16_rn(float f) { unsigned u; std::memcpy(&u, &f, 4); return {(unsigned short)((u + 0x7fffu + ((u >> 16) & 1)) >> 16)}; }
inline float __bfloat162float(__nv_bfloat16 b) { unsigned u = (unsigned)b.v << 16; float f; std::memcpy(&f, &u, 4); return f; }

#define CA_DYNAMIC_SMEM(T, name) T* name = reinterpret_cast<T*>(ca_emul::dyn_smem())

// ------------------------------------------------------------------------------------------------------------------
// kernel launch
// ------------------------------------------------------------------------------------------------------------------
namespace ca_emul {

inline void fiber_main() {
  entry(entry_arg);
  thread_exit();
#ifdef CA_EMUL_ASM_SWITCH
  void* dead;
  ca_emul_switch(&dead, &sched_sp);   // never resumed
  __builtin_trap();
#else
  setcontext(&sched);   // never returns here
#endif
}

// fiber stacks are recycled across launches through a process-wide free list (a 256 KB malloc is an mmap + page
// faults every time otherwise)
struct StackCache {
  std::mutex mu;
  std::vector<char*> free_list;
  char* take() {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!free_list.empty()) {
        char* p = free_list.back();
        free_list.pop_back();
        return p;
      }
    }
    return (char*)malloc(kStackBytes);
  }
  void give(std::vector<char*>& v) {
    std::lock_guard<std::mutex> lk(mu);
    for (char* p : v) free_list.push_back(p);
    v.clear();
  }
};
inline StackCache& stack_cache() { static StackCache c; return c; }
struct StackPool {
  std::vector<char*> all;
  ~StackPool() { stack_cache().give(all); }
  char* get(size_t i) {
    while (all.size() <= i) all.push_back(stack_cache().take());
    return all[i];
  }
};

inline void run_block(void (*body)(void*), void* body_arg, dim3 grid, dim3 block, size_t dyn_bytes, unsigned bx, unsigned by,
                      unsigned bz, StackPool& pool) {
  const int nthreads = (int)(block.x * block.y * block.z);
  Block b;
  b.nthreads = b.live = nthreads;
  b.warps.resize((nthreads + 31) / 32);
  for (int w = 0; w < (int)b.warps.size(); ++w) b.warps[w].live = std::min(32, nthreads - w * 32);
  b.dyn.assign(dyn_bytes + 256, 0xFF);   // 256-byte red zone behind the dynamic shared memory, checked after the block
  std::vector<Fiber> fibers(nthreads);
  blk = &b;
  blockIdx = {bx, by, bz};
  blockDim = block;
  gridDim = grid;
  entry = body;
  entry_arg = body_arg;
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = fibers[t];
    f.tid = t;
    f.stack = pool.get(t);
#ifdef CA_EMUL_ASM_SWITCH
    {
      // initial frame: six callee-saved register slots, then the entry point as the return address of the first switch;
      // at entry rsp must be 8 mod 16, exactly as after a call instruction
      uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
      void** sp = (void**)top;
      *--sp = nullptr;                       // fake return address of fiber_main (it never returns)
      *--sp = (void*)&fiber_main;
      for (int i = 0; i < 6; ++i) *--sp = nullptr;
      f.sp = sp;
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_main, 0);
#endif
  }
  // Scheduling order of the fibers between synchronisation points.  Correct CUDA code may not depend on it; a missing
  // __syncthreads / __syncwarp often only shows under another order, so CA_EMUL_ORDER=reverse|random re-runs the same
  // tests with the threads of a block visited last-to-first or in a fresh pseudo-random permutation every round.
  static const int order_mode = [] {
    const char* e = getenv("CA_EMUL_ORDER");
    return !e ? 0 : (!strcmp(e, "reverse") ? 1 : (!strcmp(e, "random") ? 2 : 0));
  }();
  std::vector<int> order(nthreads);
  for (int t = 0; t < nthreads; ++t) order[t] = order_mode == 1 ? nthreads - 1 - t : t;
  uint64_t rng = 0x9E3779B97F4A7C15ull ^ ((uint64_t)bx * 1315423911u + by * 2654435761u + bz);
  int remaining = nthreads;
  while (remaining > 0) {
    const uint64_t before = b.progress;
    bool ran = false;
    if (order_mode == 2)
      for (int i = nthreads - 1; i > 0; --i) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[i], order[(int)(rng % (uint64_t)(i + 1))]);
      }
    for (int oi = 0; oi < nthreads; ++oi) {
      const int t = order[oi];
      Fiber& f = fibers[t];
      if (f.done) continue;
      if (f.wait_kind == 1 && b.bar_gen == f.wait_gen) continue;
      if (f.wait_kind == 2 && b.warps[t >> 5].gen == f.wait_gen) continue;
      f.wait_kind = 0;
      fib = &f;
      threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      ran = true;
#ifdef CA_EMUL_ASM_SWITCH
      ca_emul_switch(&sched_sp, &f.sp);
#else
      swapcontext(&sched, &f.ctx);
#endif
      if (f.done) --remaining;
    }
    if (remaining > 0 && (!ran || b.progress == before)) {
      fprintf(stderr, "cuda_emul: dead-lock in block (%u,%u,%u): %d threads wait at barriers that can never complete\n",
              bx, by, bz, remaining);
      abort();
    }
  }
  for (size_t i = dyn_bytes; i < dyn_bytes + 256; ++i)
    if (b.dyn[i] != 0xFF) {
      fprintf(stderr, "cuda_emul: block (%u,%u,%u) wrote past its %zu bytes of dynamic shared memory (offset +%zu)\n", bx, by, bz,
              dyn_bytes, i - dyn_bytes);
      abort();
    }
  blk = nullptr;
  fib = nullptr;
}

inline int host_workers() {
  static int n = [] {
    const char* e = getenv("CA_EMUL_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    (void)v;
    return 64;   // the blocks of a spin-waiting kernel (k_p2p_allreduce, <= 64) must be co-resident, whatever the host has
  }();
  return n;
}

// Launch limits of the hardware (sm_100): a launch that the driver would reject with cudaErrorInvalidConfiguration /
// cudaErrorInvalidValue aborts here with a message, so that a grid that only gets too large at full problem sizes, or
// a kernel whose dynamic shared memory was raised past 48 KB without cudaFuncSetAttribute, is found without a device.
struct FuncAttrRegistry {
  std::mutex mu;
  std::vector<std::pair<const void*, size_t>> max_dyn;   // kernel -> cudaFuncAttributeMaxDynamicSharedMemorySize
  size_t get(const void* k) {
    std::lock_guard<std::mutex> lk(mu);
    size_t v = 48 * 1024;
    for (auto& e : max_dyn)
      if (e.first == k) v = e.second;
    return v;
  }
  void set(const void* k, size_t v) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : max_dyn)
      if (e.first == k) { e.second = v; return; }
    max_dyn.push_back({k, v});
  }
};
inline FuncAttrRegistry& func_attrs() { static FuncAttrRegistry r; return r; }
constexpr size_t kMaxOptinSmem = 227 * 1024;   // per-block opt-in maximum on sm_100

inline void check_launch_limits(const void* kernel, dim3 grid, dim3 block, size_t dyn_bytes) {
  const uint64_t nthreads = (uint64_t)block.x * block.y * block.z;
  const char* why = nullptr;
  if (grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u) why = "grid dimension over the hardware limit (x <= 2^31-1, y/z <= 65535)";
  else if (nthreads > 1024 || block.x > 1024 || block.y > 1024 || block.z > 64) why = "more than 1024 threads per block";
  else if (dyn_bytes > kMaxOptinSmem) why = "dynamic shared memory over the 227 KB opt-in maximum";
  else if (dyn_bytes > func_attrs().get(kernel)) why = "dynamic shared memory over 48 KB without a matching cudaFuncSetAttribute(MaxDynamicSharedMemorySize)";
  if (why) {
    fprintf(stderr, "cuda_emul: invalid launch configuration: %s [grid (%u,%u,%u) block (%u,%u,%u) smem %zu]\n", why, grid.x, grid.y,
            grid.z, block.x, block.y, block.z, dyn_bytes);
    abort();
  }
}

template <typename K, typename... A>
void launch(K kernel, dim3 grid, dim3 block, size_t dyn_bytes, A... args) {
  const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
  check_launch_limits((const void*)kernel, grid, block, dyn_bytes);
  if (nblocks == 0 || block.x * block.y * block.z == 0) {
    // a zero-sized grid or block is cudaErrorInvalidConfiguration on hardware
    fprintf(stderr, "cuda_emul: invalid launch configuration: empty grid or block [grid (%u,%u,%u) block (%u,%u,%u)]\n", grid.x, grid.y,
            grid.z, block.x, block.y, block.z);
    abort();
  }
  struct Call {
    K kernel;
    std::tuple<A...> tup;
  } call{kernel, std::make_tuple(args...)};
  void (*body)(void*) = [](void* c) {
    Call* k = static_cast<Call*>(c);
    std::apply(k->kernel, k->tup);
  };
  std::atomic<uint64_t> next{0};
  auto worker = [&]() {
    StackPool pool;
    for (;;) {
      uint64_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y), bz = (unsigned)(i / ((uint64_t)grid.x * grid.y));
      run_block(body, &call, grid, block, dyn_bytes, bx, by, bz, pool);
    }
  };
  const int nw = (int)std::min<uint64_t>(nblocks, (uint64_t)host_workers());
  if (nw <= 1) {
    worker();
  } else {
    std::vector<std::thread> ts;
    for (int i = 0; i < nw; ++i) ts.emplace_back(worker);
    for (auto& t : ts) t.join();
  }
}

// CA_LAUNCH(kernel, grid, block, smem, stream)(args...) under emulation
template <typename K>
struct Launcher {
  K kernel;
  dim3 grid, block;
  size_t smem;
  template <typename... A>
  void operator()(A... args) const { launch(kernel, grid, block, smem, args...); }
};
template <typename K>
Launcher<K> launcher(K kernel, dim3 grid, dim3 block, size_t smem) { return {kernel, grid, block, smem}; }

}  // namespace ca_emul

#define CA_LAUNCH(kernel, grid, block, smem, stream) ca_emul::launcher(kernel, dim3(grid), dim3(block), (size_t)(smem))

// ------------------------------------------------------------------------------------------------------------------
// the part of the CUDA runtime API that core.cu uses (synchronous, host memory)
// ------------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct ca_emul_stream* cudaStream_t;
struct ca_emul_event { std::chrono::steady_clock::time_point t; };
typedef ca_emul_event* cudaEvent_t;
// CUDA graphs are never built under emulation (platform_emul.h: kGraphsAvailable == false); the calls only have to compile
typedef void* cudaGraphExec_t;
typedef void* cudaGraph_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorInvalidValue; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorInvalidValue; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorInvalidValue; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorInvalidValue; }
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };

inline const char* cudaGetErrorName(cudaError_t e) { return e ? "cudaErrorEmulated" : "cudaSuccess"; }
inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated failure" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  const char* e = getenv("CA_EMUL_SMS");
  *p = {10, 0, e ? atoi(e) : 3};
  return cudaSuccess;
}
// Device memory = host memory with a 256-byte red zone on either side (0xFF like fresh memory, checked at cudaFree: an
// out-of-bounds WRITE near a buffer aborts with a message) and 0xFF-filled contents (reading memory nobody wrote, or the
// red zone, shows up as NaN in the parity tests).
namespace ca_emul {
constexpr size_t kRedZone = 256;
struct AllocRegistry {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> live;
};
inline AllocRegistry& alloc_registry() { static AllocRegistry r; return r; }
}  // namespace ca_emul
inline cudaError_t cudaMalloc(void** p, size_t bytes) {
  const size_t body = (bytes + 255) / 256 * 256;
  unsigned char* raw = (unsigned char*)aligned_alloc(256, body + 2 * ca_emul::kRedZone);
  if (!raw) return cudaErrorInvalidValue;
  memset(raw, 0xFF, ca_emul::kRedZone);
  memset(raw + ca_emul::kRedZone, 0xFF, body);
  memset(raw + ca_emul::kRedZone + bytes, 0xFF, body - bytes + ca_emul::kRedZone);
  *p = raw + ca_emul::kRedZone;
  std::lock_guard<std::mutex> lk(ca_emul::alloc_registry().mu);
  ca_emul::alloc_registry().live.push_back({*p, bytes});
  return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
inline cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  size_t bytes = 0;
  bool found = false;
  {
    std::lock_guard<std::mutex> lk(ca_emul::alloc_registry().mu);
    auto& v = ca_emul::alloc_registry().live;
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i].first == p) { bytes = v[i].second; v[i] = v.back(); v.pop_back(); found = true; break; }
  }
  if (!found) { fprintf(stderr, "cuda_emul: cudaFree of an unknown pointer %p\n", p); abort(); }
  const size_t body = (bytes + 255) / 256 * 256;
  unsigned char* raw = (unsigned char*)p - ca_emul::kRedZone;
  for (size_t i = 0; i < ca_emul::kRedZone; ++i)
    if (raw[i] != 0xFF) { fprintf(stderr, "cuda_emul: write BEFORE a %zu-byte device buffer (offset -%zu)\n", bytes, ca_emul::kRedZone - i); abort(); }
  for (size_t i = bytes; i < body + ca_emul::kRedZone; ++i)
    if (raw[ca_emul::kRedZone + i] != 0xFF) { fprintf(stderr, "cuda_emul: write PAST the end of a %zu-byte device buffer (offset +%zu)\n", bytes, i - bytes); abort(); }
  free(raw);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new ca_emul_event(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <typename K> inline cudaError_t cudaFuncSetAttribute(K k, cudaFuncAttribute a, int v) {
  if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) {
    if (v < 0 || (size_t)v > ca_emul::kMaxOptinSmem) return cudaErrorInvalidValue;
    ca_emul::func_attrs().set((const void*)k, (size_t)v);
  }
  return cudaSuccess;
}
// CUDA IPC between the "ranks" of the emulation (threads of one process): a handle is the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
