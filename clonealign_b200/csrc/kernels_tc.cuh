// tcgen05 (5th-gen tensor core) contraction kernels for the reference's default model (K = 1, P = 0).
//
//   FWD  Zx[n][j]  = sum_g E_ng * Mx[g][j]     M-dim = 128 cells, N-dim = J = 2*SCp, K-dim = genes
//   BWD  dMx[g][j] = sum_n E_ng * Rx[n][j]     M-dim = 128 genes, N-dim = J,          K-dim = cells
//   E_ng = exp(psi_n w_g - m_n)   (random_fixed_effects, R/inference-tflow.R:243,280; never materialised)
//
// The A operand (E) is GENERATED on chip: 8 generator warps evaluate exp2 on the fly and store the 16-bit
// tile straight into shared memory in the canonical K-major SWIZZLE_128B UMMA layout; the B operand
// (Mx^T / Rx^T, bf16, K-major) arrives by TMA; one elected thread issues tcgen05.mma with the fp32
// accumulator in TMEM; the same 8 warps drain TMEM with tcgen05.ld in the epilogue.  A and B travel through
// two independent shared-memory rings (A is produced by compute, B by the copy engine) so that the scarce
// shared memory buys latency tolerance where it is needed.
//
// Precision (scripts/emulate_precision.py).  log Z is multiplied by the library size s_n ~ 1e3..1e4
// before the clone softmax, and d psi_n = (YW)_n - sum R Z' is a difference of two terms ~s_n |w| whose true
// value is O(sqrt(s_n)): both Z and Z' need ~2^-16 operands.  FWD therefore uses a 3-term bf16 split on every
// column (E_hi*M_hi + E_hi*M_lo + E_lo*M_hi, ~2^-17 relative, fp32-equivalent).  BWD feeds Adam-normalised gene
// gradients only: one fp16 x fp16 term (2^-12); R is rescaled per cell by a power of two that is folded into the
// generated A operand's exponent, so both operands stay inside the fp16 range.
//
// Column layout of Mx / Zx / Rx / dMx (J = 2*SCp, SCp = S*C rounded up to 16):
//   [0, SCp)      : (s,c) -> mu_sg L_gc            | Z      | R            | dM
//   [SCp, 2*SCp)  : (s,c) -> w_g mu_sg L_gc        | Z'     | psi_n R      | dM'
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "common.cuh"

namespace ca {

constexpr int kTcGenWarps = 16;                      // generator / epilogue warps: two groups of 8
constexpr int kTcGroupThreads = kTcGenWarps / 2 * 32; // threads per generator group (one group per 128-row sub-tile)
constexpr int kTcThreads = (kTcGenWarps + 2) * 32;   // + TMA warp + MMA warp
constexpr int kTcBM = 128;                           // UMMA M
constexpr int kTcBK = 64;                            // K elements per stage (one 128-byte swizzle row of 16-bit)
constexpr float kLog2e = 1.4426950408889634f;

constexpr bool kTcAvailable = true;   // false in the CPU-emulation stub (tests/cuda_emul/kernels_tc_stub.h)
struct TcPlan {
  bool ok = false;
  int dev = 0, num_sms = 0;
  int64_t N = 0, Nld = 0, Gld = 0;
  int G = 0, SCp = 0, J = 0;
  int fsplit = 1, nsplit = 1;          // K-dim splits of FWD (genes) and BWD (cells)
  int64_t genes_per_fsplit = 0, cells_per_split = 0;
  int fwd_na = 0, fwd_nb = 0, bwd_na = 0, bwd_nb = 0;   // ring depths
  size_t fwd_smem = 0, bwd_smem = 0;
  int tmem_cols = 256;
  int dbg = 0;
  alignas(64) CUtensorMap tm_mhi, tm_mlo, tm_rx;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier / bulk-copy / proxy-fence wrappers shared with the Y pass live in platform.cuh)
// ------------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], 16-bit inputs, fp32 accumulate, single CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(tm), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16-byte shared store.  No "memory" clobber on purpose: the compiler may hoist the (read-only) global loads of
// later chunk-tasks above it; ordering against the proxy fence / mbarrier arrive is kept by asm volatile.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace ptx

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1, layout type 2).
// Rows are 128 bytes (64 x 16-bit); 8-row groups are 1024 bytes apart (SBO); LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32, A/B formats (0 = F16, 1 = BF16), both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {   // lo half = a, round-to-nearest-even
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// the kernel.  FWD: rows = cells (row scalar psi_n, shift m_n per row), K index = genes (w_g).
//              BWD: rows = genes (row scalar w_g), K index = cells (psi_n and shift_bwd_n per K index).
//
// CTA tile = 256 rows = two 128-row sub-tiles that share every B stage (halves the L2 -> SM traffic of B and
// doubles the time a B stage lives, i.e. the latency the TMA ring can hide).  Generator warps form two groups of
// 8, one per sub-tile, each with its own A ring and its own barriers: the groups run out of phase, so while one
// group sits in its barrier / fence phase the other keeps the MUFU (ex2) pipe busy.
// ------------------------------------------------------------------------------------------------
struct TcArgs {
  const float* rowv;     // FWD: psi [N]   BWD: w [G]
  const float* kv;       // FWD: w [G]     BWD: psi [N]         (allocation padded by 64 elements: bulk copies of 64)
  const float* shift;    // FWD: m [N] (natural log units)   BWD: shift_bwd [Nld] = m*log2(e) - a_n
  float* out;            // FWD: Zx [fsplit][N][J]   BWD: dMx [nsplit][G][J]
  int64_t rows;          // valid rows (N or G)
  int64_t kdim;          // valid K extent (G or N)
  int64_t k_per_split;   // K elements handled per blockIdx.y (multiple of 64)
  int J, SCp;
  int na, nb;            // A ring depth PER GROUP / B ring depth
  int tmem_cols;         // 256 or 512
  int dbg;               // timing experiments only (CLONEALIGN_B200_TC_DBG): 1 = no ex2, 2 = no B loads, 4 = 1/4 MMA, 8 = no proxy fence
};

template <bool FWD>
__global__ void __launch_bounds__(kTcThreads, 1)
k_expgemm_tc(const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1, TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A rings [2 groups][na] x { hi 16K | (FWD) lo 16K }, B ring [nb] x { B0 J*128 | (FWD) B1 J*128 | K-side
  // scalars 1K }, then barriers.  The K-side scalars of a block (FWD: w_g; BWD: psi_n and shift_bwd_n; 64 floats each)
  // ride in the B ring: one bulk copy per block by the TMA warp instead of redundant global loads per thread.
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t a_tile = kTcBM * 128;
  constexpr uint32_t a_stage = a_tile * (FWD ? 2 : 1);
  const uint32_t b_tile = (uint32_t)a.J * 128;
  const uint32_t b_stage = b_tile * (FWD ? 2 : 1) + 1024;
  const uint32_t ks_off = b_tile * (FWD ? 2 : 1);   // offset of the K-side scalars inside a B stage
  const int na = a.na, nb = a.nb;
  uint8_t* b_base = base + (size_t)2 * na * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)nb * b_stage);
  // bars: full_a[2][na] | empty_a[2][na] | full_b[nb] | empty_b[nb] | acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * na + 2 * nb + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_u32 = ptx::smem_u32(base), b_u32 = ptx::smem_u32(b_base), bar_u32 = ptx::smem_u32(bars);
  auto full_a = [&](int t, int s) { return bar_u32 + 8u * (t * na + s); };
  auto empty_a = [&](int t, int s) { return bar_u32 + 8u * (2 * na + t * na + s); };
  auto full_b = [&](int s) { return bar_u32 + 8u * (4 * na + s); };
  auto empty_b = [&](int s) { return bar_u32 + 8u * (4 * na + nb + s); };
  const uint32_t acc_full = bar_u32 + 8u * (4 * na + 2 * nb);

  const int64_t row0 = (int64_t)blockIdx.x * (2 * kTcBM);
  const int64_t kbeg = (int64_t)blockIdx.y * a.k_per_split;
  int64_t kend = kbeg + a.k_per_split;
  if (kend > a.kdim) kend = a.kdim;
  const int nkb = kend > kbeg ? (int)((kend - kbeg + kTcBK - 1) / kTcBK) : 0;

  if (warp == kTcGenWarps + 1) {
    if (lane == 0) {
      for (int s = 0; s < 2 * na; ++s) {
        ptx::mbar_init(bar_u32 + 8u * s, kTcGenWarps / 2);    // full_a: one arrival per warp of the group
        ptx::mbar_init(bar_u32 + 8u * (2 * na + s), 1);      // empty_a
      }
      for (int s = 0; s < nb; ++s) {
        ptx::mbar_init(full_b(s), 1);
        ptx::mbar_init(empty_b(s), 1);
      }
      ptx::mbar_init(acc_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kTcGenWarps) {
    // ===================== A-operand generators: group t = sub-tile t =====================
    const int t = warp / (kTcGenWarps / 2);
    const int gt = tid - t * kTcGroupThreads;   // thread index inside the group, 0..255
    const int r = gt & (kTcBM - 1);             // row of the sub-tile owned by this thread
    const int cpar = gt >> 7;                   // 0/1: 16-byte chunks cpar, cpar+2, cpar+4, cpar+6 of the row
    const int64_t grow = row0 + t * kTcBM + r;
    float rv = 0.f, rsh = 0.f;
    if (grow < a.rows) {
      rv = a.rowv[grow] * kLog2e;
      if (FWD) rsh = a.shift[grow] * kLog2e;
    }
    const uint32_t a_row = a_u32 + (uint32_t)(t * na) * a_stage + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    // ring positions and phases are carried incrementally (runtime '%' and '/' cost ~40 instructions each)
    int st = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      ptx::mbar_wait(full_b(sb), pb);                                    // K-side scalars of this block have landed
      const uint32_t ks = b_u32 + (uint32_t)sb * b_stage + ks_off;
      // all exponentials of this thread's 4 chunk-tasks are evaluated BEFORE waiting for the A slot: the MUFU work of
      // block kb overlaps the tensor core still reading block kb - na; only the pack + store needs the slot
      float e[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t c = (uint32_t)(cpar + 2 * i);
        const float4 k0 = ptx::ld_shared_v4f(ks + c * 32);
        const float4 k1 = ptx::ld_shared_v4f(ks + c * 32 + 16);
        const float kvv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        if (FWD) {
#pragma unroll
          for (int j = 0; j < 8; ++j) e[i][j] = (a.dbg & 1) ? fmaf(rv, kvv[j], -rsh) : ptx::ex2(fmaf(rv, kvv[j], -rsh));
        } else {
          const float4 s0 = ptx::ld_shared_v4f(ks + 256 + c * 32);
          const float4 s1 = ptx::ld_shared_v4f(ks + 256 + c * 32 + 16);
          const float shv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) e[i][j] = (a.dbg & 1) ? fmaf(rv, kvv[j], -shv[j]) : ptx::ex2(fmaf(rv, kvv[j], -shv[j]));
        }
      }
      ptx::mbar_wait(empty_a(t, st), pa ^ 1u);                           // A slot free
      const uint32_t sA = a_row + (uint32_t)st * a_stage;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = ((uint32_t)(cpar + 2 * i) ^ sw) << 4;
        if (FWD) {
          uint4 hi, lo;
          hi.x = pack_bf16x2(e[i][0], e[i][1]); hi.y = pack_bf16x2(e[i][2], e[i][3]);
          hi.z = pack_bf16x2(e[i][4], e[i][5]); hi.w = pack_bf16x2(e[i][6], e[i][7]);
          lo.x = pack_bf16x2(e[i][0] - __uint_as_float(hi.x << 16), e[i][1] - __uint_as_float(hi.x & 0xffff0000u));
          lo.y = pack_bf16x2(e[i][2] - __uint_as_float(hi.y << 16), e[i][3] - __uint_as_float(hi.y & 0xffff0000u));
          lo.z = pack_bf16x2(e[i][4] - __uint_as_float(hi.z << 16), e[i][5] - __uint_as_float(hi.z & 0xffff0000u));
          lo.w = pack_bf16x2(e[i][6] - __uint_as_float(hi.w << 16), e[i][7] - __uint_as_float(hi.w & 0xffff0000u));
          ptx::st_shared_v4(sA + off, hi);
          ptx::st_shared_v4(sA + a_tile + off, lo);
        } else {
          uint4 h;
          h.x = pack_f16x2(e[i][0], e[i][1]); h.y = pack_f16x2(e[i][2], e[i][3]);
          h.z = pack_f16x2(e[i][4], e[i][5]); h.w = pack_f16x2(e[i][6], e[i][7]);
          ptx::st_shared_v4(sA + off, h);
        }
      }
      if (!(a.dbg & 8)) ptx::fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(full_a(t, st));    // one arrival per warp (256 same-address arrivals would serialise)
      if (++st == na) { st = 0; pa ^= 1u; }
      if (++sb == nb) { sb = 0; pb ^= 1u; }
    }
    // ===================== epilogue: TMEM -> registers -> global =====================
    if (nkb > 0) {
      ptx::mbar_wait(acc_full, 0);
      ptx::tc_fence_after();
    }
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int wg = (warp >> 2) & 1;         // the two warps sharing (sub-tile, quarter) alternate 16-column chunks
    const int64_t orow = row0 + t * kTcBM + q * 32 + lane;
    float* obase = a.out + ((int64_t)blockIdx.y * a.rows + orow) * a.J;
    for (int col = wg * 16; col < a.J; col += 32) {
      uint32_t v[16];
      if (nkb > 0) {
        ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * a.J + col), v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (orow < a.rows) {
        float4* o4 = reinterpret_cast<float4*>(obase + col);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
      }
    }
  } else if (warp == kTcGenWarps) {
    // ===================== TMA producer for the B operand + K-side scalars =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(empty_b(st), ph ^ 1u);
        const uint32_t sB = b_u32 + (uint32_t)st * b_stage;
        const int kx = (int)(kbeg + (int64_t)kb * kTcBK);
        if (a.dbg & 2) {
          ptx::mbar_expect_tx(full_b(st), FWD ? 256u : 512u);
        } else {
          ptx::mbar_expect_tx(full_b(st), ks_off + (FWD ? 256u : 512u));
          ptx::tma_load_2d(sB, &tm_b0, kx, 0, full_b(st));
          if (FWD) ptx::tma_load_2d(sB + b_tile, &tm_b1, kx, 0, full_b(st));
        }
        ptx::bulk_load_1d(sB + ks_off, a.kv + kx, 256u, full_b(st));
        if (!FWD) ptx::bulk_load_1d(sB + ks_off + 256u, a.shift + kx, 256u, full_b(st));
        if (++st == nb) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // FWD: A bf16 (hi, lo), B bf16 (hi, lo).  BWD: A fp16, B fp16 (mixed f16/bf16 operands trap as illegal).
      const uint32_t idesc = FWD ? umma_idesc(a.J, 1u, 1u) : umma_idesc(a.J, 0u, 0u);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(full_b(sb), pb);
        const uint32_t sB = b_u32 + (uint32_t)sb * b_stage;
        const uint64_t dB = umma_desc_k128(sB), dBlo = umma_desc_k128(sB + b_tile);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(full_a(t, sa), pa);
          ptx::tc_fence_after();
          const uint32_t sA = a_u32 + (uint32_t)(t * na + sa) * a_stage;
          const uint64_t dA = umma_desc_k128(sA), dAlo = umma_desc_k128(sA + a_tile);
          const uint32_t acc = tmem_base + (uint32_t)(t * a.J);
#pragma unroll
          for (int k = 0; k < ((a.dbg & 4) ? 1 : kTcBK / 16); ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);   // +32 bytes along K inside the swizzle atom
            ptx::umma_f16(acc, dA + adv, dB + adv, idesc, (kb | k) != 0 ? 1u : 0u);
            if (FWD) {
              ptx::umma_f16(acc, dA + adv, dBlo + adv, idesc, 1u);   // E_hi * M_lo
              ptx::umma_f16(acc, dAlo + adv, dB + adv, idesc, 1u);   // E_lo * M_hi
            }
          }
          ptx::umma_commit(empty_a(t, sa));   // A slot of this group is free once these MMAs have read it
        }
        ptx::umma_commit(empty_b(sb));        // B stage is free once both sub-tiles have consumed it
        if (++sa == na) { sa = 0; pa ^= 1u; }
        if (++sb == nb) { sb = 0; pb ^= 1u; }
      }
      if (nkb > 0) ptx::umma_commit(acc_full);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kTcGenWarps + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline void tc_make_map(CUtensorMap* tm, void* gptr, uint64_t inner, uint64_t outer, uint32_t box_outer, bool fp16 = false) {
  static tc_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    fn = (tc_encode_fn)p;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};   // bytes, row pitch
  cuuint32_t box[2] = {kTcBK, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
}

inline int tc_pick_split(int64_t tiles, int64_t kchunks, int num_sms, int max_split) {
  // choose the K-split that minimises tail-wave waste; prefer fewer splits on ties
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= max_split; ++ns) {
    if (kchunks / ns < 8 && ns > 1) break;
    int64_t total = tiles * ns;
    int64_t waves = (total + num_sms - 1) / num_sms;
    double eff = (double)total / (double)(waves * num_sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = ns; }
  }
  return best;
}

// ring depths: `na` A stages per generator group (two groups), `nb` B stages, inside ~220 KB
inline void tc_pick_rings(size_t a_stage, size_t b_stage, int& na, int& nb) {
  const size_t budget = 220 * 1024;
  na = 2;
  nb = (2 * na * a_stage < budget) ? (int)std::min<size_t>(6, (budget - 2 * na * a_stage) / b_stage) : 0;
  if (nb < 3) {
    na = 1;
    nb = (int)std::min<size_t>(6, (budget - 2 * na * a_stage) / b_stage);
  }
  if (nb < 2) throw std::runtime_error("tensor path: tile does not fit in shared memory");
}

inline void tc_plan_create(TcPlan& p, int dev, int64_t N, int64_t Nld, int G, int64_t Gld, int SCp, int J, __nv_bfloat16* MxT_hi,
                           __nv_bfloat16* MxT_lo, __half* RxT) {
  p.dev = dev; p.N = N; p.Nld = Nld; p.G = G; p.Gld = Gld; p.SCp = SCp; p.J = J;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) throw std::runtime_error("cudaGetDeviceProperties failed");
  p.num_sms = prop.multiProcessorCount;
  tc_make_map(&p.tm_mhi, MxT_hi, (uint64_t)Gld, (uint64_t)J, (uint32_t)J);
  tc_make_map(&p.tm_mlo, MxT_lo, (uint64_t)Gld, (uint64_t)J, (uint32_t)J);
  tc_make_map(&p.tm_rx, RxT, (uint64_t)Nld, (uint64_t)J, (uint32_t)J, true);
  const size_t fa = 2 * kTcBM * 128, fb = 2 * (size_t)J * 128 + 1024, ba = kTcBM * 128, bb = (size_t)J * 128 + 1024;
  tc_pick_rings(fa, fb, p.fwd_na, p.fwd_nb);
  tc_pick_rings(ba, bb, p.bwd_na, p.bwd_nb);
  p.fwd_smem = 2 * p.fwd_na * fa + p.fwd_nb * fb + 1024 + 512;
  p.bwd_smem = 2 * p.bwd_na * ba + p.bwd_nb * bb + 1024 + 512;
  p.tmem_cols = (2 * J > 256) ? 512 : 256;
  const int64_t rows_per_cta = 2 * kTcBM;
  int64_t ftiles = (N + rows_per_cta - 1) / rows_per_cta, btiles = (G + rows_per_cta - 1) / rows_per_cta;
  p.fsplit = tc_pick_split(ftiles, Gld / kTcBK, p.num_sms, 4);
  p.nsplit = tc_pick_split(btiles, Nld / kTcBK, p.num_sms, 16);
  p.genes_per_fsplit = ((Gld / kTcBK + p.fsplit - 1) / p.fsplit) * kTcBK;
  p.cells_per_split = ((Nld / kTcBK + p.nsplit - 1) / p.nsplit) * kTcBK;
  if (cudaFuncSetAttribute(k_expgemm_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.fwd_smem) != cudaSuccess ||
      cudaFuncSetAttribute(k_expgemm_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.bwd_smem) != cudaSuccess)
    throw std::runtime_error("tensor path: cannot raise the dynamic shared memory limit");
  const char* dbg = getenv("CLONEALIGN_B200_TC_DBG");
  p.dbg = dbg ? atoi(dbg) : 0;
  p.ok = true;
}
inline void tc_plan_destroy(TcPlan& p) { p.ok = false; }

// FWD: psi = U [N] (K == 1), w = V [G]; out Zx [fsplit][N][J]
inline void tc_launch_fwd(const TcPlan& p, const float* psi, const float* w, const float* shift, float* Zx, cudaStream_t st) {
  TcArgs a;
  a.rowv = psi; a.kv = w; a.shift = shift; a.out = Zx; a.rows = p.N; a.kdim = p.G; a.k_per_split = p.genes_per_fsplit;
  a.J = p.J; a.SCp = p.SCp; a.na = p.fwd_na; a.nb = p.fwd_nb; a.tmem_cols = p.tmem_cols; a.dbg = p.dbg;
  dim3 grid((unsigned)((p.N + 2 * kTcBM - 1) / (2 * kTcBM)), p.fsplit);
  k_expgemm_tc<true><<<grid, kTcThreads, p.fwd_smem, st>>>(p.tm_mhi, p.tm_mlo, a);
}
// BWD: out dMx [nsplit][G][J]; shift_bwd is written by the per-cell epilogue together with the scaled fp16 R^T
inline void tc_launch_bwd(const TcPlan& p, const float* psi, const float* w, const float* shift, float* dMx, cudaStream_t st) {
  TcArgs a;
  a.rowv = w; a.kv = psi; a.shift = shift; a.out = dMx; a.rows = p.G; a.kdim = p.N; a.k_per_split = p.cells_per_split;
  a.J = p.J; a.SCp = p.SCp; a.na = p.bwd_na; a.nb = p.bwd_nb; a.tmem_cols = p.tmem_cols; a.dbg = p.dbg;
  dim3 grid((unsigned)((p.G + 2 * kTcBM - 1) / (2 * kTcBM)), p.nsplit);
  k_expgemm_tc<false><<<grid, kTcThreads, p.bwd_smem, st>>>(p.tm_rx, p.tm_rx, a);
}

}  // namespace ca
