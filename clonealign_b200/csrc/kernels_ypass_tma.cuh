// k_ypass_k1_v7: the Y pass of the K = 1 step (row sums YW and column sums Y^T psi from ONE stream over the stored u8 counts) as exact
// integer contractions on the tensor pipe, fed by 2-D TENSOR copies.  What `path = auto` runs on u8 matrices of benchmark size
// (core_build.inl; profiles/r02_notes.md section 3d has the measurements behind every choice below).
//   * W and psi are cut into four base-128 s8 digits of a power-of-two tile scale (k_ypass_k1_v5 in kernels_ypass.cuh introduced the
//     arithmetic); products are mma.sync.m16n8k32 on u8 x s8 / s8 x u8 with s32 accumulators, digits recombined in fp64 in a fixed order:
//     the results do not depend on the order of accumulation or on which CTA ran a tile.
//   * One persistent CTA per SM: a producer warp and six consumer warps (128 registers: fits next to the 16 x 64-register kernels of the
//     co-scheduled step).  A stage is 32 rows x up to 12 boxes of 128 columns, one cp.async.bulk.tensor request of 4 KB per box (12 requests
//     instead of the 32 row copies that bound the row-copy versions by their issue rate), written with the 128-byte swizzle so that
//     ldmatrix reads 8 rows x 16 bytes of a box from 8 distinct bank groups; rows outside the matrix are zero-filled by the copy engine.
//     Three stages, handed over through full / empty mbarriers only; the producer walks the stages of all the CTA's tiles as one stream.
//   * Row sums: A = the staged tile (ldmatrix), B = W digit fragments.  Column sums: B = the staged tile through ldmatrix.trans AS IT IS
//     (two rows of a column pair per register), A = psi digits with the other column parity's slots zeroed -- no byte transpose.
//   * The boxes of a row are dealt evenly over the column blocks; tiles are walked in storage order from per-CTA lists.
// Real build only (tensor maps come from the driver entry point that kernels_tc.cuh resolves); the emulated build substitutes
// tests/cuda_emul/kernels_ypass_tma_stub.h and never selects the kernel.
#pragma once
#include <vector>
#include "kernels_ypass.cuh"

namespace ca {

constexpr bool kY7Available = true;
constexpr int kY7BoxBytes = 32 * 128, kY7StageBytes = 2 * kY6Consumers * kY7BoxBytes;
inline size_t ypass7_smem_bytes() {
  return 1024 + (size_t)kY6Stages * kY7StageBytes + (size_t)(kY5MaxRows / 32) * 128 + (size_t)kY6Consumers * 8 * 32 * 8 +
         (size_t)kY6Stages * kY5StageRows * 16 + 8 * 2 * kY6Stages + 16;
}
struct Y7Plan {
  alignas(64) CUtensorMap tm;
  bool ok = false;
  int RB = 0;                      // rows of a tile (a multiple of the 32-row stage)
  std::vector<int> tiles, offs;    // tiles (rb * nCB + cb) of CTA b: tiles[offs[b] .. offs[b + 1])
  const int* d_tiles = nullptr;    // device copies (owned by the handle's allocator)
  const int* d_offs = nullptr;
};
// Tiles are (RB rows x one column block).  The ceil(ldY / 128) boxes of a row are dealt evenly over the nCB column blocks (config 3: 157
// boxes -> 3 blocks of 12 and 11 of 11, not 13 of 12 and one of 1: a stage costs a ring round trip whatever it holds, so a block of one
// box would cost most of a full one).  Every CTA walks its own list of tiles, handed over as [offsets | tiles]: a strided walk in storage
// order (the tiles in flight at any time cover whole rows of the matrix).  Longest-first / least-loaded dealing and a cost-model choice
// of RB were measured and lost to it by 2-4 % of the step (profiles/r02_notes.md section 3c).  Sums inside a tile are integers and the
// partials are added in fixed (rb, cb) order behind the pass, so results do not depend on which CTA ran a tile.
__host__ __device__ inline int y7_first_box(int cb, int nboxes, int nCB) { const int q = nboxes / nCB, r = nboxes % nCB; return cb * q + (cb < r ? cb : r); }
inline void y7_plan_tiles(Y7Plan& p, int64_t N, int64_t ldY, int grid, int RB) {
  const int nCB = (int)((ldY + kY6Cols - 1) / kY6Cols), nRB = (int)((N + RB - 1) / RB);
  p.RB = RB; p.tiles.clear(); p.offs.assign(1, 0);
  for (int b = 0; b < grid; ++b) {
    for (int t = b; t < nCB * nRB; t += grid) p.tiles.push_back(t);
    p.offs.push_back((int)p.tiles.size());
  }
}
// tensor map over the stored u8 matrix [N][ldY]: boxes of 128 columns x 32 rows, 128-byte swizzle, zero fill outside
inline void y7_plan_create(Y7Plan& p, const void* Y, int64_t N, int64_t ldY) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fp) throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {(cuuint64_t)ldY, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)ldY};
  cuuint32_t box[2] = {128, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((tc_encode_fn)fp)(&p.tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(Y), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (Y pass) failed (code " + std::to_string((int)r) + ")");
  p.ok = true;
}

// all four base-128 digits of x in [-1, 1] from one pass of y5_digit's recursion (the same roundings: identical digits)
__device__ __forceinline__ void y7_digits4(float x, int (&D)[4]) {
  float y = x * 64.f;
  float d = rintf(y);
  D[0] = (int)d;
#pragma unroll
  for (int i = 1; i < 4; ++i) { y = (y - d) * 128.f; d = rintf(y); D[i] = (int)d; }
}
// ldmatrix on shared-space addresses; products without `volatile` (pure functions of their operands: the compiler orders them by data flow)
__device__ __forceinline__ void y7_ldsm(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void y7_ldsm_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void y7_mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A (s8): rows 0..7 = (a0 | a2), rows 8..15 zero; B (u8)
__device__ __forceinline__ void y7_mma_s8u8_half(int (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %5}, {%7, %8}, {%0, %1, %2, %3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kY6Threads, 2)
k_ypass_k1_v7(const __grid_constant__ CUtensorMap tmY, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart, const int* __restrict__ tlist,
              const int* __restrict__ toffs) {
  CA_DYNAMIC_SMEM(unsigned char, ring7raw);
  // swizzle atoms are 1 KB: skip to the next multiple (an offset from the shared array, so that the accesses below stay shared-space ones)
  unsigned char* ring6 = ring7raw + ((1024u - (ptx::smem_u32(ring7raw) & 1023u)) & 1023u);
  constexpr int kWarpCols = 256, kKB = 8, kGB = 16;
  unsigned char* ring = ring6;                                                                   // [stage][consumer][box][32 rows][128 bytes], 128-byte swizzle
  unsigned char* p0 = ring6 + (size_t)kY6Stages * kY7StageBytes;
  uint2* psd = reinterpret_cast<uint2*>(p0);                                                      // [32-row block][digit][t]
  uint2* bws = reinterpret_cast<uint2*>(p0 + (size_t)(kY5MaxRows / 32) * 128);                    // [warp][kb][lane]: W digit fragments
  int* rsum = reinterpret_cast<int*>(p0 + (size_t)(kY5MaxRows / 32) * 128 + (size_t)kY6Consumers * 8 * 32 * 8);   // [stage][row][digit]
  uint64_t* full = reinterpret_cast<uint64_t*>(rsum + kY6Stages * kY5StageRows * 4);
  uint64_t* empty = full + kY6Stages;
  __shared__ float sred[16];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool producer = wid == kY6Consumers;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < kY6Stages * kY5StageRows * 4; i += kY6Threads) rsum[i] = 0;
  if (tid == 0)
    for (int st = 0; st < kY6Stages; ++st) { bar_init(full + st, 1); bar_init(empty + st, kY6Consumers); }
  fence_bar_init();
  fence_proxy_async();
  __syncthreads();
  uint32_t j0 = 0;                                                   // stages handed over so far (all tiles): stage j lives in slot j % kY6Stages
  const int tbeg = toffs[blockIdx.x], tend = toffs[blockIdx.x + 1];  // this CTA's tiles (y7_plan_tiles)
  // The producer walks the stages of ALL the CTA's tiles as one stream, kY6Stages ahead of the consumers: the ring stays full across tile
  // boundaries (the next tile's first stages land while this tile is finalised and the next one's operands are prepared).
  const int nboxes = (int)((ldY + 127) / 128);
  int pi = tbeg, psg = 0;                                            // next stage to request: tile tlist[pi], stage psg
  int pst = 0;                                                       // ... into slot pst
  int pnst = 0, nbox = 0, pcol = 0, prow = 0;                        // geometry of tile tlist[pi] (lane's box column, first row)
  auto next_tile = [&]() {
    if (pi >= tend) return;
    const int ptile = tlist[pi];
    const int pcb = ptile % nCB;
    const int pcol0 = y7_first_box(pcb, nboxes, nCB) * 128;
    const int64_t prbeg = (int64_t)(ptile / nCB) * RB;
    pnst = (int)(((prbeg + RB < N ? prbeg + RB : N) - prbeg + kY5StageRows - 1) / kY5StageRows);
    nbox = y7_first_box(pcb + 1, nboxes, nCB) - y7_first_box(pcb, nboxes, nCB);
    pcol = pcol0 + lane * 128;
    prow = (int)prbeg;
  };
  auto request = [&]() {
    if (pi >= tend) return;
    // one (32 rows x 128 columns) box per lane, 2 per consumer warp; rows outside the matrix arrive as zeros and count as bytes; boxes
    // past the tile's columns are not requested (their slots keep stale bytes: the W digits there are zero and no column sum is written)
    if (lane == 0) bar_arm(full + pst, (uint32_t)(nbox * kY7BoxBytes));
    __syncwarp();
    if (lane < nbox)
      ptx::tma_load_2d(ptx::smem_u32(ring + (size_t)pst * kY7StageBytes + (size_t)lane * kY7BoxBytes), &tmY, pcol, prow + psg * kY5StageRows,
                       ptx::smem_u32(full + pst));
    pst = (pst + 1 == kY6Stages) ? 0 : pst + 1;
    if (++psg == pnst) { psg = 0; ++pi; next_tile(); }
  };
  if (producer) next_tile();
  if (producer)
    for (int sg = 0; sg < kY6Stages; ++sg) request();
  for (int ti = tbeg; ti < tend; ++ti) {
    const int tile = tlist[ti];
    const int cb = tile % nCB;
    const int64_t rb = tile / nCB;
    const int64_t tcol0 = (int64_t)y7_first_box(cb, nboxes, nCB) * 128;
    const int64_t cend = ((int64_t)y7_first_box(cb + 1, nboxes, nCB) * 128 < G) ? (int64_t)y7_first_box(cb + 1, nboxes, nCB) * 128 : (int64_t)G;   // the tile's columns end here
    const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
    const int nrows = (int)(rend - rbeg);
    const int nstages = (nrows + kY5StageRows - 1) / kY5StageRows;
    // ---- per-tile operands: scales, W digit fragments, psi digit table (all threads) ----
    const float kInf = __int_as_float(0x7f800000);
    float wm = 0.f, pm = 0.f;
    for (int c = tid; c < kY6Cols; c += kY6Threads) {
      const int64_t col = tcol0 + c;
      if (col < cend) { const float v = fabsf(Vm[col]); wm = (v <= 3.0e38f) ? fmaxf(wm, v) : kInf; }
    }
    for (int r = tid; r < nrows; r += kY6Threads) { const float v = fabsf(U[rbeg + r]); pm = (v <= 3.0e38f) ? fmaxf(pm, v) : kInf; }
    wm = warp_max(wm); pm = warp_max(pm);
    if (lane == 0) { sred[wid] = wm; sred[8 + wid] = pm; }           // (read again only behind the barriers below)
    __syncthreads();
    wm = sred[0]; pm = sred[8];
#pragma unroll
    for (int i = 1; i < kY6Consumers + 1; ++i) { wm = fmaxf(wm, sred[i]); pm = fmaxf(pm, sred[8 + i]); }
    const float sw = y5_pow2_ceil(wm), sp = y5_pow2_ceil(pm);
    const float isw = 1.f / sw, isp = 1.f / sp;
    const bool bad = !(wm <= 3.0e38f) || !(pm <= 3.0e38f);
    if (!producer) {
      // W digit fragments [warp][kb][lane (g < 4)] = (columns 32 kb + 4 t .. + 3, the same + 16) of digit g; entries of lanes g >= 4 are
      // never read (those operands are 0).  A lane takes 4 columns of two k-blocks and writes their word in all four digits' entries.
      const int wt = lane & 3, wh = (lane >> 2) & 1, kq = lane >> 3;
      uint32_t* bw32 = reinterpret_cast<uint32_t*>(bws);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kb = 2 * kq + i;
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t col = tcol0 + wid * kWarpCols + kb * 32 + wh * 16 + wt * 4 + j;
          int D[4] = {0, 0, 0, 0};
          if (col < cend) y7_digits4(Vm[col] * isw, D);
#pragma unroll
          for (int d = 0; d < 4; ++d) pk[d] |= ((uint32_t)D[d] & 0xffu) << (8 * j);
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) bw32[((wid * kKB + kb) * 32 + d * 4 + wt) * 2 + wh] = pk[d];
      }
    }
    // psi digit table [32-row block k][digit d][tt] = (rows 2 tt, 2 tt + 1, 8 + 2 tt, 9 + 2 tt | the same + 16) of the block: an item is
    // two adjacent rows, written as a 16-bit half in all four digits' entries
    {
      unsigned short* psd16 = reinterpret_cast<unsigned short*>(psd);
      for (int it = tid; it < nstages * 16; it += kY6Threads) {
        const int k = it >> 4, h = (it >> 3) & 1, jp = (it >> 2) & 1, tt = it & 3;
        const int r0 = k * 32 + h * 16 + jp * 8 + 2 * tt;
        int D0[4] = {0, 0, 0, 0}, D1[4] = {0, 0, 0, 0};
        if (r0 < nrows) y7_digits4(U[rbeg + r0] * isp, D0);
        if (r0 + 1 < nrows) y7_digits4(U[rbeg + r0 + 1] * isp, D1);
#pragma unroll
        for (int d = 0; d < 4; ++d)
          psd16[((k * 16 + d * 4 + tt) * 2 + h) * 2 + jp] = (unsigned short)(((uint32_t)D0[d] & 0xffu) | (((uint32_t)D1[d] & 0xffu) << 8));
      }
    }
    __syncthreads();
    if (producer) {
      // ---- producer: finalise the row sums of every stage behind its consumers, then refill the slot ----
      for (int sg = 0; sg < nstages; ++sg) {
        const uint32_t j = j0 + (uint32_t)sg;
        const int st = (int)(j % kY6Stages);
        bar_wait(empty + st, (j / kY6Stages) & 1u);
        int4 d4 = reinterpret_cast<int4*>(rsum)[st * kY5StageRows + lane];
        reinterpret_cast<int4*>(rsum)[st * kY5StageRows + lane] = make_int4(0, 0, 0, 0);
        __syncwarp();
        request();                                                   // the released slot takes the next stage of the stream
        const double v = (double)d4.x * 0.015625 + (double)d4.y * 0.0001220703125 + (double)d4.z * 9.5367431640625e-07 +
                         (double)d4.w * 7.450580596923828e-09;
        if (sg * kY5StageRows + lane < nrows)
          rowpart[(int64_t)cb * N + rbeg + sg * kY5StageRows + lane] = bad ? __int_as_float(0x7fc00000) : (float)(v * (double)sw);
      }
    } else {
      // Column sums without a byte transpose: ldmatrix.trans hands thread (g, t) the bytes Y[2t][2g], Y[2t][2g+1], Y[2t+1][2g], Y[2t+1][2g+1]
      // of an (8 rows x 16 columns) block -- two rows of a column PAIR.  Taken as the B operand as it is, the contraction index of the
      // product is (row, column parity) and its n index the column pair; the A operand holds the psi digits with the other parity's slots
      // zeroed: row m = 2 * digit + parity of A is psi_digit[row] where the slot's parity is m's, else 0.  One product covers 16 rows x 16
      // columns (half of A is zeros) and lands digit d of column 16 gb + 4 t + 2 i + parity in accumulator i of lane (g = 2 d + parity, t);
      // rows 8..15 of A are zero, so accumulators 2 and 3 stay zero (the product writes aligned groups of four registers).
      int cacc[kGB][4];
#pragma unroll
      for (int gb = 0; gb < kGB; ++gb)
#pragma unroll
        for (int i = 0; i < 4; ++i) cacc[gb][i] = 0;
      const uint2* bwp = bws + (wid * kKB) * 32 + lane;              // W digit fragments of this warp's columns
      const uint32_t ring_u32 = ptx::smem_u32(ring) + (uint32_t)wid * 2u * kY7BoxBytes;
      const uint32_t x7 = (uint32_t)(lane & 7);
      // forward operand (rows x columns, k = columns): rows rh * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 16-byte chunk ((kb & 3) * 2 + (lane >> 4)) ^ (row & 7)
      const uint32_t aoff = (uint32_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * 128u, ahi = (uint32_t)(lane >> 4);
      const uint32_t toff = (uint32_t)lane * 128u;                   // transposed operand: row = lane, chunk (gb & 7) ^ (row & 7)
      const uint32_t selA = (g & 1) ? 0x1404u : 0x4140u, selB = (g & 1) ? 0x3424u : 0x4342u;
      auto load_batch = [&](uint32_t (&f)[4][4], uint32_t sb, int I) {   // kb = I for both row halves, gb = 2 I and 2 I + 1
        const uint32_t fo = (uint32_t)(I >> 2) * kY7BoxBytes + ((((uint32_t)(I & 3) << 1 | ahi) ^ x7) << 4);
        y7_ldsm(f[0], sb + aoff + fo);
        y7_ldsm(f[1], sb + aoff + 16u * 128u + fo);
        y7_ldsm_t(f[2], sb + toff + (uint32_t)((2 * I) >> 3) * kY7BoxBytes + ((((uint32_t)(2 * I) & 7u) ^ x7) << 4));
        y7_ldsm_t(f[3], sb + toff + (uint32_t)((2 * I + 1) >> 3) * kY7BoxBytes + ((((uint32_t)(2 * I + 1) & 7u) ^ x7) << 4));
      };
      for (int sg = 0; sg < nstages; ++sg) {
        const uint32_t j = j0 + (uint32_t)sg;
        const int st = (int)(j % kY6Stages);
        const uint2 pw = psd[sg * 16 + (g >> 1) * 4 + t];            // digit g / 2 of rows (2t, 2t+1, 8+2t, 9+2t | 16+.., 24+..)
        const uint32_t a10 = __byte_perm(pw.x, 0u, selA), a12 = __byte_perm(pw.x, 0u, selB);   // rows 0..15 of the stage
        const uint32_t a20 = __byte_perm(pw.y, 0u, selA), a22 = __byte_perm(pw.y, 0u, selB);   // rows 16..31
        bar_wait(full + st, (j / kY6Stages) & 1u);
        const uint32_t sb = ring_u32 + (uint32_t)st * kY7StageBytes;
        int racc[2][4];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh)
#pragma unroll
          for (int i = 0; i < 4; ++i) racc[rh][i] = 0;
        uint32_t f[2][4][4];
        load_batch(f[0], sb, 0);
#pragma unroll
        for (int I = 0; I < kKB; ++I) {                              // the next batch's loads are in flight while this one multiplies
          if (I + 1 < kKB) load_batch(f[(I + 1) & 1], sb, I + 1);
          uint32_t (&c)[4][4] = f[I & 1];
          const uint2 bw = (g < 4) ? bwp[I * 32] : make_uint2(0u, 0u);
          y7_mma_u8s8(racc[0], c[0], bw.x, bw.y);
          y7_mma_u8s8(racc[1], c[1], bw.x, bw.y);
          y7_mma_s8u8_half(cacc[2 * I], a10, a12, c[2][0], c[2][1]);
          y7_mma_s8u8_half(cacc[2 * I + 1], a10, a12, c[3][0], c[3][1]);
          y7_mma_s8u8_half(cacc[2 * I], a20, a22, c[2][2], c[2][3]);
          y7_mma_s8u8_half(cacc[2 * I + 1], a20, a22, c[3][2], c[3][3]);
        }
        // digit sums of the stage's rows: lanes t = 0 hold digits (0, 1), t = 1 digits (2, 3); integer adds commute
        if (t < 2) {
          int* rs = rsum + st * kY5StageRows * 4 + 2 * t;
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            atomicAdd(rs + (rh * 16 + g) * 4, racc[rh][0]);
            atomicAdd(rs + (rh * 16 + g) * 4 + 1, racc[rh][1]);
            atomicAdd(rs + (rh * 16 + g + 8) * 4, racc[rh][2]);
            atomicAdd(rs + (rh * 16 + g + 8) * 4 + 1, racc[rh][3]);
          }
        }
        __syncwarp();
        if (lane == 0) bar_arrive(empty + st);
      }
      // digits of a column sit in lanes g = 2 d + parity: add them in the fixed order (d0 + d1) + (d2 + d3)
      const int dgt = g >> 1;
      const double sc = dgt == 0 ? 0.015625 : (dgt == 1 ? 0.0001220703125 : (dgt == 2 ? 9.5367431640625e-07 : 7.450580596923828e-09));
#pragma unroll
      for (int gb = 0; gb < kGB; ++gb) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          double v = (double)cacc[gb][i] * sc;
          v += __shfl_xor_sync(CA_FULL, v, 8);
          v += __shfl_xor_sync(CA_FULL, v, 16);
          const int64_t col = tcol0 + wid * kWarpCols + gb * 16 + 4 * t + 2 * i + (g & 1);
          if (g < 2 && col < cend) colpart[rb * G + col] = bad ? __int_as_float(0x7fc00000) : (float)(v * (double)sp);
        }
      }
    }
    j0 += (uint32_t)nstages;
    __syncthreads();                                                 // the tile is finalised: slots, sred, psd, bws are reused
  }
}

inline void y7_launch(const Y7Plan& p, unsigned grid, cudaStream_t st, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* U,
                      const float* Vm, float* rowpart, float* colpart) {
  k_ypass_k1_v7<<<grid, kY6Threads, ypass7_smem_bytes(), st>>>(p.tm, ldY, N, G, RB, nCB, nRB, U, Vm, rowpart, colpart, p.d_tiles, p.d_offs);
}
inline cudaError_t y7_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(k_ypass_k1_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ypass7_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_ypass_k1_v7, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

}  // namespace ca
