// k_ypass_k1_v7: k_ypass_k1_v6 (integer tensor-pipe Y pass, stages handed over through mbarriers by a producer warp) with the stages
// loaded as 2-D TENSOR copies: one (32 rows x 128 columns) box of 4 KB per request, 12 requests per stage instead of 32 row copies of
// 1.5 KB -- the row-copy versions of the three-stage ring were bound by the copy issue rate (profiles/r02_notes.md section 3c) --, written
// with the 128-byte swizzle (ldmatrix reads 8 rows x 16 bytes of a box from 8 distinct bank groups without the 16-byte row skew), rows and
// columns outside the matrix zero-filled by the copy engine.  Everything else is k_ypass_k1_v6.  Real build only (tensor maps come from the
// driver entry point that kernels_tc.cuh resolves); the emulated build substitutes tests/cuda_emul/kernels_ypass_tma_stub.h.
#pragma once
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
};
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

__global__ void __launch_bounds__(kY6Threads, 2)
k_ypass_k1_v7(const __grid_constant__ CUtensorMap tmY, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  CA_DYNAMIC_SMEM(unsigned char, ring7raw);
  unsigned char* ring6 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ring7raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1 KB
  constexpr int kWarpCols = 256, kKB = 8, kGB = 16;
  unsigned char* ring = ring6;                                                                   // [stage][consumer][box][32 rows][128 bytes], 128-byte swizzle
  unsigned char* p0 = ring6 + (size_t)kY6Stages * kY7StageBytes;
  uint2* psd = reinterpret_cast<uint2*>(p0);                                                      // [32-row block][digit][t]
  uint2* bws = reinterpret_cast<uint2*>(p0 + (size_t)(kY5MaxRows / 32) * 128);                    // [warp][kb][lane]: W digit fragments
  int* rsum = reinterpret_cast<int*>(p0 + (size_t)(kY5MaxRows / 32) * 128 + (size_t)kY6Consumers * 8 * 32 * 8);   // [stage][row][digit]
  uint64_t* full = reinterpret_cast<uint64_t*>(rsum + kY6Stages * kY5StageRows * 4);
  uint64_t* empty = full + kY6Stages;
  __shared__ float sred[8];
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
  const int64_t ntiles = (int64_t)nCB * nRB;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int cb = (int)(tile % nCB);
    const int64_t rb = tile / nCB;
    const int64_t tcol0 = (int64_t)cb * kY6Cols;
    const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
    const int nrows = (int)(rend - rbeg);
    const int nstages = (nrows + kY5StageRows - 1) / kY5StageRows;
    auto issue = [&](int sg) {                                       // producer warp: stage sg of this tile -> slot (j0 + sg) % kY6Stages
      if (sg >= nstages) return;
      const int st = (int)((j0 + (uint32_t)sg) % kY6Stages);
      const int r0 = sg * kY5StageRows;
      // one (32 rows x 128 columns) box per lane: 2 per consumer warp; rows / columns outside the matrix arrive as zeros and count as bytes
      if (lane == 0) bar_arm(full + st, (uint32_t)kY7StageBytes);
      __syncwarp();
      if (lane < 2 * kY6Consumers)
        ptx::tma_load_2d(ptx::smem_u32(ring + (size_t)st * kY7StageBytes + (size_t)lane * kY7BoxBytes), &tmY, (int)(tcol0 + lane * 128), (int)(rbeg + r0),
                         ptx::smem_u32(full + st));
    };
    if (producer)                                                    // (the slots were released when the previous tile was finalised)
      for (int sg = 0; sg < kY6Stages; ++sg) issue(sg);
    CA_SYNC_AFTER_SYNCHRONOUS_COPY();
    // ---- per-tile operands: scales, W digit fragments, psi digit table (all threads) ----
    const float kInf = __int_as_float(0x7f800000);
    float wm = 0.f, pm = 0.f;
    for (int c = tid; c < kY6Cols; c += kY6Threads) {
      const int64_t col = tcol0 + c;
      if (col < G) { const float v = fabsf(Vm[col]); wm = (v <= 3.0e38f) ? fmaxf(wm, v) : kInf; }
    }
    for (int r = tid; r < nrows; r += kY6Threads) { const float v = fabsf(U[rbeg + r]); pm = (v <= 3.0e38f) ? fmaxf(pm, v) : kInf; }
    wm = warp_max(wm); pm = warp_max(pm);
    if (lane == 0) sred[wid] = wm;
    __syncthreads();
    wm = sred[0];
#pragma unroll
    for (int i = 1; i < kY6Consumers + 1; ++i) wm = fmaxf(wm, sred[i]);
    __syncthreads();
    if (lane == 0) sred[wid] = pm;
    __syncthreads();
    pm = sred[0];
#pragma unroll
    for (int i = 1; i < kY6Consumers + 1; ++i) pm = fmaxf(pm, sred[i]);
    const float sw = y5_pow2_ceil(wm), sp = y5_pow2_ceil(pm);
    const float isw = 1.f / sw, isp = 1.f / sp;
    const bool bad = !(wm <= 3.0e38f) || !(pm <= 3.0e38f);
    if (!producer) {
#pragma unroll
      for (int kb = 0; kb < kKB; ++kb) {
        uint32_t w2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pk = 0u;
          if (g < 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t col = tcol0 + wid * kWarpCols + kb * 32 + h * 16 + t * 4 + j;
              const int D = col < G ? y5_digit(Vm[col] * isw, g) : 0;
              pk |= ((uint32_t)D & 0xffu) << (8 * j);
            }
          }
          w2[h] = pk;
        }
        bws[(wid * kKB + kb) * 32 + lane] = make_uint2(w2[0], w2[1]);
      }
    }
    for (int it = tid; it < nstages * 16; it += kY6Threads) {
      const int k = it >> 4, d = (it >> 2) & 3, tt = it & 3;
      uint32_t w2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = k * 32 + h * 16 + (j >> 1) * 8 + 2 * tt + (j & 1);
          const int D = r < nrows ? y5_digit(U[rbeg + r] * isp, d) : 0;
          pk |= ((uint32_t)D & 0xffu) << (8 * j);
        }
        w2[h] = pk;
      }
      psd[it] = make_uint2(w2[0], w2[1]);
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
        const double v = (double)d4.x * 0.015625 + (double)d4.y * 0.0001220703125 + (double)d4.z * 9.5367431640625e-07 +
                         (double)d4.w * 7.450580596923828e-09;
        if (sg * kY5StageRows + lane < nrows)
          rowpart[(int64_t)cb * N + rbeg + sg * kY5StageRows + lane] = bad ? __int_as_float(0x7fc00000) : (float)(v * (double)sw);
        __syncwarp();
        issue(sg + kY6Stages);
      }
    } else {
      int cacc[kGB][4];
#pragma unroll
      for (int gb = 0; gb < kGB; ++gb)
#pragma unroll
        for (int i = 0; i < 4; ++i) cacc[gb][i] = 0;
      for (int sg = 0; sg < nstages; ++sg) {
        const uint32_t j = j0 + (uint32_t)sg;
        const int st = (int)(j % kY6Stages);
        bar_wait(full + st, (j / kY6Stages) & 1u);
        const unsigned char* sbase = ring + (size_t)st * kY7StageBytes + (size_t)wid * 2 * kY7BoxBytes;   // this warp's two boxes
        int racc[2][4];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
          for (int i = 0; i < 4; ++i) racc[rh][i] = 0;
          const int arow = rh * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const unsigned char* ap = sbase + (size_t)arow * 128;
#pragma unroll
          for (int kb = 0; kb < kKB; ++kb) {
            uint32_t a[4];
            // box kb / 4; 16-byte chunk (kb % 4) * 2 + (lane >> 4) of the row, XOR-swizzled with the row (CU_TENSOR_MAP_SWIZZLE_128B)
            ldmatrix_x4(a, ap + (kb >> 2) * kY7BoxBytes + (((((kb & 3) << 1) | (lane >> 4)) ^ (arow & 7)) << 4));
            const uint2 b = bws[(wid * kKB + kb) * 32 + lane];
            mma_u8s8(racc[rh], a, b.x, b.y);
          }
        }
        const uint2 pb = (g < 4) ? psd[sg * 16 + g * 4 + t] : make_uint2(0u, 0u);
        const unsigned char* tp = sbase + (size_t)lane * 128;
#pragma unroll
        for (int gb = 0; gb < kGB; ++gb) {
          uint32_t r[4], a[4];
          ldmatrix_x4_trans(r, tp + (gb >> 3) * kY7BoxBytes + (((gb & 7) ^ (lane & 7)) << 4));
          a[0] = __byte_perm(r[0], r[1], 0x6420u);
          a[1] = __byte_perm(r[0], r[1], 0x7531u);
          a[2] = __byte_perm(r[2], r[3], 0x6420u);
          a[3] = __byte_perm(r[2], r[3], 0x7531u);
          mma_u8s8(cacc[gb], a, pb.x, pb.y);
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
#pragma unroll
      for (int gb = 0; gb < kGB; ++gb) {
        const double s0 = t == 0 ? 0.015625 : (t == 1 ? 9.5367431640625e-07 : 0.0);
        const double s1 = t == 0 ? 0.0001220703125 : (t == 1 ? 7.450580596923828e-09 : 0.0);
        double ev = (double)cacc[gb][0] * s0 + (double)cacc[gb][1] * s1;
        double od = (double)cacc[gb][2] * s0 + (double)cacc[gb][3] * s1;
        ev += __shfl_xor_sync(CA_FULL, ev, 1);
        od += __shfl_xor_sync(CA_FULL, od, 1);
        if (t == 0) {
          const int64_t col = tcol0 + wid * kWarpCols + gb * 16 + 2 * g;
          if (col < G) colpart[rb * G + col] = bad ? __int_as_float(0x7fc00000) : (float)(ev * (double)sp);
          if (col + 1 < G) colpart[rb * G + col + 1] = bad ? __int_as_float(0x7fc00000) : (float)(od * (double)sp);
        }
      }
    }
    j0 += (uint32_t)nstages;
    __syncthreads();                                                 // the tile is finalised: slots, sred, psd, bws are reused
  }
}

inline void y7_launch(const Y7Plan& p, unsigned grid, cudaStream_t st, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* U,
                      const float* Vm, float* rowpart, float* colpart) {
  k_ypass_k1_v7<<<grid, kY6Threads, ypass7_smem_bytes(), st>>>(p.tm, ldY, N, G, RB, nCB, nRB, U, Vm, rowpart, colpart);
}
inline cudaError_t y7_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(k_ypass_k1_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ypass7_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_ypass_k1_v7, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

}  // namespace ca
