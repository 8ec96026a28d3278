// CUDA-core (fp32) contraction with an on-the-fly exponential operand:
//
//   FWD:  Zx[n][j]  = sum_g exp(U_n . V_g - m_n) * Mx[g][j]      (normaliser, R/inference-tflow.R:288-290)
//   BWD:  dMx[g][j] = sum_n exp(U_n . V_g - m_n) * Rx[n][j]      (its reverse-mode gradient)
//
// E = exp(eta - m) is never materialised (the reference materialises (S,G,C,N), :289).  This is the
// general path (any K + P <= 8, any S*C); the tcgen05 path in kernels_tc.cuh replaces it for the
// reference's default K = 1, P = 0 model.  Each output element is produced by one thread in a fixed
// order, so results are run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace ca {

template <bool FWD>
__global__ void __launch_bounds__(256)
k_expgemm(const float* __restrict__ rowP, const float* __restrict__ conP, const float* __restrict__ shift,
          const float* __restrict__ Bmat, float* __restrict__ out, int64_t Mdim, int64_t Kdim, int J, int ld, int KP) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float rowS[BM][kMaxKP];
  __shared__ float conS[BK][kMaxKP];
  __shared__ float shRow[BM];
  __shared__ float shCon[BK];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  for (int i = tid; i < BM * kMaxKP; i += 256) {
    int r = i / kMaxKP, kp = i % kMaxKP;
    rowS[r][kp] = (m0 + r < Mdim && kp < KP) ? rowP[(m0 + r) * KP + kp] : 0.f;
  }
  if (tid < BM) shRow[tid] = (FWD && m0 + tid < Mdim) ? shift[m0 + tid] : 0.f;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < Kdim; k0 += BK) {
    __syncthreads();   // previous tile fully consumed (also orders the rowS/shRow prologue)
    if (tid < BK * kMaxKP) {
      int kk = tid / kMaxKP, kp = tid % kMaxKP;
      conS[kk][kp] = (k0 + kk < Kdim && kp < KP) ? conP[(k0 + kk) * KP + kp] : 0.f;
    }
    if (tid < BK) shCon[tid] = (!FWD && k0 + tid < Kdim) ? shift[k0 + tid] : 0.f;
    for (int i = tid; i < BK * BN; i += 256) {
      int kk = i / BN, nn = i % BN;
      Bs[kk][nn] = (k0 + kk < Kdim && n0 + nn < J) ? Bmat[(k0 + kk) * ld + n0 + nn] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < BK * BM; i += 256) {
      int kk = i / BM, mm = i % BM;
      float eta = 0.f;
#pragma unroll
      for (int kp = 0; kp < kMaxKP; ++kp) eta = fmaf(rowS[mm][kp], conS[kk][kp], eta);
      float sh = FWD ? shRow[mm] : shCon[kk];
      As[kk][mm] = (k0 + kk < Kdim) ? expf(eta - sh) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= Mdim) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c < J) out[r * ld + c] = acc[i][j];
    }
  }
}

}  // namespace ca
