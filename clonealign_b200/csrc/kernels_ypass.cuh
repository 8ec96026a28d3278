// K3: the one pass over the count matrix Y per parameter update.
//
// The multinomial log-probability sum_g y_ng log pi_scng (tfd$Multinomial$log_prob, R/inference-tflow.R
// :294-296) is linear in Y; after factorisation Y enters an iteration only through
//     YV  = Y  V   (N x KP)   -> ELBO term sum_ng y eta and d psi
//     YtU = Y^T U  (G x KP)   -> d W, d beta
// with U = [psi | X], V = [W | beta].  Both come out of ONE streaming read of Y.  Y is the only
// HBM-sized operand of the whole step, so this kernel is the HBM-roofline kernel.
//
// Layout: Y row-major [N][ldY] (cell-major, genes contiguous), element type float / uint16 / uint8.
// Tiling: CTA = 256 threads owns RB rows x 2048 columns; a thread owns 8 consecutive columns
// (one or two 16-byte loads per row) and keeps their column partials in registers for the whole tile;
// row partials are reduced with a 9-shuffle butterfly per 8 rows.  Partials are written per tile and
// summed in a fixed order by the consumers (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace ca {

constexpr int kYCB = 2048;   // columns per CTA tile
constexpr int kYRI = 8;      // rows reduced together by one butterfly

template <typename T> struct YLoad;
template <> struct YLoad<float> {
  static __device__ __forceinline__ void ld8(const float* p, float (&o)[8]) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(p));
    float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
};
template <> struct YLoad<uint16_t> {
  static __device__ __forceinline__ void ld8(const uint16_t* p, float (&o)[8]) {
    uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
    o[0] = (float)(a.x & 0xffffu); o[1] = (float)(a.x >> 16);
    o[2] = (float)(a.y & 0xffffu); o[3] = (float)(a.y >> 16);
    o[4] = (float)(a.z & 0xffffu); o[5] = (float)(a.z >> 16);
    o[6] = (float)(a.w & 0xffffu); o[7] = (float)(a.w >> 16);
  }
};
template <> struct YLoad<uint8_t> {
  static __device__ __forceinline__ void ld8(const uint8_t* p, float (&o)[8]) {
    uint2 a = __ldcs(reinterpret_cast<const uint2*>(p));
    o[0] = (float)(a.x & 0xffu); o[1] = (float)((a.x >> 8) & 0xffu);
    o[2] = (float)((a.x >> 16) & 0xffu); o[3] = (float)(a.x >> 24);
    o[4] = (float)(a.y & 0xffu); o[5] = (float)((a.y >> 8) & 0xffu);
    o[6] = (float)((a.y >> 16) & 0xffu); o[7] = (float)(a.y >> 24);
  }
};

// Reduce 8 per-lane values across the warp with 9 shuffles.  On return v[0] of lane l holds the
// warp total of value index ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1).
__device__ __forceinline__ float butterfly8(float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = b4 ? v[i] : v[i + 4];
    float keep = b4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(CA_FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = b3 ? v[i] : v[i + 2];
    float keep = b3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(CA_FULL, send, 8);
  }
  {
    float send = b2 ? v[0] : v[1];
    float keep = b2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(CA_FULL, send, 4);
  }
  v[0] += __shfl_xor_sync(CA_FULL, v[0], 2);
  v[0] += __shfl_xor_sync(CA_FULL, v[0], 1);
  return v[0];
}

// KP == 1 (the reference's default model: K = 1 latent dimension, no covariates)
// LIGHT = true halves the rows in flight per thread (4 instead of 8) so that the kernel needs <= 80 registers:
// two of its CTAs then fit on an SM NEXT TO one tcgen05 contraction CTA (27.6k registers, 211 KB smem), which is
// what lets the Y stream (HBM-bound) overlap the forward contraction (tensor-bound) on a second stream.
template <typename T, bool LIGHT>
__device__ __forceinline__ void ypass_tile(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB,
                                           const float* __restrict__ U, const float* __restrict__ Vm,
                                           float* __restrict__ rowpart, float* __restrict__ colpart, int cb, int64_t rb,
                                           float (*red)[8][kYRI]) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t col0 = (int64_t)cb * kYCB + tid * 8;
  const bool colok = col0 < ldY;
  float vr[8], cacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    vr[j] = (col0 + j < G) ? Vm[col0 + j] : 0.f;
    cacc[j] = 0.f;
  }
  const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  int buf = 0;
  constexpr int kInFlight = LIGHT ? 4 : 8;
  for (int64_t r0 = rbeg; r0 < rend; r0 += kYRI) {
    float rp[kYRI];
#pragma unroll
    for (int h = 0; h < kYRI / kInFlight; ++h) {
      float y[kInFlight][8];
      float u[kInFlight];
#pragma unroll
      for (int i = 0; i < kInFlight; ++i) {
        const int64_t r = r0 + h * kInFlight + i;
        if (colok && r < rend) {
          YLoad<T>::ld8(Y + r * ldY + col0, y[i]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[i][j] = 0.f;
        }
        u[i] = (r < rend) ? __ldg(U + r) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kInFlight; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a = fmaf(y[i][j], vr[j], a);
          cacc[j] = fmaf(y[i][j], u[i], cacc[j]);
        }
        rp[h * kInFlight + i] = a;
      }
    }
    float tot = butterfly8(rp, lane);
    if ((lane & 3) == 0) red[buf][wid][ridx] = tot;
    __syncthreads();
    if (tid < kYRI && r0 + tid < rend) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[buf][w][tid];
      rowpart[(int64_t)cb * N + r0 + tid] = a;
    }
    buf ^= 1;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (col0 + j < G) colpart[rb * G + col0 + j] = cacc[j];
  __syncthreads();   // `red` is reused by the next tile of a persistent CTA
}

// one CTA per tile (used when the Y pass runs alone on the GPU)
template <typename T>
__global__ void __launch_bounds__(256, 2)
k_ypass_k1(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, const float* __restrict__ U,
           const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  __shared__ float red[2][8][kYRI];
  ypass_tile<T, false>(Y, ldY, N, G, RB, U, Vm, rowpart, colpart, blockIdx.x, blockIdx.y, red);
}

// persistent, register-light variant for overlap with the contraction kernels: exactly two CTAs per SM
// (grid = 2 x #SM), 72 registers per thread, tiles taken round-robin in a fixed assignment (deterministic)
template <typename T>
__global__ void __maxnreg__(72)
k_ypass_k1_persistent(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB,
                      const float* __restrict__ U, const float* __restrict__ Vm, float* __restrict__ rowpart,
                      float* __restrict__ colpart) {
  __shared__ float red[2][8][kYRI];
  const int64_t ntiles = (int64_t)nCB * nRB;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    ypass_tile<T, true>(Y, ldY, N, G, RB, U, Vm, rowpart, colpart, (int)(t % nCB), t / nCB, red);
}

// generic K + P (slow path, reads Y twice): rows then columns
template <typename T>
__global__ void k_ypass_rows_generic(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int KP,
                                     const float* __restrict__ Vm, float* __restrict__ rowpart) {
  int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[kMaxKP];
#pragma unroll
  for (int kp = 0; kp < kMaxKP; ++kp) acc[kp] = 0.f;
  for (int g = lane; g < G; g += 32) {
    float y = (float)Y[n * ldY + g];
    if (y != 0.f) {
#pragma unroll
      for (int kp = 0; kp < kMaxKP; ++kp)
        if (kp < KP) acc[kp] = fmaf(y, Vm[(int64_t)g * KP + kp], acc[kp]);
    }
  }
#pragma unroll
  for (int kp = 0; kp < kMaxKP; ++kp) {
    if (kp < KP) {
      float t = warp_sum(acc[kp]);
      if (lane == 0) rowpart[n * KP + kp] = t;
    }
  }
}
template <typename T>
__global__ void k_ypass_cols_generic(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int KP, int RB,
                                     const float* __restrict__ U, float* __restrict__ colpart) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int64_t rb = blockIdx.y;
  if (g >= G) return;
  int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
  float acc[kMaxKP];
#pragma unroll
  for (int kp = 0; kp < kMaxKP; ++kp) acc[kp] = 0.f;
  for (int64_t r = rbeg; r < rend; ++r) {
    float y = (float)Y[r * ldY + g];
    if (y != 0.f) {
#pragma unroll
      for (int kp = 0; kp < kMaxKP; ++kp)
        if (kp < KP) acc[kp] = fmaf(y, __ldg(U + r * KP + kp), acc[kp]);
    }
  }
  for (int kp = 0; kp < KP; ++kp) colpart[(rb * G + g) * KP + kp] = acc[kp];
}

}  // namespace ca
