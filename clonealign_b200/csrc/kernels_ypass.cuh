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

// Per storage type: Raw = the 8 consecutive counts a thread owns in one row, exactly as loaded (kept packed in
// registers until use), kRows = rows in flight per thread so that every type keeps 128-256 bytes per thread in flight.
// Narrow integer counts are widened without the (slow, XU-pipe) I2F conversion: a byte permute drops the value
// into the mantissa of 2^23 (0x4B000000) and one FADD removes the 2^23: exact for values < 2^23.
__device__ __forceinline__ float magic_to_float(uint32_t x, uint32_t sel) {
  return __uint_as_float(__byte_perm(x, 0x4B000000u, sel)) - 8388608.0f;
}
template <typename T> struct YLoad;
template <> struct YLoad<float> {
  struct Raw { float4 a, b; };
  static constexpr int kRows = 8;
  static __device__ __forceinline__ Raw ld(const float* p) {
    Raw r;
    r.a = __ldcs(reinterpret_cast<const float4*>(p));
    r.b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ Raw zero() { Raw r; r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; return r; }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&o)[8]) {
    o[0] = r.a.x; o[1] = r.a.y; o[2] = r.a.z; o[3] = r.a.w; o[4] = r.b.x; o[5] = r.b.y; o[6] = r.b.z; o[7] = r.b.w;
  }
};
template <> struct YLoad<uint16_t> {
  typedef uint4 Raw;
  static constexpr int kRows = 16;
  static __device__ __forceinline__ Raw ld(const uint16_t* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& a, float (&o)[8]) {
    o[0] = magic_to_float(a.x, 0x7510u); o[1] = magic_to_float(a.x, 0x7532u);
    o[2] = magic_to_float(a.y, 0x7510u); o[3] = magic_to_float(a.y, 0x7532u);
    o[4] = magic_to_float(a.z, 0x7510u); o[5] = magic_to_float(a.z, 0x7532u);
    o[6] = magic_to_float(a.w, 0x7510u); o[7] = magic_to_float(a.w, 0x7532u);
  }
};
template <> struct YLoad<uint8_t> {
  typedef uint2 Raw;
  static constexpr int kRows = 16;
  static __device__ __forceinline__ Raw ld(const uint8_t* p) { return __ldcs(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& a, float (&o)[8]) {
    o[0] = magic_to_float(a.x, 0x7540u); o[1] = magic_to_float(a.x, 0x7541u);
    o[2] = magic_to_float(a.x, 0x7542u); o[3] = magic_to_float(a.x, 0x7543u);
    o[4] = magic_to_float(a.y, 0x7540u); o[5] = magic_to_float(a.y, 0x7541u);
    o[6] = magic_to_float(a.y, 0x7542u); o[7] = magic_to_float(a.y, 0x7543u);
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
// LIGHT = true halves the rows in flight per thread so that the kernel fits in 72 registers: two of its CTAs then
// fit on an SM NEXT TO one tcgen05 contraction CTA (27.6k registers, 211 KB smem) when the Y stream is overlapped
// with the forward contraction on a second stream.
constexpr int kYMaxRows = 16;
template <typename T, bool LIGHT>
__device__ __forceinline__ void ypass_tile(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB,
                                           const float* __restrict__ U, const float* __restrict__ Vm,
                                           float* __restrict__ rowpart, float* __restrict__ colpart, int cb, int64_t rb,
                                           float (*red)[8][kYMaxRows]) {
  using L = YLoad<T>;
  constexpr int kRows = LIGHT ? L::kRows / 2 : L::kRows;   // rows per iteration (multiple of 4)
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
  for (int64_t r0 = rbeg; r0 < rend; r0 += kRows) {
    typename L::Raw raw[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) raw[i] = (colok && r0 + i < rend) ? L::ld(Y + (r0 + i) * ldY + col0) : L::zero();
    float u[kRows];   // U is allocated with 64 elements of slack, r0 is a multiple of 4: vector loads stay in bounds
#pragma unroll
    for (int i = 0; i < kRows; i += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(U + r0 + i));
      u[i] = t4.x; u[i + 1] = t4.y; u[i + 2] = t4.z; u[i + 3] = t4.w;
    }
    float rp[(kRows + 7) / 8 * 8];
#pragma unroll
    for (int i = 0; i < (kRows + 7) / 8 * 8; ++i) rp[i] = 0.f;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      float y[8];
      L::unpack(raw[i], y);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc = fmaf(y[j], vr[j], acc);
        cacc[j] = fmaf(y[j], u[i], cacc[j]);
      }
      rp[i] = acc;
    }
#pragma unroll
    for (int h = 0; h < (kRows + 7) / 8; ++h) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = rp[h * 8 + i];
      float tot = butterfly8(v8, lane);
      if ((lane & 3) == 0) red[buf][wid][h * 8 + ridx] = tot;
    }
    __syncthreads();
    if (tid < kRows && r0 + tid < rend) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += red[buf][w][tid];
      rowpart[(int64_t)cb * N + r0 + tid] = acc;
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
  __shared__ float red[2][8][kYMaxRows];
  ypass_tile<T, false>(Y, ldY, N, G, RB, U, Vm, rowpart, colpart, blockIdx.x, blockIdx.y, red);
}

// persistent, register-light variant for overlap with the contraction kernels: exactly two CTAs per SM
// (grid = 2 x #SM), 72 registers per thread, tiles taken round-robin in a fixed assignment (deterministic)
template <typename T>
__global__ void __maxnreg__(72)
k_ypass_k1_persistent(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB,
                      const float* __restrict__ U, const float* __restrict__ Vm, float* __restrict__ rowpart,
                      float* __restrict__ colpart) {
  __shared__ float red[2][8][kYMaxRows];
  const int64_t ntiles = (int64_t)nCB * nRB;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    ypass_tile<T, true>(Y, ldY, N, G, RB, U, Vm, rowpart, colpart, (int)(t % nCB), t / nCB, red);
}

// ---------------------------------------------------------------------------------------------------------------
// v2 of the KP == 1 tile: identical tiling / partial layout, arithmetic on PACKED fp32 pairs.
// Why (profiles/r01_notes.md, ncu of round 1): with u8 counts the v1 kernel is bound by the FMA pipe, not by HBM
// (per 8 counts: 8 PRMT on the ALU pipe, 8 FADD + 16 FFMA on the FMA pipe, one warp instruction per 2 cycles per
// sub-partition => 0.32 ms of FMA-pipe time at c3 against a 0.31 ms HBM floor; measured 0.44 ms).  sm_100 has
// two-wide fp32 instructions (add/fma.rn.f32x2 = __fadd2_rn / __ffma2_rn): the magic-number FADD and both FMAs go
// through them, halving the FMA-pipe work (8 PRMT + 4 FADD2 + 8 FFMA2 per 8 counts).  Full tiles also skip the
// per-row bounds predicates.  Sums are re-associated (even / odd columns of a thread are accumulated separately and
// added at the end), so results differ from v1 in the last bits; they stay run-to-run deterministic.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct YLoad2;
template <> struct YLoad2<float> {
  static __device__ __forceinline__ void unpack(const YLoad<float>::Raw& r, float2 (&o)[4]) {
    o[0] = make_float2(r.a.x, r.a.y); o[1] = make_float2(r.a.z, r.a.w);
    o[2] = make_float2(r.b.x, r.b.y); o[3] = make_float2(r.b.z, r.b.w);
  }
};
__device__ __forceinline__ float2 magic_pair(uint32_t x, uint32_t sel0, uint32_t sel1) {
  const float2 m = make_float2(__uint_as_float(__byte_perm(x, 0x4B000000u, sel0)), __uint_as_float(__byte_perm(x, 0x4B000000u, sel1)));
  return __fadd2_rn(m, make_float2(-8388608.0f, -8388608.0f));
}
template <> struct YLoad2<uint16_t> {
  static __device__ __forceinline__ void unpack(const uint4& a, float2 (&o)[4]) {
    o[0] = magic_pair(a.x, 0x7510u, 0x7532u); o[1] = magic_pair(a.y, 0x7510u, 0x7532u);
    o[2] = magic_pair(a.z, 0x7510u, 0x7532u); o[3] = magic_pair(a.w, 0x7510u, 0x7532u);
  }
};
template <> struct YLoad2<uint8_t> {
  static __device__ __forceinline__ void unpack(const uint2& a, float2 (&o)[4]) {
    o[0] = magic_pair(a.x, 0x7540u, 0x7541u); o[1] = magic_pair(a.x, 0x7542u, 0x7543u);
    o[2] = magic_pair(a.y, 0x7540u, 0x7541u); o[3] = magic_pair(a.y, 0x7542u, 0x7543u);
  }
};

template <typename T>
__global__ void __launch_bounds__(256, 2)
k_ypass_k1_v2(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  using L = YLoad<T>;
  constexpr int kRows = L::kRows;
  __shared__ float red[2][8][kYMaxRows];
  const int cb = blockIdx.x;
  const int64_t rb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t col0 = (int64_t)cb * kYCB + tid * 8;
  const bool colok = col0 < ldY;
  float2 vr[4], cacc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    vr[j] = make_float2((col0 + 2 * j < G) ? Vm[col0 + 2 * j] : 0.f, (col0 + 2 * j + 1 < G) ? Vm[col0 + 2 * j + 1] : 0.f);
    cacc[j] = make_float2(0.f, 0.f);
  }
  const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const T* yp = Y + rbeg * ldY + col0;
  int buf = 0;
  for (int64_t r0 = rbeg; r0 < rend; r0 += kRows, yp += (int64_t)kRows * ldY) {
    typename L::Raw raw[kRows];
    if (colok && r0 + kRows <= rend) {       // full tile: no per-row predicates
#pragma unroll
      for (int i = 0; i < kRows; ++i) raw[i] = L::ld(yp + (int64_t)i * ldY);
    } else {
#pragma unroll
      for (int i = 0; i < kRows; ++i) raw[i] = (colok && r0 + i < rend) ? L::ld(yp + (int64_t)i * ldY) : L::zero();
    }
    float u[kRows];   // U is allocated with 64 elements of slack, r0 is a multiple of 4: vector loads stay in bounds
#pragma unroll
    for (int i = 0; i < kRows; i += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(U + r0 + i));
      u[i] = t4.x; u[i + 1] = t4.y; u[i + 2] = t4.z; u[i + 3] = t4.w;
    }
    float rp[(kRows + 7) / 8 * 8];
#pragma unroll
    for (int i = 0; i < (kRows + 7) / 8 * 8; ++i) rp[i] = 0.f;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      float2 y[4];
      YLoad2<T>::unpack(raw[i], y);
      const float2 u2 = make_float2(u[i], u[i]);
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = __ffma2_rn(y[j], vr[j], acc);
        cacc[j] = __ffma2_rn(y[j], u2, cacc[j]);
      }
      rp[i] = acc.x + acc.y;
    }
#pragma unroll
    for (int h = 0; h < (kRows + 7) / 8; ++h) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = rp[h * 8 + i];
      float tot = butterfly8(v8, lane);
      if ((lane & 3) == 0) red[buf][wid][h * 8 + ridx] = tot;
    }
    __syncthreads();
    if (tid < kRows && r0 + tid < rend) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += red[buf][w][tid];
      rowpart[(int64_t)cb * N + r0 + tid] = acc;
    }
    buf ^= 1;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (col0 + 2 * j < G) colpart[rb * G + col0 + 2 * j] = cacc[j].x;
    if (col0 + 2 * j + 1 < G) colpart[rb * G + col0 + 2 * j + 1] = cacc[j].y;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v3 of the KP == 1 tile (variant YPASS3), from the SASS of v2 (538 warp instructions per 4096 counts for u8: 128 PRMT,
// 128 FFMA2, 64 FADD2, 88 for the row butterfly, ~130 address / predicate / loop; issue-limited at ~0.23 ms for the
// 2e9 counts of config 3, uncomfortably close to the 0.31 ms HBM floor):
//   * integer counts are NOT converted: the byte (or half-word) is moved into the low bits of an otherwise zero word,
//     which read as fp32 is the DENORMAL count * 2^-149, exactly.  NVIDIA FMA units take denormal operands at full rate
//     and an FMA rounds once, after the exact product, so with the other operand pre-scaled by 2^100 (psi_n, w_g: one
//     multiply per row / column, exact) every product and partial sum is the v2 value times 2^-49, in the normal range,
//     with the same rounding (fp32 rounding is scale-invariant there); the partial sums are multiplied by 2^49 (exact)
//     when they are stored.  Bit-identical to the widened arithmetic unless |w| or |psi| >= 2^27 (overflow) or a product
//     is below 2^-77 (underflow); the magic-number FADD2 per pair of counts disappears;
//   * u8: a thread owns 16 consecutive columns (one 16-byte load per row, 8 rows in flight) instead of 8, so the row
//     butterfly, the loads of psi and the loop overhead are spread over twice the counts (CTA tile = 4096 columns).
// Same partial-sum layouts (rowpart [nCB][N], colpart [nRB][G]) with nCB = ceil(ldY / (256 * kCols)); summation order
// differs from v1 / v2 in the grouping of columns only.  ~340 warp instructions per 4096 u8 counts.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct Y3;
template <> struct Y3<float> {
  static constexpr int kCols = 8, kRows = 8;
  static constexpr float kPre = 1.f, kPost = 1.f;
  typedef YLoad<float>::Raw Raw;
  static __device__ __forceinline__ Raw ld(const float* p) { return YLoad<float>::ld(p); }
  static __device__ __forceinline__ Raw zero() { return YLoad<float>::zero(); }
  static __device__ __forceinline__ void unpack(const Raw& r, float2 (&o)[4]) { YLoad2<float>::unpack(r, o); }
};
template <> struct Y3<uint16_t> {
  static constexpr int kCols = 8, kRows = 16;
  static constexpr float kPre = 1.2676506002282294e30f /* 2^100 */, kPost = 562949953421312.f /* 2^49 */;
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw ld(const uint16_t* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ float2 pair(uint32_t x) { return make_float2(__uint_as_float(x & 0xffffu), __uint_as_float(x >> 16)); }
  static __device__ __forceinline__ void unpack(const Raw& a, float2 (&o)[4]) {
    o[0] = pair(a.x); o[1] = pair(a.y); o[2] = pair(a.z); o[3] = pair(a.w);
  }
};
template <> struct Y3<uint8_t> {
  static constexpr int kCols = 16, kRows = 8;
  static constexpr float kPre = 1.2676506002282294e30f /* 2^100 */, kPost = 562949953421312.f /* 2^49 */;
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw ld(const uint8_t* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ float2 lo(uint32_t x) {
    return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x4440u)), __uint_as_float(__byte_perm(x, 0u, 0x4441u)));
  }
  static __device__ __forceinline__ float2 hi(uint32_t x) {
    return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x4442u)), __uint_as_float(x >> 24));
  }
  static __device__ __forceinline__ void unpack(const Raw& a, float2 (&o)[8]) {
    o[0] = lo(a.x); o[1] = hi(a.x); o[2] = lo(a.y); o[3] = hi(a.y);
    o[4] = lo(a.z); o[5] = hi(a.z); o[6] = lo(a.w); o[7] = hi(a.w);
  }
};
template <typename T> constexpr int ypass3_tile_cols() { return 256 * Y3<T>::kCols; }

template <typename T>
__global__ void __launch_bounds__(256, 2)
k_ypass_k1_v3(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  using L = Y3<T>;
  constexpr int kRows = L::kRows, kCols = L::kCols, kPairs = kCols / 2;
  __shared__ float red[2][8][kYMaxRows];
  const int cb = blockIdx.x;
  const int64_t rb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t col0 = (int64_t)cb * (256 * kCols) + tid * kCols;
  const bool colok = col0 < ldY;
  float2 vr[kPairs], cacc[kPairs];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    vr[j] = make_float2((col0 + 2 * j < G) ? Vm[col0 + 2 * j] * L::kPre : 0.f, (col0 + 2 * j + 1 < G) ? Vm[col0 + 2 * j + 1] * L::kPre : 0.f);
    cacc[j] = make_float2(0.f, 0.f);
  }
  const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const T* yp = Y + rbeg * ldY + col0;
  int buf = 0;
  for (int64_t r0 = rbeg; r0 < rend; r0 += kRows, yp += (int64_t)kRows * ldY) {
    typename L::Raw raw[kRows];
    if (colok && r0 + kRows <= rend) {       // full tile: no per-row predicates
#pragma unroll
      for (int i = 0; i < kRows; ++i) raw[i] = L::ld(yp + (int64_t)i * ldY);
    } else {
#pragma unroll
      for (int i = 0; i < kRows; ++i) raw[i] = (colok && r0 + i < rend) ? L::ld(yp + (int64_t)i * ldY) : L::zero();
    }
    float u[kRows];   // U is allocated with 64 elements of slack, r0 is a multiple of 4: vector loads stay in bounds
#pragma unroll
    for (int i = 0; i < kRows; i += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(U + r0 + i));
      u[i] = t4.x * L::kPre; u[i + 1] = t4.y * L::kPre; u[i + 2] = t4.z * L::kPre; u[i + 3] = t4.w * L::kPre;
    }
    float rp[(kRows + 7) / 8 * 8];
#pragma unroll
    for (int i = 0; i < (kRows + 7) / 8 * 8; ++i) rp[i] = 0.f;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      float2 y[kPairs];
      L::unpack(raw[i], y);
      const float2 u2 = make_float2(u[i], u[i]);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);   // two chains: the 8 pairs of a u8 row
#pragma unroll
      for (int j = 0; j < kPairs; j += 2) {
        acc0 = __ffma2_rn(y[j], vr[j], acc0);
        acc1 = __ffma2_rn(y[j + 1], vr[j + 1], acc1);
        cacc[j] = __ffma2_rn(y[j], u2, cacc[j]);
        cacc[j + 1] = __ffma2_rn(y[j + 1], u2, cacc[j + 1]);
      }
      const float2 a2 = __fadd2_rn(acc0, acc1);
      rp[i] = a2.x + a2.y;
    }
#pragma unroll
    for (int h = 0; h < (kRows + 7) / 8; ++h) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = rp[h * 8 + i];
      float tot = butterfly8(v8, lane);
      if ((lane & 3) == 0) red[buf][wid][h * 8 + ridx] = tot;
    }
    __syncthreads();
    if (tid < kRows && r0 + tid < rend) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += red[buf][w][tid];
      rowpart[(int64_t)cb * N + r0 + tid] = acc * L::kPost;
    }
    buf ^= 1;
  }
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    if (col0 + 2 * j < G) colpart[rb * G + col0 + 2 * j] = cacc[j].x * L::kPost;
    if (col0 + 2 * j + 1 < G) colpart[rb * G + col0 + 2 * j + 1] = cacc[j].y * L::kPost;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v4 of the KP == 1 tile (variant YPASS4), from the round-2 ncu captures (profiles/r02_notes.md).  v3: 16 warps per SM,
// issue slots 41 % busy, top stall long_scoreboard -- every warp loads 8 rows into REGISTERS, waits about a microsecond,
// then spends about as long on its ~430 instructions: nothing is in flight while it computes (128 registers per thread
// leave no room to double-buffer), and 2 CTAs take the whole register file, so nothing can share the SM with the stream.
// Two staging schemes were measured and dropped: per-thread cp.async (LDGSTS) copies removed the stall but cost +57 %
// instructions (address arithmetic, a commit / wait and three compiler-inserted dummy LDS per 16-byte copy): issue-bound
// at the same speed; 512-byte bulk copies per (warp, row) ran at 3.4 TB/s -- the TMA engine serves about one request
// per 46 cycles per SM, so small requests starve it.
// Here a row of the CTA's column tile (256 threads x 16 bytes = 4 KB contiguous) is ONE bulk copy (cp.async.bulk, the TMA
// engine) into a shared-memory ring of 16 rows (4 stages of 4 rows, one mbarrier per stage counting bytes), issued by one
// thread; every thread reads its 16-byte piece of a row with one LDS.128.  The block barrier the row sums need every 8
// rows doubles as the "stage is free" signal: right after it one thread re-arms and refills the two stages just consumed,
// so the next 8 rows are in flight WHILE the CTA computes.  Bytes in flight cost neither registers nor issue slots, the
// kernel needs half the registers, and it is launched as a PERSISTENT grid of 2 CTAs per SM (fixed tile assignment:
// deterministic) that leaves half of the register file and ~100 KB of shared memory of every SM free: started first in
// the step on the second stream (variant COSCHED), the HBM-bound stream runs NEXT TO the issue-bound kernels of the step.
// Arithmetic, tiling (256 x kCols columns) and partial-sum layouts are those of v3: bit-identical results.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kY4StageRows = 4, kY4Stages = 4;
template <typename T> constexpr size_t ypass4_smem_bytes() {
  return (size_t)kY4Stages * kY4StageRows * 256 * sizeof(typename Y3<T>::Raw) + 8 * kY4Stages + 16;
}

// (bar_init / bar_arm / bar_wait / bulk_copy / fence_*: platform.cuh)

// MINB = CTAs per SM the register allocation is sized for: 4 -> 64 registers, 3 -> 80
template <typename T, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_ypass_k1_v4(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  using L = Y3<T>;
  using Raw = typename L::Raw;
  constexpr int kSR = kY4StageRows, kCols = L::kCols, kPairs = kCols / 2;
  constexpr uint32_t kRowBytes = 256 * sizeof(Raw);              // one row of the CTA's column tile
  CA_DYNAMIC_SMEM(unsigned char, ring_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  Raw* ring = reinterpret_cast<Raw*>(ring_raw);                  // [stage][row][thread]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + (size_t)kY4Stages * kSR * kRowBytes);
  __shared__ float red[2][8][8];
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  // the ring starts zeroed: threads past the last stored column never receive data and must not read NaN patterns
  for (int i = tid; i < kY4Stages * kSR * 256; i += 256) ring[i] = L::zero();
  if (tid == 0)
    for (int st = 0; st < kY4Stages; ++st) bar_init(bars + st, 1);
  fence_bar_init();
  fence_proxy_async();                                              // the zero fill (generic stores) before any bulk copy
  __syncthreads();
  uint32_t ph = 0;                                               // bit st: parity the next wait on stage st has to see
  const int64_t ntiles = (int64_t)nCB * nRB;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int cb = (int)(tile % nCB);
    const int64_t rb = tile / nCB;
    const int64_t tcol0 = (int64_t)cb * (256 * kCols);           // first column of the CTA's tile
    const int64_t col0 = tcol0 + tid * kCols;
    // bytes of a row that exist for this tile (the stored row ends at ldY; ldY * sizeof(T) and the tile start are multiples of 16)
    const int64_t avail = (ldY - tcol0) * (int64_t)sizeof(T);
    const uint32_t tbytes = avail < (int64_t)kRowBytes ? (uint32_t)avail : kRowBytes;
    float2 vr[kPairs], cacc[kPairs];
#pragma unroll
    for (int j = 0; j < kPairs; ++j) {
      vr[j] = make_float2((col0 + 2 * j < G) ? Vm[col0 + 2 * j] * L::kPre : 0.f, (col0 + 2 * j + 1 < G) ? Vm[col0 + 2 * j + 1] * L::kPre : 0.f);
      cacc[j] = make_float2(0.f, 0.f);
    }
    const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
    const int nrows = (int)(rend - rbeg);
    const int ngroups = (nrows + 7) / 8;
    const T* ybase = Y + rbeg * ldY + tcol0;
    // group `g` of the tile (rows 8 g ..) lives in stages S0, S0 + 1 with S0 = 2 (g & 1): even groups in stages 0 / 1, odd
    // groups in 2 / 3 (the loop below is unrolled by two, so stage addresses are compile-time offsets); thread 0 issues.
    auto issue = [&](auto s0c, int g) {
      constexpr int S0 = decltype(s0c)::value;
      if (g >= ngroups) return;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r0 = g * 8 + hh * kSR;
        const int nv = nrows - r0 < kSR ? nrows - r0 : kSR;
        if (nv <= 0) break;
        bar_arm(bars + S0 + hh, (uint32_t)nv * tbytes);
        for (int i = 0; i < nv; ++i)
          bulk_copy(ring + ((size_t)(S0 + hh) * kSR + i) * 256, ybase + (int64_t)(r0 + i) * ldY, tbytes, bars + S0 + hh);
      }
    };
    if (tid == 0) { issue(std::integral_constant<int, 0>{}, 0); issue(std::integral_constant<int, 2>{}, 1); }
    CA_SYNC_AFTER_SYNCHRONOUS_COPY();                            // nothing on the device (see platform.cuh)
    auto group = [&](auto s0c, int g) {
      constexpr int S0 = decltype(s0c)::value;
      const int64_t r0 = rbeg + (int64_t)g * 8;
      float u[8];   // U is allocated with 64 elements of slack, r0 is a multiple of 4: vector loads stay in bounds
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(U + r0 + i));
        u[i] = t4.x * L::kPre; u[i + 1] = t4.y * L::kPre; u[i + 2] = t4.z * L::kPre; u[i + 3] = t4.w * L::kPre;
      }
      float rp[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nv = nrows - (g * 8 + hh * kSR);               // valid rows of this stage
        if (nv > 0) {
          bar_wait(bars + S0 + hh, (ph >> (S0 + hh)) & 1u);
          ph ^= 1u << (S0 + hh);
          Raw* src = ring + (size_t)(S0 + hh) * kSR * 256 + tid;
          if (nv < kSR) {
            // last stage of the tile: its trailing rows were not copied and hold stale ring contents.  Every thread
            // zeroes its own pieces (it reads them back itself below); the proxy fence orders these generic stores
            // before the bulk copy that refills the stage later.
            for (int i = nv; i < kSR; ++i) src[(size_t)i * 256] = L::zero();
            fence_proxy_async();
          }
#pragma unroll
          for (int i = 0; i < kSR; ++i) {
            const Raw raw = src[(size_t)i * 256];
            float2 y[kPairs];
            L::unpack(raw, y);
            const float2 u2 = make_float2(u[hh * kSR + i], u[hh * kSR + i]);
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kPairs; j += 2) {
              acc0 = __ffma2_rn(y[j], vr[j], acc0);
              acc1 = __ffma2_rn(y[j + 1], vr[j + 1], acc1);
              cacc[j] = __ffma2_rn(y[j], u2, cacc[j]);
              cacc[j + 1] = __ffma2_rn(y[j + 1], u2, cacc[j + 1]);
            }
            const float2 a2 = __fadd2_rn(acc0, acc1);
            rp[hh * kSR + i] = a2.x + a2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < kSR; ++i) rp[hh * kSR + i] = 0.f;
        }
      }
      const float tot = butterfly8(rp, lane);
      const int buf = S0 >> 1;                                   // even / odd groups alternate between the two exchange buffers
      if ((lane & 3) == 0) red[buf][wid][ridx] = tot;
      __syncthreads();
      // every thread has issued the FMAs that consume its pieces of this group (in-order issue: their LDS have returned)
      // and passed the barrier: the two stages are free and are refilled with the group after next
      if (tid == 0) issue(s0c, g + 2);
      if (tid < 8 && r0 + tid < rend) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += red[buf][w][tid];
        rowpart[(int64_t)cb * N + r0 + tid] = acc * L::kPost;
      }
    };
    for (int g = 0; g < ngroups; g += 2) {
      group(std::integral_constant<int, 0>{}, g);
      if (g + 1 < ngroups) group(std::integral_constant<int, 2>{}, g + 1);
    }
#pragma unroll
    for (int j = 0; j < kPairs; ++j) {
      if (col0 + 2 * j < G) colpart[rb * G + col0 + 2 * j] = cacc[j].x * L::kPost;
      if (col0 + 2 * j + 1 < G) colpart[rb * G + col0 + 2 * j + 1] = cacc[j].y * L::kPost;
    }
    __syncthreads();   // `red` is reused by the next tile; every stage this CTA issued has been consumed
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_ypass_k1_v5 (variant YPASS5, u8 storage): the two products of the Y pass on the INTEGER tensor pipe.
// The co-scheduled timeline of round 2 (profiles/r02_notes.md section 3b) shows the step bound by instruction issue, and the
// Y pass owns 74 % of the step's warp instructions: byte-wise arithmetic on the FMA pipe costs one PRMT and half a packed FMA
// per count and product.  Both products are exact integer contractions once the fp32 operand is written in base-128 digits:
//     w_g   = 2^e_w   sum_d D_d(g) 2^(-6 - 7 d),   psi_n = 2^e_psi sum_d P_d(n) 2^(-6 - 7 d),   D_d, P_d in [-64, 64], d = 0 .. 3
// (four digits: absolute error 2^-28 of the tile's power-of-two scale -- finer than an fp32 ulp of the largest operand and
// than the rounding of an fp32 accumulation chain; the digit recursion y -> rint(y), 128 (y - rint(y)) is exact in fp32), so
//     (Y W)_n     = 2^e_w   sum_d 2^(-6 - 7 d) [ sum_g y_ng D_d(g) ]      mma.sync m16n8k32: A = Y tile (u8), B = W digits (s8)
//     (Y^T psi)_g = 2^e_psi sum_d 2^(-6 - 7 d) [ sum_n y_ng P_d(n) ]      A = Y tile TRANSPOSED (ldmatrix.trans + 4 PRMT per
//                                                                        512 counts regroup bytes by gene), B = psi digits
// with s32 accumulation (bounds: 255 * 64 * 256 columns resp. * 4 096 rows < 2^27).  Per 512 counts a warp issues ~10
// instructions instead of ~55, the results do not depend on the order of accumulation at all, and the FMA / ALU pipes stay
// free for the kernels that run next to the pass.  One persistent CTA per SM (8 warps, <= 128 registers, 161 KB): the other
// half of the register file and 66 KB of shared memory are left to the per-cell kernel and the gene-level launches.
//   tile    = kY5Cols (2 048) columns x RB rows (multiple of 32, <= kY5MaxRows); warp w owns columns [256 w, 256 w + 256)
//   ring    = 2 stages x 32 rows x (2 048 + 16) bytes: a row is one bulk copy (TMA engine); the 16-byte skew puts the 8 rows
//             of an ldmatrix tile on distinct banks; one mbarrier per stage, the block barrier behind a stage's row sums frees it
//   stage   = 2 x 8 (ldmatrix + mma) for the row sums of its 32 rows, 16 x (ldmatrix.trans + 4 PRMT + mma) for the column sums
// Rows of a stage beyond the tile have psi digits 0 (stale ring contents times 0), columns beyond G have W digits 0.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kY5Cols = 2048, kY5StageRows = 32, kY5Stages = 2, kY5Pitch = kY5Cols + 16, kY5MaxRows = 2048;
inline size_t ypass5_smem_bytes() {
  return (size_t)kY5Stages * kY5StageRows * kY5Pitch + (size_t)(kY5MaxRows / 32) * 128 + 8 * kY5Stages + 16;
}
// power of two >= |v| (v finite): exponent arithmetic only
__device__ __forceinline__ float y5_pow2_ceil(float amax) {
  if (!(amax > 0.f)) return 1.f;
  int e = ((__float_as_int(amax) >> 23) & 0xff) + 1;            // 2^(e - 127) > amax
  e = e > 254 ? 254 : e;
  return __int_as_float(e << 23);
}
// digit d (0 .. 3) of x in [-1, 1]: y_0 = 64 x, D = rint(y), y <- 128 (y - D)
__device__ __forceinline__ int y5_digit(float x, int d) {
  float y = x * 64.f;
  float D = rintf(y);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < d) { y = (y - D) * 128.f; D = rintf(y); }
  }
  return (int)D;
}
// NW warps per CTA (8: 128 registers each, 16: 64 -- twice the warps to hide the ldmatrix -> mma latency next to other kernels)
template <int NW>
__global__ void __launch_bounds__(NW * 32, 2)
k_ypass_k1_v5(const uint8_t* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  CA_DYNAMIC_SMEM(unsigned char, ring5);
  unsigned char* ring = ring5;                                                                   // [stage][row][kY5Pitch]
  uint2* psd = reinterpret_cast<uint2*>(ring5 + (size_t)kY5Stages * kY5StageRows * kY5Pitch);      // [32-row block][digit][t]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring5 + (size_t)kY5Stages * kY5StageRows * kY5Pitch + (size_t)(kY5MaxRows / 32) * 128);
  constexpr int kY5WarpCols = kY5Cols / NW, kY5KB = kY5WarpCols / 32, kY5GB = kY5WarpCols / 16, kThreads = NW * 32;
  __shared__ double red[2][NW][kY5StageRows];
  __shared__ float sred[NW];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;                           // fragment coordinates (groupID, threadID_in_group)
  for (int i = tid; i < kY5Stages * kY5StageRows * kY5Pitch / 16; i += kThreads) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0)
    for (int st = 0; st < kY5Stages; ++st) bar_init(bars + st, 1);
  fence_bar_init();
  fence_proxy_async();
  __syncthreads();
  uint32_t ph = 0;
  const int64_t ntiles = (int64_t)nCB * nRB;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int cb = (int)(tile % nCB);
    const int64_t rb = tile / nCB;
    const int64_t tcol0 = (int64_t)cb * kY5Cols;
    const int64_t avail = ldY - tcol0;
    const uint32_t tbytes = avail < (int64_t)kY5Cols ? (uint32_t)avail : (uint32_t)kY5Cols;
    const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
    const int nrows = (int)(rend - rbeg);
    const int nstages = (nrows + kY5StageRows - 1) / kY5StageRows;
    const uint8_t* ybase = Y + rbeg * ldY + tcol0;
    // stage group sg -> ring stage sg % kY5Stages; issued by the 32 lanes of warp 0, one row each (a single issuing thread
    // is a bottleneck of its own: ~80 cycles per bulk copy, measured)
    auto issue = [&](int sg) {
      if (sg >= nstages) return;
      const int st = sg % kY5Stages;
      const int r0 = sg * kY5StageRows;
      const int nv = nrows - r0 < kY5StageRows ? nrows - r0 : kY5StageRows;
      if (lane == 0) bar_arm(bars + st, (uint32_t)nv * tbytes);
      __syncwarp();
      if (lane < nv) bulk_copy(ring + ((size_t)st * kY5StageRows + lane) * kY5Pitch, ybase + (int64_t)(r0 + lane) * ldY, tbytes, bars + st);
    };
    if (wid == 0)
      for (int sg = 0; sg < kY5Stages; ++sg) issue(sg);
    CA_SYNC_AFTER_SYNCHRONOUS_COPY();
    // ---- per-tile operands (computed while the first stages travel): scales, W digit fragments, psi digit table ----
    float wm = 0.f, pm = 0.f;
    // (a NaN / infinite operand turns the maximum into +inf: fmaxf alone would drop a NaN)
    const float kInf = __int_as_float(0x7f800000);
    for (int c = tid; c < kY5Cols; c += kThreads) {
      const int64_t col = tcol0 + c;
      if (col < G) { const float v = fabsf(Vm[col]); wm = (v <= 3.0e38f) ? fmaxf(wm, v) : kInf; }
    }
    for (int r = tid; r < nrows; r += kThreads) { const float v = fabsf(U[rbeg + r]); pm = (v <= 3.0e38f) ? fmaxf(pm, v) : kInf; }
    wm = warp_max(wm); pm = warp_max(pm);
    __syncthreads();                                                 // sred / psd of the previous tile are no longer read
    if (lane == 0) sred[wid] = wm;
    __syncthreads();
    wm = sred[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) wm = fmaxf(wm, sred[i]);
    __syncthreads();
    if (lane == 0) sred[wid] = pm;
    __syncthreads();
    pm = sred[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) pm = fmaxf(pm, sred[i]);
    const float sw = y5_pow2_ceil(wm), sp = y5_pow2_ceil(pm);        // NaN operands: scale 1, digits of NaN are garbage -> results
    const float isw = 1.f / sw, isp = 1.f / sp;                      // are flagged below
    const bool bad = !(wm <= 3.0e38f) || !(pm <= 3.0e38f);           // NaN / inf in W or psi: the partials of the tile are NaN
    // W digits of this warp's 256 columns as B fragments: bw[kb][0] = digit g of columns 32 kb + 4 t .. + 3, [1]: + 16
    uint32_t bw[kY5KB][2];
#pragma unroll
    for (int kb = 0; kb < kY5KB; ++kb) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk = 0u;
        if (g < 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t col = tcol0 + wid * kY5WarpCols + kb * 32 + h * 16 + t * 4 + j;
            const int D = col < G ? y5_digit(Vm[col] * isw, g) : 0;
            pk |= ((uint32_t)D & 0xffu) << (8 * j);
          }
        }
        bw[kb][h] = pk;
      }
    }
    // psi digits of the tile's rows, laid out as the B fragments of the column-sum products: block k (32 rows), digit d, lane
    // coordinate t: word 0 = digit d of rows (2 t, 2 t + 1, 8 + 2 t, 9 + 2 t), word 1 = rows (16 + 2 t, 17 + 2 t, 24 + 2 t, 25 + 2 t)
    for (int it = tid; it < nstages * 16; it += kThreads) {
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
    int cacc[kY5GB][4];                                            // column sums: blocks of 16 genes x (digit columns)
#pragma unroll
    for (int gb = 0; gb < kY5GB; ++gb)
#pragma unroll
      for (int i = 0; i < 4; ++i) cacc[gb][i] = 0;
    for (int sg = 0; sg < nstages; ++sg) {
      const int st = sg % kY5Stages;
      bar_wait(bars + st, (ph >> st) & 1u);
      ph ^= 1u << st;
      const unsigned char* sbase = ring + (size_t)st * kY5StageRows * kY5Pitch + wid * kY5WarpCols;
      // ---- row sums: 2 halves of 16 rows x 8 blocks of 32 columns ----
      int racc[2][4];
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
        for (int i = 0; i < 4; ++i) racc[rh][i] = 0;
        const unsigned char* ap = sbase + (size_t)(rh * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kY5Pitch + 16 * (lane >> 4);
#pragma unroll
        for (int kb = 0; kb < kY5KB; ++kb) {
          uint32_t a[4];
          ldmatrix_x4(a, ap + kb * 32);
          mma_u8s8(racc[rh], a, bw[kb][0], bw[kb][1]);
        }
      }
      // ---- column sums: 16 blocks of 16 genes x the 32 rows of the stage ----
      const uint2 pb = (g < 4) ? psd[sg * 16 + g * 4 + t] : make_uint2(0u, 0u);
      const unsigned char* tp = sbase + (size_t)lane * kY5Pitch;
#pragma unroll
      for (int gb = 0; gb < kY5GB; ++gb) {
        uint32_t r[4], a[4];
        ldmatrix_x4_trans(r, tp + gb * 16);
        a[0] = __byte_perm(r[0], r[1], 0x6420u);                   // gene 2 g   : rows (2 t, 2 t + 1, 8 + 2 t, 9 + 2 t)
        a[1] = __byte_perm(r[0], r[1], 0x7531u);                   // gene 2 g + 1
        a[2] = __byte_perm(r[2], r[3], 0x6420u);                   // rows (16 + 2 t, ...)
        a[3] = __byte_perm(r[2], r[3], 0x7531u);
        mma_u8s8(cacc[gb], a, pb.x, pb.y);
      }
      // ---- row sums of the stage: digits -> value, lanes t = 0 (digits 0, 1) + t = 1 (digits 2, 3), then over the warps ----
      const int buf = sg & 1;
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const double s0 = t == 0 ? 0.015625 : (t == 1 ? 9.5367431640625e-07 : 0.0);                 // 2^-6, 2^-20
        const double s1 = t == 0 ? 0.0001220703125 : (t == 1 ? 7.450580596923828e-09 : 0.0);       // 2^-13, 2^-27
        double lo = (double)racc[rh][0] * s0 + (double)racc[rh][1] * s1;                           // row g
        double hi = (double)racc[rh][2] * s0 + (double)racc[rh][3] * s1;                           // row g + 8
        lo += __shfl_xor_sync(CA_FULL, lo, 1);
        hi += __shfl_xor_sync(CA_FULL, hi, 1);
        if (t == 0) { red[buf][wid][rh * 16 + g] = lo; red[buf][wid][rh * 16 + g + 8] = hi; }
      }
      __syncthreads();
      // every warp has consumed the stage (its ldmatrix results are in registers): refill it with the stage after next
      if (wid == 0) issue(sg + kY5Stages);
      if (tid < kY5StageRows && sg * kY5StageRows + tid < nrows) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += red[buf][w][tid];
        rowpart[(int64_t)cb * N + rbeg + sg * kY5StageRows + tid] = bad ? __int_as_float(0x7fc00000) : (float)(acc * (double)sw);
      }
    }
    // ---- column sums of the tile: digits -> value (lanes t = 0, 1), gene 2 g and 2 g + 1 of every block ----
#pragma unroll
    for (int gb = 0; gb < kY5GB; ++gb) {
      const double s0 = t == 0 ? 0.015625 : (t == 1 ? 9.5367431640625e-07 : 0.0);
      const double s1 = t == 0 ? 0.0001220703125 : (t == 1 ? 7.450580596923828e-09 : 0.0);
      double ev = (double)cacc[gb][0] * s0 + (double)cacc[gb][1] * s1;                             // gene 16 gb + 2 g
      double od = (double)cacc[gb][2] * s0 + (double)cacc[gb][3] * s1;                             // gene 16 gb + 2 g + 1
      ev += __shfl_xor_sync(CA_FULL, ev, 1);
      od += __shfl_xor_sync(CA_FULL, od, 1);
      if (t == 0) {
        const int64_t col = tcol0 + wid * kY5WarpCols + gb * 16 + 2 * g;
        if (col < G) colpart[rb * G + col] = bad ? __int_as_float(0x7fc00000) : (float)(ev * (double)sp);
        if (col + 1 < G) colpart[rb * G + col + 1] = bad ? __int_as_float(0x7fc00000) : (float)(od * (double)sp);
      }
    }
    __syncthreads();                                                 // red / psd / sred are reused by the next tile
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_ypass_k1_v6: the arithmetic of k_ypass_k1_v5 without a block barrier per stage (CLONEALIGN_B200_Y5_SPEC=1).
// v5 next to the other kernels of the step is held up as much as the FMA-pipe pass (profiles/r02_notes.md section 3c): its 8 warps
// meet at a block barrier after every 32-row stage, so the slowest warp -- delayed by whatever else the schedulers issue -- sets the
// pace of the CTA.  Here the stages are handed over through mbarriers only:
//   * warps 0 .. 5 (consumers): wait for full[stage] -> row-sum and column-sum products of their 256 columns -> integer digit sums
//     of the 32 rows added into shared memory with red.shared.add.s32 (integer: order-independent) -> arrive on empty[stage];
//   * warp 6 (producer): waits for empty[stage] (6 arrivals), turns the stage's digit sums into the row partials (lane = row),
//     zeroes them, re-arms full[stage] and issues the 32 bulk copies of the stage after next.
// Block barriers remain at tile boundaries only (the per-tile scales and digit tables are built by all threads).  Tiles are 1 536 columns
// wide in a ring of THREE stages (two travel while one is consumed; two stages of 1 792 columns left the pass 24 % above the time of its
// bytes): 6 consumer warps + the producer, at most 2 warps per scheduler (a ninth warp takes a third register slice on one scheduler and
// pushes the co-scheduled 16-warp kernels off the SM: measured, the step serialised).  W digit fragments live in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kY6Consumers = 6, kY6Threads = 32 * (kY6Consumers + 1), kY6Cols = kY6Consumers * 256, kY6Pitch = kY6Cols + 16, kY6Stages = 3;   // 1 536-column tiles
inline size_t ypass6_smem_bytes() {
  return (size_t)kY6Stages * kY5StageRows * kY6Pitch + (size_t)(kY5MaxRows / 32) * 128 + (size_t)kY6Consumers * 8 * 32 * 8 +
         (size_t)kY6Stages * kY5StageRows * 16 + 8 * 2 * kY6Stages + 16;
}
__global__ void __launch_bounds__(kY6Threads, 2)
k_ypass_k1_v6(const uint8_t* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, int nCB, int nRB, const float* __restrict__ U,
              const float* __restrict__ Vm, float* __restrict__ rowpart, float* __restrict__ colpart) {
  CA_DYNAMIC_SMEM(unsigned char, ring6);
  constexpr int kWarpCols = 256, kKB = 8, kGB = 16;
  unsigned char* ring = ring6;                                                                   // [stage][row][kY6Pitch]
  unsigned char* p0 = ring6 + (size_t)kY6Stages * kY5StageRows * kY6Pitch;
  uint2* psd = reinterpret_cast<uint2*>(p0);                                                      // [32-row block][digit][t]
  uint2* bws = reinterpret_cast<uint2*>(p0 + (size_t)(kY5MaxRows / 32) * 128);                    // [warp][kb][lane]: W digit fragments
  int* rsum = reinterpret_cast<int*>(p0 + (size_t)(kY5MaxRows / 32) * 128 + (size_t)kY6Consumers * 8 * 32 * 8);   // [stage][row][digit]
  uint64_t* full = reinterpret_cast<uint64_t*>(rsum + kY6Stages * kY5StageRows * 4);
  uint64_t* empty = full + kY6Stages;
  __shared__ float sred[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool producer = wid == kY6Consumers;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < kY6Stages * kY5StageRows * kY6Pitch / 16; i += kY6Threads) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
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
    const int64_t avail = ldY - tcol0;
    const uint32_t tbytes = avail < (int64_t)kY6Cols ? (uint32_t)avail : (uint32_t)kY6Cols;
    const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
    const int nrows = (int)(rend - rbeg);
    const int nstages = (nrows + kY5StageRows - 1) / kY5StageRows;
    const uint8_t* ybase = Y + rbeg * ldY + tcol0;
    auto issue = [&](int sg) {                                       // producer warp: stage sg of this tile -> slot (j0 + sg) % kY6Stages
      if (sg >= nstages) return;
      const int st = (int)((j0 + (uint32_t)sg) % kY6Stages);
      const int r0 = sg * kY5StageRows;
      const int nv = nrows - r0 < kY5StageRows ? nrows - r0 : kY5StageRows;
      if (lane == 0) bar_arm(full + st, (uint32_t)nv * tbytes);
      __syncwarp();
      if (lane < nv) bulk_copy(ring + ((size_t)st * kY5StageRows + lane) * kY6Pitch, ybase + (int64_t)(r0 + lane) * ldY, tbytes, full + st);
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
        const unsigned char* sbase = ring + (size_t)st * kY5StageRows * kY6Pitch + wid * kWarpCols;
        int racc[2][4];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
          for (int i = 0; i < 4; ++i) racc[rh][i] = 0;
          const unsigned char* ap = sbase + (size_t)(rh * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kY6Pitch + 16 * (lane >> 4);
#pragma unroll
          for (int kb = 0; kb < kKB; ++kb) {
            uint32_t a[4];
            ldmatrix_x4(a, ap + kb * 32);
            const uint2 b = bws[(wid * kKB + kb) * 32 + lane];
            mma_u8s8(racc[rh], a, b.x, b.y);
          }
        }
        const uint2 pb = (g < 4) ? psd[sg * 16 + g * 4 + t] : make_uint2(0u, 0u);
        const unsigned char* tp = sbase + (size_t)lane * kY6Pitch;
#pragma unroll
        for (int gb = 0; gb < kGB; ++gb) {
          uint32_t r[4], a[4];
          ldmatrix_x4_trans(r, tp + gb * 16);
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

// ---------------------------------------------------------------------------------------------------------------
// Batched Y pass for R fits that share one count matrix (restarts of run_clonealign on one device):
// ONE stream over Y produces (Y W_r, Y^T psi_r) for every fit r -- the widening / magic-number work is shared and the
// matrix leaves HBM once instead of R times.  Same tiling and partial layouts as k_ypass_k1_v2 (each fit's own rowpart /
// colpart buffers, summed in the same fixed order by its consumers), 8 rows per iteration to keep R x 8 row partials
// and R x 8 column accumulators in registers.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kYMultiMax = 4;
struct YMultiArgs {
  const float* U[kYMultiMax];
  const float* Vm[kYMultiMax];
  float* rowpart[kYMultiMax];
  float* colpart[kYMultiMax];
};

template <typename T, int R>
__global__ void __launch_bounds__(256, 1)
k_ypass_k1_multi(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RB, YMultiArgs a) {
  using L = YLoad<T>;
  constexpr int kRows = 8;
  __shared__ float red[2][8][R * kRows];
  const int cb = blockIdx.x;
  const int64_t rb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t col0 = (int64_t)cb * kYCB + tid * 8;
  const bool colok = col0 < ldY;
  float2 vr[R][4], cacc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      vr[r][j] = make_float2((col0 + 2 * j < G) ? a.Vm[r][col0 + 2 * j] : 0.f, (col0 + 2 * j + 1 < G) ? a.Vm[r][col0 + 2 * j + 1] : 0.f);
      cacc[r][j] = make_float2(0.f, 0.f);
    }
  const int64_t rbeg = rb * RB, rend = (rbeg + RB < N) ? rbeg + RB : N;
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const T* yp = Y + rbeg * ldY + col0;
  int buf = 0;
  for (int64_t r0 = rbeg; r0 < rend; r0 += kRows, yp += (int64_t)kRows * ldY) {
    typename L::Raw raw[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) raw[i] = (colok && r0 + i < rend) ? L::ld(yp + (int64_t)i * ldY) : L::zero();
    float rp[R][kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      float2 y[4];
      YLoad2<T>::unpack(raw[i], y);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float u = __ldg(a.U[r] + r0 + i);     // U has 64 elements of slack: in bounds for the masked tail rows too
        const float2 u2 = make_float2(u, u);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc = __ffma2_rn(y[j], vr[r][j], acc);
          cacc[r][j] = __ffma2_rn(y[j], u2, cacc[r][j]);
        }
        rp[r][i] = acc.x + acc.y;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = rp[r][i];
      const float tot = butterfly8(v8, lane);
      if ((lane & 3) == 0) red[buf][wid][r * kRows + ridx] = tot;
    }
    __syncthreads();
    if (tid < R * kRows) {
      const int r = tid / kRows, i = tid % kRows;
      if (r0 + i < rend) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += red[buf][w][tid];
        a.rowpart[r][(int64_t)cb * N + r0 + i] = acc;
      }
    }
    buf ^= 1;
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col0 + 2 * j < G) a.colpart[r][rb * G + col0 + 2 * j] = cacc[r][j].x;
      if (col0 + 2 * j + 1 < G) a.colpart[r][rb * G + col0 + 2 * j + 1] = cacc[r][j].y;
    }
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
