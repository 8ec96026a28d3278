// psi initialisation on the device: the leading principal component of the centred, scaled log2(Y + 1)
// -- prcomp(log2(Y_dat + 1), center = TRUE, scale = TRUE)$x[, 1] of R/inference-tflow.R:203-204 -- by power iteration
// on the count matrix that is already resident in HBM, instead of a host-side O(N G^2) prcomp.
// With l = log2(y + 1), column means m_g and standard deviations s_g (n - 1 denominator, as R's scale()):
//     X = (l - m) / s,   t = X v = l (v / s) - m . (v / s),   X^T t = (l^T t - m sum(t)) / s
// so one iteration is a row pass and a column pass over Y with the transform applied on the fly; X is never formed.
// All accumulations are fp64 and in a fixed order.
#pragma once
#include "common.cuh"

namespace ca {

__device__ __forceinline__ float pca_l2(float y) { return log2f(y + 1.0f); }

// column slices: part[rs][g] = (sum l, sum l^2)
template <typename T>
__global__ void __launch_bounds__(128)
k_pca_colstats(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RS, double* __restrict__ part) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int64_t rps = ceil_div64(N, RS);
  const int64_t r0 = blockIdx.y * rps, r1 = r0 + rps < N ? r0 + rps : N;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const double l = (double)pca_l2((float)Y[r * ldY + g]);
    s1 += l;
    s2 += l * l;
  }
  part[((int64_t)blockIdx.y * G + g) * 2] = s1;
  part[((int64_t)blockIdx.y * G + g) * 2 + 1] = s2;
}
// sums[g] = (sum l, sum l^2) over this rank's row slices (all-reduced over the ranks of a cell-sharded fit by the host code)
__global__ void k_pca_colstats_reduce(const double* __restrict__ part, int RS, int G, double* __restrict__ sums) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s1 = 0.0, s2 = 0.0;
  for (int r = 0; r < RS; ++r) {
    s1 += part[((int64_t)r * G + g) * 2];
    s2 += part[((int64_t)r * G + g) * 2 + 1];
  }
  sums[2 * g] = s1;
  sums[2 * g + 1] = s2;
}
// mean, 1 / sd from the global sums; *bad is set when a column is constant ("cannot rescale a constant/zero column to
// unit variance")
__global__ void k_pca_colstats_final(const double* __restrict__ sums, int G, double n, double* __restrict__ mean,
                                     double* __restrict__ inv_sd, int* __restrict__ bad) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const double s1 = sums[2 * g], s2 = sums[2 * g + 1];
  const double m = s1 / n;
  const double var = (s2 - n * m * m) / (n - 1.0);
  mean[g] = m;
  if (!(var > 1e-300)) {
    inv_sd[g] = 0.0;
    atomicOr(bad, 1);
  } else {
    inv_sd[g] = 1.0 / sqrt(var);
  }
}
// a = v * inv_sd, b = sum_g mean_g a_g  (one block)
__global__ void k_pca_prepare(const double* __restrict__ v, const double* __restrict__ mean, const double* __restrict__ inv_sd,
                              int G, double* __restrict__ a, double* __restrict__ b_out) {
  __shared__ double scratch[32];
  double b = 0.0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const double ag = v[g] * inv_sd[g];
    a[g] = ag;
    b += mean[g] * ag;
  }
  const double t = block_sum(b, scratch);
  if (threadIdx.x == 0) b_out[0] = t;
}
// t_n = sum_g l_ng a_g - b : one warp per cell
template <typename T>
__global__ void __launch_bounds__(256)
k_pca_rows(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, const double* __restrict__ a, const double* __restrict__ b,
           double* __restrict__ t) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  double acc = 0.0;
  for (int g = lane; g < G; g += 32) {
    const float y = (float)Y[n * ldY + g];
    if (y != 0.f) acc += (double)pca_l2(y) * a[g];
  }
  acc = warp_sum(acc);
  if (lane == 0) t[n] = acc - b[0];
}
// column slices: part[rs][g] = sum_n l_ng t_n ; tsum_part[rs] = sum_n t_n
template <typename T>
__global__ void __launch_bounds__(128)
k_pca_cols(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RS, const double* __restrict__ t,
           double* __restrict__ part, double* __restrict__ tsum_part) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t rps = ceil_div64(N, RS);
  const int64_t r0 = blockIdx.y * rps, r1 = r0 + rps < N ? r0 + rps : N;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += t[r];
    tsum_part[blockIdx.y] = s;
  }
  if (g >= G) return;
  double acc = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const float y = (float)Y[r * ldY + g];
    if (y != 0.f) acc += (double)pca_l2(y) * t[r];
  }
  part[(int64_t)blockIdx.y * G + g] = acc;
}
// qs[g] = sum_n l_ng t_n over this rank's slices, qs[G] = sum_n t_n (all-reduced over the ranks by the host code)
__global__ void k_pca_cols_reduce(const double* __restrict__ part, const double* __restrict__ tsum_part, int RS, int G,
                                  double* __restrict__ qs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  double q = 0.0;
  if (g == G) {
    for (int r = 0; r < RS; ++r) q += tsum_part[r];
  } else {
    for (int r = 0; r < RS; ++r) q += part[(int64_t)r * G + g];
  }
  qs[g] = q;
}
// w = X^T t = (q - mean sum(t)) / sd, v <- w / |w|; out2 = (|w|, 1 - |<v_new, v_old>|)  (one block; identical on every rank)
__global__ void k_pca_update(const double* __restrict__ qs, int G, const double* __restrict__ mean,
                             const double* __restrict__ inv_sd, double* __restrict__ v, double* __restrict__ w,
                             double* __restrict__ out2) {
  __shared__ double scratch[32];
  __shared__ double sh[2];
  const double ts = qs[G];
  double nn = 0.0, dot = 0.0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const double wg = (qs[g] - mean[g] * ts) * inv_sd[g];
    w[g] = wg;
    nn += wg * wg;
    dot += wg * v[g];
  }
  const double n2 = block_sum(nn, scratch);
  const double d = block_sum(dot, scratch);
  if (threadIdx.x == 0) { sh[0] = sqrt(n2); sh[1] = d; }
  __syncthreads();
  const double nrm = sh[0];
  for (int g = threadIdx.x; g < G; g += blockDim.x) v[g] = nrm > 0.0 ? w[g] / nrm : 0.0;
  if (threadIdx.x == 0) {
    out2[0] = nrm;
    out2[1] = nrm > 0.0 ? 1.0 - fabs(sh[1]) / nrm : 0.0;   // v_old has unit norm
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host-side initial values of R/inference-tflow.R:210-222 from the resident count matrix (ca_core_data_stats):
//   s_init = rowSums(Y) (:210), colSums(Y) (gene filter, :117), mu_guess = colMeans(Y / rowMeans(Y)) (:222)
// One warp per cell for the row sums, then column slices for the two column quantities; fp64, fixed order.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_stats_rows(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, double* __restrict__ rowsum) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  double acc = 0.0;
  for (int g = lane; g < G; g += 32) acc += (double)(float)Y[n * ldY + g];
  acc = warp_sum(acc);
  if (lane == 0) rowsum[n] = acc;
}
// rowSums(Y[, keep]) (R/preprocess.R:138): the cell filter of preprocess_for_clonealign counts only the genes that survived
// the gene filters; keep[g] != 0 marks them
template <typename T>
__global__ void __launch_bounds__(256)
k_stats_rows_masked(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, const unsigned char* __restrict__ keep,
                    double* __restrict__ rowsum) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  double acc = 0.0;
  for (int g = lane; g < G; g += 32)
    if (keep[g]) acc += (double)(float)Y[n * ldY + g];
  acc = warp_sum(acc);
  if (lane == 0) rowsum[n] = acc;
}
// part[rs][g] = (sum_n y_ng, sum_n y_ng / rowsum_n) over the row slice
template <typename T>
__global__ void __launch_bounds__(128)
k_stats_cols(const T* __restrict__ Y, int64_t ldY, int64_t N, int G, int RS, const double* __restrict__ rowsum,
             double* __restrict__ part) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int64_t rps = ceil_div64(N, RS);
  const int64_t r0 = blockIdx.y * rps, r1 = r0 + rps < N ? r0 + rps : N;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const float y = (float)Y[r * ldY + g];
    if (y != 0.f) {
      s1 += (double)y;
      s2 += (double)y / rowsum[r];
    }
  }
  part[((int64_t)blockIdx.y * G + g) * 2] = s1;
  part[((int64_t)blockIdx.y * G + g) * 2 + 1] = s2;
}

}  // namespace ca
