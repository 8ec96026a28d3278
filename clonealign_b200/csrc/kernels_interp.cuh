// "interp" path (K = 1, P = 0): both contractions collapse to UNIVARIATE functions.
//
// With a single latent dimension eta_ng = psi_n w_g, so
//     Zx[n][j]  = sum_g Mx[g][j] exp(psi_n w_g - m(psi_n)) = F_j(psi_n)     (normaliser, R/inference-tflow.R:288-290)
//     dMx[g][j] = sum_n Rx[n][j] exp(psi_n w_g - m_n)       = H_j(w_g)       (its reverse-mode gradient)
// where F_j, H_j are sums of exponentials: entire functions of ONE real variable.  Piecewise Chebyshev
// interpolation on panels whose exponent half-range is <= kAmax converges spectrally (24 nodes: ~1e-14 relative, 16 nodes: ~6e-9;
// scripts/interp_prototype.py, tests/test_interp_model.py), so the (N x G) x (G x J) contraction is replaced by
//     nodes:  (n_nodes x G) x (G x J)   with n_nodes ~ 50..200 instead of N = 100 000        (k_interp_nodes<FWD>)
//     DCT  :  node values -> Chebyshev coefficients per panel                                  (k_interp_coeffs)
//     eval :  one Clenshaw recurrence per (cell, column)                                       (k_interp_eval)
// and symmetrically for the backward pass (nodes over w, reduction over cells, evaluation per gene).
// Everything is fp64 (node exponentials, accumulation, recurrence); inputs/outputs keep the fp32 layouts of the
// CUDA-core path (Mx [G][J], Zx [N][J], Rx [N][J], dMx [G][J]), so the per-cell epilogue and the gene-gradient
// kernels are shared.  All reductions run in a fixed order (deterministic).
//
// The shift convention is the one used everywhere else: m(x) = max(x w_max, x w_min), i.e. the exponent is
// x (w_g - w_max) <= 0 for x >= 0 and x (w_g - w_min) <= 0 for x < 0: two smooth pieces, panelled separately.
#pragma once
#include "common.cuh"

namespace ca {

constexpr int kIP = 16;            // Chebyshev nodes per panel (multiple of 8); with kIAmax = 4: ~6e-9 relative
constexpr int kIMaxPanF = 64;      // forward panels (both signs of psi together)
constexpr int kIMaxPanB = 64;      // backward panels over [w_min, w_max]
constexpr int kISplitF = 32;       // fixed split of the gene reduction in the forward node kernel (4 active node groups x 32 = 128 blocks)
constexpr int kISplitB = 64;       // fixed split of the cell reduction in the backward node kernel
constexpr int kIGroupsY = 8;       // grid.y of the node kernels: blocks stride over the ACTIVE groups of 8 nodes
constexpr double kIAmax = 4.0;     // exponent half-range per panel
constexpr double kPi = 3.14159265358979323846;

struct InterpPlan {
  double wmin, wmax, pmin, pmax;
  int nf_neg, nf_pos;              // forward panels on [pmin, 0) and [0, pmax]
  double f_neg_w, f_pos_w;         // panel widths (0 when the side is empty)
  int nb;                          // backward panels on [wmin, wmax]
  double b_w;
};

// panel -> (mid, half) helpers
__device__ __forceinline__ void fwd_panel(const InterpPlan& pl, int pf, double& mid, double& half, double& wref) {
  if (pf < pl.nf_neg) {
    double lo = pl.pmin + pf * pl.f_neg_w;
    half = 0.5 * pl.f_neg_w;
    mid = lo + half;
    wref = pl.wmin;
  } else {
    double lo = (pf - pl.nf_neg) * pl.f_pos_w;
    half = 0.5 * pl.f_pos_w;
    mid = lo + half;
    wref = pl.wmax;
  }
}
__device__ __forceinline__ void bwd_panel(const InterpPlan& pl, int pb, double& mid, double& half) {
  half = 0.5 * pl.b_w;
  mid = pl.wmin + pb * pl.b_w + half;
}
__device__ __forceinline__ double cheb_node(int p) { return cos(kPi * (p + 0.5) / kIP); }

// panel structure from the current ranges of psi and w
__device__ __forceinline__ InterpPlan interp_make_plan(double wmin, double wmax, double pmin, double pmax) {
  InterpPlan pl;
  pl.wmin = wmin; pl.wmax = wmax;
  pl.pmin = pmin; pl.pmax = pmax;
  const double D = pl.wmax - pl.wmin;
  auto count = [](double width, double scale, int cap) {
    int n = (int)ceil(scale * width / 2.0 / kIAmax);
    return n < 1 ? 1 : (n > cap ? cap : n);
  };
  if (pl.pmin < 0.0) {
    pl.nf_neg = count(-pl.pmin, D, kIMaxPanF / 2);
    pl.f_neg_w = -pl.pmin / pl.nf_neg;
  } else {
    pl.nf_neg = 0; pl.f_neg_w = 0.0;
  }
  if (pl.pmax >= 0.0) {
    pl.nf_pos = count(pl.pmax, D, kIMaxPanF / 2);
    pl.f_pos_w = pl.pmax / pl.nf_pos;
  } else {
    pl.nf_pos = 0; pl.f_pos_w = 0.0;
  }
  const double A = fmax(fabs(pl.pmin), fabs(pl.pmax));
  pl.nb = count(D, A, kIMaxPanB);
  pl.b_w = D / pl.nb;
  return pl;
}
// one thread
__global__ void k_interp_plan(const float* __restrict__ mm_w, const float* __restrict__ mm_psi, InterpPlan* __restrict__ plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *plan = interp_make_plan((double)mm_w[0], (double)mm_w[1], (double)mm_psi[0], (double)mm_psi[1]);
}

// Node values.  FWD: part[z][node][j] = sum_{g in split z} Mx[g][j] exp(x_node (w_g - wref));  reduction index = genes.
//               BWD: part[z][node][j] = sum_{n in split z} Rx[n][j] exp(psi_n y_node - m_n);    reduction index = cells.
// Block = 8 warps, one group of 8 nodes x (32 NC) columns; warps stride over the reduction index, a lane owns NC
// columns (j = cblock + lane + 32 c) of all 8 nodes in registers; the 8 exponentials of a row are computed by 8 lanes
// and broadcast with shuffles, each of which now feeds NC fp64 FMAs (with NC = 1 the kernel was shuffle-bound: one
// SHFL per DFMA).  Cross-warp reduction through shared memory in warp order (deterministic).
// Grid = (column blocks, kIGroupsY, reduction splits): how many panels are active is only known on the device (the
// plan), so blocks stride over the active groups of 8 nodes instead of launching (and retiring) one block per possible
// group; the reduction index is split over blockIdx.z (kISplitF / kISplitB partials, summed by k_interp_coeffs) so that
// a handful of active groups still spreads over all SMs instead of running as a few long serial loops.
constexpr int kINodeMaxNC = 8;
template <bool FWD, int NC>
__global__ void __launch_bounds__(256)
k_interp_nodes(const InterpPlan* __restrict__ plan, const float* __restrict__ rv /*FWD: w[G]  BWD: psi[N]*/,
               const float* __restrict__ shift /*BWD: m[N]*/, const float* __restrict__ B /*[R][J]*/, int64_t R, int J,
               double* __restrict__ vals) {
  __shared__ double red[8][32 * NC];
  const InterpPlan pl = *plan;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int npan = FWD ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const int ngroups = npan * (kIP / 8);
  const int j0 = blockIdx.x * 32 * NC + lane;
  // reduction range of this block
  const int64_t per = (R + gridDim.z - 1) / gridDim.z;
  const int64_t rbeg = (int64_t)blockIdx.z * per;
  const int64_t rend = rbeg + per < R ? rbeg + per : R;
  const int64_t nodes_total = (int64_t)(FWD ? kIMaxPanF : kIMaxPanB) * kIP;
  for (int grp = blockIdx.y; grp < ngroups; grp += gridDim.y) {
    const int node0 = grp * 8;
    const int panel = node0 / kIP;
    double mid, half, wref = 0.0;
    if (FWD) fwd_panel(pl, panel, mid, half, wref);
    else bwd_panel(pl, panel, mid, half);
    const double xq = mid + half * cheb_node((node0 % kIP) + (lane & 7));   // node handled by this lane (lanes 0..7 used)
    double acc[8][NC];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[q][c] = 0.0;
    // 4 rows of the reduction index per warp iteration: lane (q = lane & 7, sub = lane >> 3) evaluates the exponential
    // of row r + sub at node q (argument in fp64, expf in fp32: ~1e-7 relative, averaged over the sum), then every lane
    // accumulates its columns in fp64
    for (int64_t r = rbeg + (int64_t)wid * 4; r < rend; r += 32) {
      const int64_t rr = r + (lane >> 3);
      float e = 0.f;
      if (rr < rend) {
        const double v = (double)rv[rr];
        e = FWD ? expf((float)(xq * (v - wref))) : expf((float)(v * xq - (double)shift[rr]));
      }
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const int64_t r2 = r + sub;
        double b[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) b[c] = (r2 < rend && j0 + 32 * c < J) ? (double)B[r2 * J + j0 + 32 * c] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double eq = (double)__shfl_sync(CA_FULL, e, q + 8 * sub);
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[q][c] = fma(eq, b[c], acc[q][c]);
        }
      }
    }
    // cross-warp sum in warp order
    for (int w = 0; w < 8; ++w) {
      if (wid == w) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            double* slot = &red[q][c * 32 + lane];
            *slot = (w == 0) ? acc[q][c] : *slot + acc[q][c];
          }
      }
      __syncthreads();
    }
    // thread (node q = wid, lane): its NC columns
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = j0 + 32 * c;
      if (j < J) vals[((int64_t)blockIdx.z * nodes_total + node0 + wid) * J + j] = red[wid][c * 32 + lane];
    }
    __syncthreads();   // red is reused by the next group
  }
}
// host: columns per lane and grid.x for J columns
inline int interp_nodes_nc(int J) { int nc = (J + 31) / 32; return nc > kINodeMaxNC ? kINodeMaxNC : (nc < 1 ? 1 : nc); }

// Chebyshev coefficients per panel: c_k = (2/P) sum_p f(x_p) cos(pi k (p + 1/2) / P), c_0 halved.
// The node values arrive as `nsplit` partials that are summed here in a fixed order.
// Block = (kIP nodes) x (32 columns) threads for one (column block, panel): thread (p, lane) sums the partials of node
// p (coalesced over the columns), the block transposes through shared memory, then thread (k = p, lane) applies the DCT.
__global__ void __launch_bounds__(kIP * 32)
k_interp_coeffs(const InterpPlan* __restrict__ plan, const double* __restrict__ vals, int nsplit, int max_pan, int J,
                int fwd, double* __restrict__ coeff) {
  __shared__ double f[kIP][32];
  __shared__ double ct[kIP][kIP];
  const InterpPlan pl = *plan;
  const int npan = fwd ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const int panel = blockIdx.y;
  if (panel >= npan) return;
  const int lane = threadIdx.x & 31, p = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int64_t nodes_total = (int64_t)max_pan * kIP;
  if (threadIdx.x < kIP * kIP) {
    const int k = threadIdx.x / kIP, q = threadIdx.x % kIP;
    ct[k][q] = cos(kPi * k * (q + 0.5) / kIP);
  }
  double acc = 0.0;
  if (j < J)
    for (int z = 0; z < nsplit; ++z) acc += vals[((int64_t)z * nodes_total + panel * kIP + p) * J + j];
  f[p][lane] = acc;
  __syncthreads();
  const int k = p;
  double c = 0.0;
#pragma unroll
  for (int q = 0; q < kIP; ++q) c += f[q][lane] * ct[k][q];
  c *= 2.0 / kIP;
  if (k == 0) c *= 0.5;
  if (j < J) coeff[((int64_t)panel * kIP + k) * J + j] = c;
}

// Evaluation: out[i][j] = sum_k c[panel(x_i)][k][j] T_k(t_i) by Clenshaw, one warp per point, lanes over columns.
// Persistent blocks keep the coefficients of all active panels in shared memory when they fit (the common case:
// 2..6 panels); otherwise they are read through L2.  FWD: points = cells (x = psi), BWD: points = genes (x = w).
constexpr int kIEvalWarps = 16;
template <bool FWD>
__global__ void __launch_bounds__(kIEvalWarps * 32)
k_interp_eval(const InterpPlan* __restrict__ plan, const double* __restrict__ coeff, const float* __restrict__ xs, int64_t n,
              int J, float* __restrict__ out, int smem_panels) {
  CA_DYNAMIC_SMEM(double, csm);
  const InterpPlan pl = *plan;
  const int npan = FWD ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const bool in_smem = npan <= smem_panels;
  const int64_t per_panel = (int64_t)kIP * J;
  if (in_smem) {
    for (int64_t i = threadIdx.x; i < npan * per_panel; i += blockDim.x) csm[i] = coeff[i];
    __syncthreads();
  }
  const double* cbase = in_smem ? csm : coeff;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t ibeg = (int64_t)blockIdx.x * chunk;
  const int64_t iend = ibeg + chunk < n ? ibeg + chunk : n;
  for (int64_t i = ibeg + wid; i < iend; i += kIEvalWarps) {
    const double x = (double)xs[i];
    int panel;
    double mid, half, wref;
    if (FWD) {
      if (x < 0.0) {
        panel = (int)((x - pl.pmin) / pl.f_neg_w);
        panel = panel < 0 ? 0 : (panel >= pl.nf_neg ? pl.nf_neg - 1 : panel);
      } else {
        panel = pl.f_pos_w > 0.0 ? (int)(x / pl.f_pos_w) : 0;
        panel = pl.nf_neg + (panel >= pl.nf_pos ? pl.nf_pos - 1 : panel);
      }
      panel = panel < 0 ? 0 : (panel >= npan ? (npan > 0 ? npan - 1 : 0) : panel);   // NaN psi: stay inside the table
      fwd_panel(pl, panel, mid, half, wref);
    } else {
      panel = pl.b_w > 0.0 ? (int)((x - pl.wmin) / pl.b_w) : 0;
      panel = panel < 0 ? 0 : (panel >= pl.nb ? pl.nb - 1 : panel);
      bwd_panel(pl, panel, mid, half);
    }
    const double t = half > 0.0 ? (x - mid) / half : 0.0;
    const double t2 = 2.0 * t;
    const double* c = cbase + (int64_t)panel * per_panel;
    for (int j = lane; j < J; j += 32) {
      double b1 = 0.0, b2 = 0.0;
#pragma unroll 4
      for (int k = kIP - 1; k >= 1; --k) {
        const double tmp = fma(t2, b1, c[(int64_t)k * J + j] - b2);
        b2 = b1;
        b1 = tmp;
      }
      out[i * J + j] = (float)(fma(t, b1, c[j] - b2));
    }
  }
}

}  // namespace ca
