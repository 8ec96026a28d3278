// Variant EPI2 of the interp path (K = 1, P = 0, C <= 32, S*C <= 128): the Clenshaw evaluation of the normaliser
// columns (kernels_interp.cuh, k_interp_eval<FWD>) and the per-cell epilogue (kernels_small.cuh, k_cell_epilogue) in
// ONE kernel, re-engineered from the round-1 ncu source counters of k_cell_epilogue (2 059 warp instructions per cell,
// issue-bound at 52 %):
//   * Z / Z' never round-trip through HBM (Zx [N][J] fp32 written by one kernel and read back by the next);
//   * the Y-pass row partials are summed by nCB lanes + one shuffle tree instead of a 10-iteration loop that every
//     lane repeated with 64-bit index math (170 instructions per cell);
//   * the clone softmax runs in fp32 (the reference's own precision: tf$nn$softmax on float32, R/inference-tflow.R:273)
//     with one lane per clone held in registers, instead of fp64 exp/log loops over shared memory (~310 instructions);
//   * column -> clone indices, panel constants and the coefficient table are hoisted out of the per-cell loop
//     (persistent blocks, one warp per cell, the table of the active panels staged once per block in shared memory);
//   * the per-cell exponent shift m_n = max(psi_n w_min, psi_n w_max) is computed here (k_shift_k1 is not launched).
// Outputs are those of k_cell_epilogue on the CUDA-core layout (Rx [N][J], gT, gU, YV, F, partial sums per block in a
// fixed order), so the backward node kernel, the gene-gradient kernels and the optimiser are shared.
// Math: SURVEY.md App. A.2/A.3; reference graph nodes R/inference-tflow.R:272-273,288-308,318-319,332-340.
#pragma once
#include "common.cuh"
#include "kernels_interp.cuh"
#include "kernels_small.cuh"

namespace ca {

constexpr int kFusedWarps = 32;          // one block of 1024 threads per SM
constexpr int kFusedMaxNJ = 4;           // S*C <= 128
constexpr int kFusedMaxC = 32;
constexpr int kFusedPitch = kIP + 1;     // doubles per (panel, column) row of the staged coefficient table

struct FusedArgs {
  int64_t N;
  int C, S, SC, J, nCB, smem_panels;
  const InterpPlan* plan;
  const double* coeff;                   // [panel][kIP][J]
  const float* mm;                       // (w_min, w_max)
  const float *U, *Bm, *vA, *s, *log_alpha, *rowpart;
  float* t;                              // gamma_logits (written in INIT mode)
  float *gT, *Rx, *gU, *YV, *Fout, *shift;
  float* Zx;                             // optional inspection copy of (Z | Z') [N][J], nullptr in the timed path
  double *elbo_part, *gsum_part;         // one partial per block
};

inline size_t fused_smem_bytes(int SC, int C, int J, int smem_panels) {
  return ((size_t)smem_panels * kFusedPitch * J + (size_t)kFusedWarps * SC + kFusedWarps + (size_t)kFusedWarps * C) * sizeof(double) + 16;
}
// how many panels of coefficients fit next to the per-warp scratch (0: read them through L2)
inline int fused_smem_panels(int SC, int C, int J, size_t budget = 200 * 1024) {
  const size_t fixed = fused_smem_bytes(SC, C, J, 0);
  const size_t per_panel = (size_t)kFusedPitch * J * sizeof(double);
  if (fixed >= budget) return 0;
  size_t n = (budget - fixed) / per_panel;
  return (int)(n > (size_t)kIMaxPanF ? (size_t)kIMaxPanF : n);
}

// One Clenshaw pass over NJ columns per lane.  SMEM: table staged as [panel][j][kIP + 1] -- a chain reads consecutive
// doubles at compile-time offsets from its per-lane base (no address arithmetic in the recurrence; the odd row pitch
// keeps the 64-bit reads of a half-warp on distinct banks).  Otherwise: coefficients through L2 in their [k][J] layout.
template <int NJ, bool SMEM>
__device__ __forceinline__ void clenshaw_cols(const double* base, const int (&jz)[NJ], int col0, int J, double tt,
                                              double (&out)[NJ]) {
  const double t2 = 2.0 * tt;
  double b1[NJ], b2[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) b1[i] = b2[i] = 0.0;
  if (SMEM) {
    const double* p[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) p[i] = base + (size_t)(col0 + jz[i]) * kFusedPitch;
#pragma unroll
    for (int k = kIP - 1; k >= 1; --k) {
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const double tz = fma(t2, b1[i], p[i][k] - b2[i]);
        b2[i] = b1[i];
        b1[i] = tz;
      }
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) out[i] = fma(tt, b1[i], p[i][0] - b2[i]);
  } else {
#pragma unroll 1
    for (int k = kIP - 1; k >= 1; --k) {
      const double* ck = base + (int64_t)k * J + col0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const double tz = fma(t2, b1[i], ck[jz[i]] - b2[i]);
        b2[i] = b1[i];
        b1[i] = tz;
      }
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) out[i] = fma(tt, b1[i], base[col0 + jz[i]] - b2[i]);
  }
}

template <int MODE, int NJ>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) k_cell_fused(FusedArgs a) {
  CA_DYNAMIC_SMEM(double, sm);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int SC = a.SC, C = a.C;
  const size_t table = (size_t)a.smem_panels * a.J * kFusedPitch;
  double* csm = sm;                                                   // [smem_panels][J][kIP + 1]
  double* lz = sm + table + (size_t)wid * SC;                         // per warp: log Z + m
  double* blkE = sm + table + (size_t)kFusedWarps * SC;               // [warps]
  double* blkG = blkE + kFusedWarps;                                  // [warps][C]
  // panel geometry of this step (k_interp_plan): per side the panel width and 2 / width
  int nf_neg, nf_pos;
  double pmin, w_neg, w_pos, ih_neg, ih_pos;
  {
    const InterpPlan pl = *a.plan;
    nf_neg = pl.nf_neg; nf_pos = pl.nf_pos; pmin = pl.pmin;
    w_neg = pl.f_neg_w; w_pos = pl.f_pos_w;
    ih_neg = w_neg > 0.0 ? 2.0 / w_neg : 0.0;
    ih_pos = w_pos > 0.0 ? 2.0 / w_pos : 0.0;
  }
  const int npan = nf_neg + nf_pos;
  const bool in_smem = npan <= a.smem_panels;
  const int64_t per_panel = (int64_t)kIP * a.J;
  if (in_smem) {   // transpose [panel][k][j] -> [panel][j][k] with pitch kIP + 1 while staging
    for (int64_t i = threadIdx.x; i < npan * per_panel; i += blockDim.x) {
      const int j = (int)(i % a.J);
      const int64_t pk = i / a.J;               // panel * kIP + k
      const int k = (int)(pk % kIP);
      const int64_t pan = pk / kIP;
      csm[(pan * a.J + j) * kFusedPitch + k] = a.coeff[i];
    }
    __syncthreads();
  }

  // cell-independent per-lane constants
  int jz[NJ], cj[NJ];
  bool jok[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const int j = lane + 32 * i;
    jok[i] = j < SC;
    jz[i] = jok[i] ? j : SC - 1;
    cj[i] = jz[i] % C;
  }
  const bool cok = lane < C;
  const float la = cok ? a.log_alpha[lane] : 0.f;
  const float wmin = a.mm[0], wmax = a.mm[1];
  const double invS = 1.0 / (double)a.S;

  double elbo_w = 0.0, gacc = 0.0;
  const int64_t chunk = (a.N + gridDim.x - 1) / gridDim.x;
  const int64_t ibeg = (int64_t)blockIdx.x * chunk;
  const int64_t iend = ibeg + chunk < a.N ? ibeg + chunk : a.N;
  for (int64_t n = ibeg + wid; n < iend; n += kFusedWarps) {
    const float psif = a.U[n];
    const float mf = fmaxf(psif * wmin, psif * wmax);
    if (lane == 0) a.shift[n] = mf;
    const double x = (double)psif, m = (double)mf, sn = (double)a.s[n];
    // ---- panel and Chebyshev argument: tt = (x - lo) * 2 / width - 1 ----
    int panel;
    double tt;
    if (x < 0.0) {
      int pf = (int)((x - pmin) * ih_neg * 0.5);
      pf = pf < 0 ? 0 : (pf >= nf_neg ? nf_neg - 1 : pf);
      tt = (x - (pmin + pf * w_neg)) * ih_neg - 1.0;
      panel = pf;
    } else {
      int pf = (int)(x * ih_pos * 0.5);
      pf = pf >= nf_pos ? nf_pos - 1 : pf;
      tt = w_pos > 0.0 ? (x - pf * w_pos) * ih_pos - 1.0 : 0.0;
      panel = nf_neg + pf;
    }
    const double* cpan = in_smem ? csm + (size_t)panel * a.J * kFusedPitch : a.coeff + (int64_t)panel * per_panel;
    // ---- Z columns ----
    float zf[NJ];
    {
      double z[NJ];
      if (in_smem) clenshaw_cols<NJ, true>(cpan, jz, 0, a.J, tt, z);
      else clenshaw_cols<NJ, false>(cpan, jz, 0, a.J, tt, z);
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        zf[i] = (float)z[i];
        if (jok[i]) lz[jz[i]] = (double)logf(zf[i]) + m;
        if (a.Zx && jok[i]) a.Zx[n * a.J + jz[i]] = zf[i];
      }
    }
    __syncwarp();
    // ---- F_nc (lane = clone) ----
    double F = 0.0;
    if (cok) {
      double acc = 0.0;
      for (int s = 0; s < a.S; ++s) acc += lz[s * C + lane];
      const double bv = (double)a.Bm[n * C + lane] + (double)a.vA[n * C + lane];
      F = (MODE == EPI_INIT) ? (double)a.S * bv - sn * acc : bv - sn * acc * invS;
    }
    __syncwarp();   // lz is rewritten by the next cell of this warp
    if (MODE == EPI_INIT) {
      // gamma_init: t <- F - logsumexp_c F   (sum over s, R/inference-tflow.R:338-340); runs once per fit: fp64
      double mx = warp_max(cok ? F : -1e300);
      double z = warp_sum(cok ? exp(F - mx) : 0.0);
      if (cok) a.t[n * C + lane] = (float)(F - (mx + log(z)));
      continue;
    }
    // ---- gamma = softmax(t) in fp32, one clone per lane ----
    const float tv = cok ? a.t[n * C + lane] : -3.0e38f;
    const float mx = warp_max(tv);
    const float ex = cok ? expf(tv - mx) : 0.f;
    const float zs = warp_sum(ex);
    const float lg = tv - (mx + logf(zs));
    const float g = cok ? ex / zs : 0.f;
    const double H = F + (double)la - (double)lg;
    const double gh = (g == 0.f) ? 0.0 : (double)g * H;   // tf$where(gamma == 0, 0, gamma * log gamma), :333
    const double sumGH = warp_sum(gh);
    if (cok && a.Fout) a.Fout[n * C + lane] = (float)F;
    // ---- Y-linear term psi_n (YW)_n and the N(0,1) prior on psi (:318-319) ----
    double yv = 0.0;
    for (int cb = lane; cb < a.nCB; cb += 32) yv += (double)a.rowpart[(int64_t)cb * a.N + n];
    yv = warp_sum(yv);
    if (lane == 0 && a.YV) a.YV[n] = (float)yv;
    elbo_w += sumGH + x * yv - 0.5 * x * x - 0.5 * kLog2Pi;
    if (MODE == EPI_TRAIN) {
      gacc += (double)g;
      if (cok) a.gT[n * C + lane] = (g == 0.f) ? 0.f : (float)((double)g * (H - sumGH));
      // R_scn = gamma_nc s_n / (S Z_scn) and d psi_n = (YW)_n - sum_sc R Z' - psi_n
      double zp[NJ];
      if (in_smem) clenshaw_cols<NJ, true>(cpan, jz, SC, a.J, tt, zp);
      else clenshaw_cols<NJ, false>(cpan, jz, SC, a.J, tt, zp);
      const float sn_over_S = (float)(sn * invS);
      double gu = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const float gc = __shfl_sync(CA_FULL, g, cj[i]);
        if (jok[i]) {
          const float r = __fdividef(gc * sn_over_S, zf[i]);
          a.Rx[n * a.J + jz[i]] = r;
          a.Rx[n * a.J + SC + jz[i]] = psif * r;
          gu += (double)r * zp[i];
          if (a.Zx) a.Zx[n * a.J + SC + jz[i]] = (float)zp[i];
        }
      }
      gu = warp_sum(gu);
      if (lane == 0) a.gU[n] = (float)(yv - gu - x);
    }
  }
  if (MODE == EPI_INIT) return;
  // ---- per-block partials in a fixed order ----
  if (lane == 0) blkE[wid] = elbo_w;
  if (MODE == EPI_TRAIN && cok) blkG[(size_t)wid * C + lane] = gacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int w = 0; w < kFusedWarps; ++w) e += blkE[w];
    a.elbo_part[blockIdx.x] = e;
  }
  if (MODE == EPI_TRAIN && threadIdx.x < C) {
    double gs = 0.0;
    for (int w = 0; w < kFusedWarps; ++w) gs += blkG[(size_t)w * C + threadIdx.x];
    a.gsum_part[(int64_t)blockIdx.x * C + threadIdx.x] = gs;
  }
}

}  // namespace ca
