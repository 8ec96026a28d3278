// Variant EPI2 of the interp path (K = 1, P = 0, C <= 32, S*C <= 128): the Clenshaw evaluation of the normaliser
// columns (kernels_interp.cuh, k_interp_eval<FWD>) and the per-cell epilogue (kernels_small.cuh, k_cell_epilogue) in
// ONE kernel, re-engineered from the round-1 ncu source counters of k_cell_epilogue (2 059 warp instructions per cell,
// issue-bound at 52 %):
//   * Z / Z' never round-trip through HBM (Zx [N][J] fp32 written by one kernel and read back by the next);
//   * the Y-pass row partials are summed by nCB lanes + one shuffle tree instead of a 10-iteration loop that every
//     lane repeated with 64-bit index math (170 instructions per cell);
//   * the clone softmax keeps one clone per lane in registers with fp32 exponentials / logarithm (the reference's own
//     precision: tf$nn$softmax on float32, R/inference-tflow.R:273) and an fp64 normalisation (so that 1 - gamma_max keeps
//     its digits), instead of fp64 exp/log loops over shared memory (~310 instructions);
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
  int defer_yv;                          // variant DEFER: the Y-linear terms are added later (k_adam_all / k_yv_dot)
  const InterpPlan* plan;
  const double* coeff;                   // [panel][kIP][J]
  const float* mm;                       // (w_min, w_max)
  const float *U, *Bm, *vA, *s, *log_alpha, *rowpart;
  float* t;                              // gamma_logits (written in INIT mode)
  float *gT, *Rx, *gU, *YV, *Fout, *shift;
  float* Zx;                             // optional inspection copy of (Z | Z') [N][J], nullptr in the timed path
  double *elbo_part, *gsum_part;         // one partial per block
};

inline size_t fused_smem_bytes(int SC, int C, int J, int smem_panels, int warps = kFusedWarps) {
  return ((size_t)smem_panels * kFusedPitch * J + (size_t)warps * SC + warps + (size_t)warps * C) * sizeof(double) + 16;
}
// how many panels of coefficients fit next to the per-warp scratch (0: read them through L2)
inline int fused_smem_panels(int SC, int C, int J, size_t budget = 200 * 1024, int warps = kFusedWarps) {
  const size_t fixed = fused_smem_bytes(SC, C, J, 0, warps);
  const size_t per_panel = (size_t)kFusedPitch * J * sizeof(double);
  if (fixed >= budget) return 0;
  size_t n = (budget - fixed) / per_panel;
  return (int)(n > (size_t)kIMaxPanF ? (size_t)kIMaxPanF : n);
}

// log of a positive fp32 value as a double: exponent and mantissa are split, the mantissa m in [0.707, 1.414) goes through
// the hardware lg2 (__logf: absolute error <= 2^-21.4 on [0.5, 2]) and the exponent is added in fp64.  The absolute error
// (3.6e-7) does not grow with |log z|, unlike logf's 1 ulp of the result (9.5e-7 at log z ~ 10) -- log Z is multiplied by
// the library size before the clone softmax -- and it costs ~9 instructions instead of ~30 (ncu of round 2: logf was 98
// of the 1090 warp instructions per cell).  Zero, denormal, infinite and NaN arguments take the library path.
__device__ __forceinline__ double log_pos_f32(float z) {
  if (!(z >= 1.17549435e-38f && z <= 3.4e38f)) return (double)logf(z);
  const int bits = __float_as_int(z);
  int e = ((bits >> 23) & 0xff) - 127;
  float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  if (m > 1.41421356f) { m *= 0.5f; e += 1; }
  return (double)__logf(m) + (double)e * 0.69314718055994530942;
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
  const int nwarps = blockDim.x >> 5;                                 // 32, or 16 when the block shares its SM with a Y-pass CTA
  double* blkE = sm + table + (size_t)nwarps * SC;                    // [warps]
  double* blkG = blkE + nwarps;                                       // [warps][C]
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
  for (int64_t n = ibeg + wid; n < iend; n += nwarps) {
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
    // a NaN psi (diverged fit) must yield NaN results, not an out-of-range table index
    panel = panel < 0 ? 0 : (panel >= npan ? (npan > 0 ? npan - 1 : 0) : panel);
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
        if (jok[i]) lz[jz[i]] = log_pos_f32(zf[i]) + m;
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
    // ---- gamma = softmax(t), one clone per lane: exponentials in fp32, normalisation in fp64 ----
    // d t_c = gamma_c (H_c - sum_k gamma_k H_k) cancels to (1 - gamma_A) H_A - ... for a confidently assigned cell: with
    // gamma rounded to fp32 (1 - gamma_A) keeps no significant digits once gamma_A > 1 - 1e-6 and the logits of such
    // cells random-walk under Adam.  exp(t_A - max) is exactly 1 and the fp64 sum keeps 1 - gamma_A = sum_{k != A} e_k / Z
    // to the relative accuracy of the small exponentials.
    const float tv = cok ? a.t[n * C + lane] : -3.0e38f;
    const float mx = warp_max(tv);
    const double ex = cok ? (double)expf(tv - mx) : 0.0;
    const double zs = warp_sum(ex);
    const float lg = tv - (mx + logf((float)zs));
    const double g = ex / zs;
    const double H = F + (double)la - (double)lg;
    const double gh = (g == 0.0) ? 0.0 : g * H;           // tf$where(gamma == 0, 0, gamma * log gamma), :333
    const double sumGH = warp_sum(gh);
    if (cok && a.Fout) a.Fout[n * C + lane] = (float)F;
    // ---- Y-linear term psi_n (YW)_n and the N(0,1) prior on psi (:318-319) ----
    // (variant DEFER: this kernel does not read the Y-pass partials at all, so that it can run before / next to the Y pass;
    // psi_n (YW)_n joins the ELBO in k_yv_dot and (YW)_n joins d psi_n in k_adam_all)
    double yv = 0.0;
    if (!a.defer_yv) {
      for (int cb = lane; cb < a.nCB; cb += 32) yv += (double)a.rowpart[(int64_t)cb * a.N + n];
      yv = warp_sum(yv);
      if (lane == 0 && a.YV) a.YV[n] = (float)yv;
    }
    elbo_w += sumGH + x * yv - 0.5 * x * x - 0.5 * kLog2Pi;
    if (MODE == EPI_TRAIN) {
      gacc += g;
      if (cok) a.gT[n * C + lane] = (g == 0.0) ? 0.f : (float)(g * (H - sumGH));
      // R_scn = gamma_nc s_n / (S Z_scn) and d psi_n = (YW)_n - sum_sc R Z' - psi_n
      double zp[NJ];
      if (in_smem) clenshaw_cols<NJ, true>(cpan, jz, SC, a.J, tt, zp);
      else clenshaw_cols<NJ, false>(cpan, jz, SC, a.J, tt, zp);
      const float sn_over_S = (float)(sn * invS);
      double gu = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const float gc = __shfl_sync(CA_FULL, (float)g, cj[i]);
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
    for (int w = 0; w < nwarps; ++w) e += blkE[w];
    a.elbo_part[blockIdx.x] = e;
  }
  if (MODE == EPI_TRAIN && threadIdx.x < C) {
    double gs = 0.0;
    for (int w = 0; w < nwarps; ++w) gs += blkG[(size_t)w * C + threadIdx.x];
    a.gsum_part[(int64_t)blockIdx.x * C + threadIdx.x] = gs;
  }
}


// =====================================================================================================================
// Variant LEAN (with EPI2): fewer, fatter launches for the gene-level / scalar work that every rank of a cell-sharded
// fit replicates and that therefore bounds strong scaling (profiles/r01_notes.md: ~0.25 ms per step independent of N).
//   k_prologue   = k_alpha + k_minmax(W) + k_wsq + k_minmax(psi) + k_interp_plan + k_sample_mu   (6 launches -> 1)
//   k_gene_fused = k_interp_eval<BWD> + k_gene_grads_warp + k_reduce_gsum                  (3 launches -> 1, no dMx round trip)
//   k_adam_all   = k_gene_adam + k_scalar_adam + k_cell_adam                               (3 launches -> 1)
// =====================================================================================================================
constexpr int kProThreads = 512;
constexpr int kProPsiBlocks = 30;   // psi min/max partials (blocks 2 .. 2 + kProPsiBlocks - 1)

struct PrologueArgs {
  int64_t N;
  int G, C, K;
  const float *u, *chi_raw, *Vm, *U;
  float *log_alpha, *mm, *mm_psi;
  double *scal_elbo, *wsq, *chi_cur;   // chi_cur[k] = exp(chi_raw[k]) of THIS step (read by k_adam_all)
  float* pmm_part;       // [kProPsiBlocks][2]
  unsigned* ticket;      // zero before the first launch; the last block resets it
  InterpPlan* plan;
  double dirichlet_const;   // C lgamma(1/C) - lgamma(1)   (host)
  StepState* state;         // draw counter read by every gene block, advanced (with lr_t of this step) by the last block
  double lr;                // learning rate
  SampleMuArgs mu;          // gene blocks (k_sample_mu): blocks 2 + kProPsiBlocks ..; gene_part has one partial per gene block
  int mu_vec4;              // C % 4 == 0: 16-byte stores of the contraction operand
  int wide_panels;          // CELL2 set: interp_make_plan(..., wide = true)
};

// roles by block: 0 = alpha / scalar priors (k_alpha), 1 = W range and sum of squares (k_minmax, k_wsq; K == 1),
// 2 .. 2 + kProPsiBlocks - 1 = psi range partials, the rest = gene blocks (k_sample_mu: draws, mu, contraction operand,
// gene-level ELBO terms); the block that arrives last combines the ranges into the panel plan (k_interp_plan).
__global__ void __launch_bounds__(kProThreads) k_prologue(PrologueArgs a) {
  __shared__ double dscr[32];
  __shared__ float smin[32], smax[32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned long long draw = a.state->draw;   // every block reads it before it arrives at the ticket below
  if (blockIdx.x == 0) {
    if (wid == 0) {
      // log_softmax(alpha_unconstr) (R/inference-tflow.R:255), Dirichlet(1/C) prior on alpha + 1e-3 (:324),
      // chi = exp(chi_raw) (:241) and its Gamma(2,1) prior (:315)
      double mx = -1e300;
      for (int c = lane; c < a.C; c += 32) mx = fmax(mx, (double)a.u[c]);
      mx = warp_max(mx);
      double z = 0.0;
      for (int c = lane; c < a.C; c += 32) z += exp((double)a.u[c] - mx);
      z = warp_sum(z);
      const double lzv = mx + log(z);
      double e = 0.0;
      for (int c = lane; c < a.C; c += 32) {
        const double la = (double)a.u[c] - lzv;
        a.log_alpha[c] = (float)la;
        e += (1.0 / a.C - 1.0) * log(exp(la) + 1e-3);
      }
      for (int k = lane; k < a.K; k += 32) {
        const double ch = exp((double)a.chi_raw[k]);
        a.chi_cur[k] = ch;
        e += (double)a.chi_raw[k] - ch;
      }
      e = warp_sum(e);
      if (lane == 0) a.scal_elbo[0] = e - a.dirichlet_const;
    }
  } else if (blockIdx.x == 1) {
    float mn = 3.4e38f, mx = -3.4e38f;
    double sq = 0.0;
    for (int g = threadIdx.x; g < a.G; g += blockDim.x) {
      const float v = a.Vm[g];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
      sq += (double)v * (double)v;
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { smin[wid] = mn; smax[wid] = mx; }
    const double tot = block_sum(sq, dscr);   // contains the barriers that also publish smin / smax
    if (threadIdx.x == 0) {
      for (int i = 1; i < nw; ++i) { mn = fminf(mn, smin[i]); mx = fmaxf(mx, smax[i]); }
      a.mm[0] = mn;
      a.mm[1] = mx;
      a.wsq[0] = tot;
    }
  } else if (blockIdx.x >= 2 + kProPsiBlocks) {
    const int gb = blockIdx.x - 2 - kProPsiBlocks;
    SampleMuArgs m = a.mu;
    m.draw = draw;
    const int ngb = (int)gridDim.x - 2 - kProPsiBlocks;
    if (a.mu_vec4) sample_mu_body<true>(m, gb, ngb, dscr, m.gene_part);
    else sample_mu_body<false>(m, gb, ngb, dscr, m.gene_part);
  } else {
    const int pb = blockIdx.x - 2;
    const int64_t per = (a.N + kProPsiBlocks - 1) / kProPsiBlocks;
    const int64_t beg = (int64_t)pb * per, end = beg + per < a.N ? beg + per : a.N;
    float mn = 3.4e38f, mx = -3.4e38f;
    for (int64_t n = beg + threadIdx.x; n < end; n += blockDim.x) {
      const float v = a.U[n];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { smin[wid] = mn; smax[wid] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < nw; ++i) { mn = fminf(mn, smin[i]); mx = fmaxf(mx, smax[i]); }
      a.pmm_part[2 * pb] = mn;
      a.pmm_part[2 * pb + 1] = mx;
    }
  }
  // last block to arrive builds the plan
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* pp = a.pmm_part;
    const volatile float* mw = a.mm;
    float mn = 3.4e38f, mx = -3.4e38f;
    for (int i = 0; i < kProPsiBlocks; ++i) { mn = fminf(mn, pp[2 * i]); mx = fmaxf(mx, pp[2 * i + 1]); }
    a.mm_psi[0] = mn;
    a.mm_psi[1] = mx;
    *a.plan = interp_make_plan((double)mw[0], (double)mw[1], (double)mn, (double)mx, a.wide_panels != 0);
    *a.ticket = 0u;
    // this forward pass has consumed draw `draw`; the optimiser step that may follow is step adam_t + 1
    a.state->draw = draw + 1ull;
    const double t = (double)(a.state->adam_t + 1);
    a.state->lr_t = (float)(a.lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// gene kernel: dMx[g][j] = H_j(w_g) by Clenshaw (k_interp_eval<BWD>) consumed in place by the gene-gradient reductions
// (k_gene_grads_warp); the last block reduces the sum-gamma partials into the allreduce buffer (k_reduce_gsum).
// One warp per gene, persistent blocks, coefficients of the active backward panels staged in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kGeneWarps = 16;
struct GeneFusedArgs {
  int G, C, S, SC, J, nRB, smem_panels;
  const InterpPlan* plan;
  const double* coeff;       // [panel][kIP][J] backward coefficients
  const float *Vm, *colpart, *mu, *sig, *eps, *lsd, *L;
  float *ar, *YtU, *dM_out;  // dM_out: inspection copy [G][J] or nullptr
  const double* gsum_part;   // [n_parts][C]
  int64_t n_parts;
};
inline size_t gene_fused_smem_bytes(int J, int smem_panels) { return (size_t)smem_panels * kFusedPitch * J * sizeof(double) + 16; }
inline int gene_fused_smem_panels(int J, size_t budget = 96 * 1024) {
  size_t n = budget / ((size_t)kFusedPitch * J * sizeof(double));
  return (int)(n > (size_t)kIMaxPanB ? (size_t)kIMaxPanB : n);
}

// Latency structure (ncu of round 2: 75 us at 26 % issue, stall samples spread over the Clenshaw chains, the loads of a
// gene's inputs issued behind them and the gather of the column partials): per gene a lane now owns NJ columns (template,
// as in k_cell_fused), loads sigma / L / eps / mu of its columns BEFORE the recurrences, and runs the 2 NJ Clenshaw chains
// (dM and dM') of all its columns interleaved; the column partials of the NEXT round of 16 genes are fetched while this
// round computes.
template <int NJ>
__global__ void __launch_bounds__(kGeneWarps * 32, 2) k_gene_fused(GeneFusedArgs a) {
  CA_DYNAMIC_SMEM(double, csm);
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (blockIdx.x == gridDim.x - 1) {   // role: sum_n gamma_nc partials -> float slots of the allreduce buffer
    float* out = a.ar + 3 * (int64_t)a.G;
    for (int c = 0; c < a.C; ++c) {
      double s = 0.0;
      for (int64_t i = threadIdx.x; i < a.n_parts; i += blockDim.x) s += a.gsum_part[i * a.C + c];
      const double t = block_sum(s, scratch);
      if (threadIdx.x == 0) out[c] = (float)t;
    }
    return;
  }
  const InterpPlan pl = *a.plan;
  const int npan = pl.nb;
  const bool in_smem = npan <= a.smem_panels;
  const int64_t per_panel = (int64_t)kIP * a.J;
  const double ih = pl.b_w > 0.0 ? 2.0 / pl.b_w : 0.0;
  const int nblk = gridDim.x - 1;
  const int SC = a.SC;
  // Y^T psi partials of the Y pass, colpart [nRB][G]: the 16 genes a block works on in one round are consecutive, so
  // the block sums their partials cooperatively -- thread (slice, gene) strides over the row blocks and reads 16
  // consecutive floats per row block (whole 32-byte sectors); slices are combined in a fixed order.  kGMaxPer loads per
  // thread are in flight at once (up to 32 x kGMaxPer row blocks; more are summed by a serial tail).
  constexpr int kGMaxPer = 8;
  __shared__ double cpart[32][kGeneWarps];
  const int gi = threadIdx.x % kGeneWarps, sl = threadIdx.x / kGeneWarps;            // 512 threads = 32 slices x 16 genes
  auto fetch = [&](int gbase, float (&v)[kGMaxPer]) {
    const bool ok = gbase + gi < a.G;
    const float* cp = a.colpart + gbase + gi;
#pragma unroll
    for (int u = 0; u < kGMaxPer; ++u) v[u] = (ok && sl + 32 * u < a.nRB) ? cp[(int64_t)(sl + 32 * u) * a.G] : 0.f;
  };
  // cell-independent per-lane constants: columns j = lane + 32 i, their (sample, clone)
  int jz[NJ], sj[NJ], cj[NJ];
  bool jok[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const int j = lane + 32 * i;
    jok[i] = j < SC;
    jz[i] = jok[i] ? j : SC - 1;
    sj[i] = jz[i] / a.C;
    cj[i] = jz[i] - sj[i] * a.C;
  }
  float vcur[kGMaxPer];
  fetch(blockIdx.x * kGeneWarps, vcur);
  if (in_smem) {
    for (int64_t i = threadIdx.x; i < npan * per_panel; i += blockDim.x) {
      const int j = (int)(i % a.J);
      const int64_t pk = i / a.J;
      csm[((pk / kIP) * a.J + j) * kFusedPitch + (int)(pk % kIP)] = a.coeff[i];
    }
  }
  for (int gbase = blockIdx.x * kGeneWarps; gbase < a.G; gbase += nblk * kGeneWarps) {   // block-uniform trip count
    {
      double part = 0.0;
#pragma unroll
      for (int u = 0; u < kGMaxPer; ++u) part += (double)vcur[u];
      if (gbase + gi < a.G)
        for (int rb = sl + 32 * kGMaxPer; rb < a.nRB; rb += 32) part += (double)a.colpart[(int64_t)rb * a.G + gbase + gi];
      cpart[sl][gi] = part;
    }
    __syncthreads();                                    // also publishes the staged coefficient table (first round)
    fetch(gbase + nblk * kGeneWarps, vcur);             // next round's partials travel while this round computes
    const int g = gbase + wid;
    if (g < a.G) {
      const float wf = a.Vm[g];
      const float lsdf = a.lsd[g];
      // inputs of this lane's columns, requested before the recurrences
      float sg[NJ], lc[NJ], ep[NJ], mu[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const int64_t o = (int64_t)sj[i] * a.G + g;
        sg[i] = a.sig[o]; ep[i] = a.eps[o]; mu[i] = a.mu[o];
        lc[i] = a.L[(int64_t)g * a.C + cj[i]];
      }
      const double x = (double)wf;
      int pb = (int)((x - pl.wmin) * ih * 0.5);
      pb = pb < 0 ? 0 : (pb >= pl.nb ? pl.nb - 1 : pb);
      const double tt = pl.b_w > 0.0 ? (x - (pl.wmin + pb * pl.b_w)) * ih - 1.0 : 0.0;
      const double* cpan = in_smem ? csm + (size_t)pb * a.J * kFusedPitch : a.coeff + (int64_t)pb * per_panel;
      double d[NJ], d2[NJ];
      if (in_smem) {
        clenshaw_cols<NJ, true>(cpan, jz, 0, a.J, tt, d);
        clenshaw_cols<NJ, true>(cpan, jz, SC, a.J, tt, d2);
      } else {
        clenshaw_cols<NJ, false>(cpan, jz, 0, a.J, tt, d);
        clenshaw_cols<NJ, false>(cpan, jz, SC, a.J, tt, d2);
      }
      const float sd = expf(lsdf);
      double aloc = 0.0, alsd = 0.0, gv = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        if (jok[i]) {
          const float df = (float)d[i], d2f = (float)d2[i];   // the unfused path rounds dMx to fp32: keep its numerics
          if (a.dM_out) {
            a.dM_out[(int64_t)g * a.J + jz[i]] = df;
            a.dM_out[(int64_t)g * a.J + SC + jz[i]] = d2f;
          }
          const double dx = -(double)sg[i] * (double)lc[i] * (double)df;
          aloc += dx;
          alsd += dx * (double)sd * (double)ep[i];
          gv -= (double)(mu[i] * lc[i]) * (double)d2f;
        }
      }
      aloc = warp_sum(aloc);
      alsd = warp_sum(alsd);
      gv = warp_sum(gv);
      const double acc = warp_sum(cpart[lane][wid]);
      if (lane == 0) {
        a.YtU[g] = (float)acc;
        a.ar[2 * (int64_t)a.G + g] = (float)(acc + gv);
        a.ar[g] = (float)aloc;
        a.ar[a.G + g] = (float)alsd;
      }
    }
    __syncthreads();   // cpart is rewritten by the next round
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// optimiser: gene-level (k_gene_adam), per-cell (k_cell_adam) and scalar (k_scalar_adam) Adam updates in one launch.
// Gene blocks read chi from chi_cur (written by k_prologue from the OLD chi_raw), so the scalar block may update
// chi_raw concurrently.
// ---------------------------------------------------------------------------------------------------------------------
struct AdamAllArgs {
  GeneAdamArgs ga;
  const double* chi_cur;
  ScalarAdamArgs sa;
  int64_t N;
  int C;
  float *t, *m_t, *v_t, *U, *m_U, *v_U;
  const float* gT;
  float* gU;                 // written back in full when the Y-linear term was deferred
  int n_gene_blocks;
  int64_t n_cell_blocks;
  int defer_yv, nCB;         // variant DEFER: d psi_n += (YW)_n = sum_cb rowpart[cb][n] (fixed order)
  const float* rowpart;
  float* YV;
  StepState* state;          // lr_t of this step (k_prologue); the scalar block advances adam_t
  // CELL2 set, unsharded fit: the gene kernel ran in front of the Y-pass join, (Y^T psi)_g = sum_rb colpart[rb][g] is added here
  // (a sharded fit adds it before the all-reduce: k_colpart_add); nullptr otherwise
  const float* colpart;
  int nRB;
  float* YtU;
  int t_done;                // the gamma logits were already updated by k_cell_fused2: the cell blocks only hold psi
};
__global__ void __launch_bounds__(256) k_adam_all(AdamAllArgs a) {
  const int64_t b = blockIdx.x;
  a.ga.h.lr_t = a.sa.h.lr_t = a.state->lr_t;   // nobody writes lr_t while this kernel runs
  if (b < a.n_gene_blocks) {
    const GeneAdamArgs& q = a.ga;
    const int g = (int)b * blockDim.x + threadIdx.x;
    if (g >= q.G) return;
    // This launch is all that is left behind the Y-pass join and it is a chain of dependent loads (ncu of round 2: 21 us for
    // 1.3e6 instructions): everything a gene needs is requested up front -- the first column partials, then the samples in
    // batches of 8 -- instead of one round trip per sample and per 8 partials.
    float cpv[8];
    if (a.colpart) {
#pragma unroll
      for (int u = 0; u < 8; ++u) cpv[u] = u < a.nRB ? a.colpart[(int64_t)u * q.G + g] : 0.f;
    }
    const float lsd = q.lsd[g];
    const float ar0 = q.ar[g], ar1 = q.ar[q.G + g], ar2 = q.ar[2 * (int64_t)q.G + g], wg = q.Vm[g];
    const double sd = exp((double)lsd), cs = (double)q.colsum[g];
    double gl = (double)ar0, gs = (double)ar1;
    for (int s0 = 0; s0 < q.S; s0 += 8) {
      float mu8[8], sg8[8], lm8[8], ep8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t o = (int64_t)(s0 + i < q.S ? s0 + i : q.S - 1) * q.G + g;
        mu8[i] = q.mu[o]; sg8[i] = q.sig[o]; lm8[i] = q.logmu[o]; ep8[i] = q.eps[o];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (s0 + i < q.S) {
          const double mu = (double)mu8[i], sg = (double)sg8[i];
          const double dmu = (cs - (double)lm8[i]) / ((double)q.S * mu);
          const double dx = sg * dmu + (1.0 - sg) / (double)q.S;
          gl += dx;
          gs += dx * sd * (double)ep8[i];
        }
      }
    }
    gs += 1.0;
    const float gloc = (float)gl, glsd = (float)gs;
    q.g_loc[g] = gloc;
    q.g_lsd[g] = glsd;
    double ytu = 0.0;
    if (a.colpart) {   // fixed order: row blocks ascending, 16 partials in flight
      const float* cp = a.colpart + g;
#pragma unroll
      for (int u = 0; u < 8; ++u) ytu += (double)cpv[u];
      int rb = 8;
      for (; rb + 16 <= a.nRB; rb += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = cp[(int64_t)(rb + u) * q.G];
#pragma unroll
        for (int u = 0; u < 16; ++u) ytu += (double)v[u];
      }
      for (; rb < a.nRB; ++rb) ytu += (double)cp[(int64_t)rb * q.G];
      a.YtU[g] = (float)ytu;
    }
    const float gw = (float)((double)ar2 + ytu - a.chi_cur[0] * (double)wg);   // K == 1, P == 0
    q.g_V[g] = gw;
    if (q.h.apply) {
      adam_update(q.loc[g], q.m_loc[g], q.v_loc[g], gloc, q.h);
      adam_update(q.lsd[g], q.m_lsd[g], q.v_lsd[g], glsd, q.h);
      adam_update(q.Vm[g], q.m_V[g], q.v_V[g], gw, q.h);
    }
  } else if (b < a.n_gene_blocks + a.n_cell_blocks) {
    // cell blocks: thread i < nt4 updates 4 consecutive gamma logits (16-byte loads / stores: a quarter of the threads,
    // 4x the bytes in flight per thread -- the kernel is latency-bound), the threads behind them one psi each
    const int64_t i = (b - a.n_gene_blocks) * blockDim.x + threadIdx.x;
    const int64_t nt = a.N * a.C, nt4 = a.t_done ? 0 : (nt + 3) / 4;
    if (i < nt4) {
      if (a.ga.h.apply) {
        const int64_t e0 = 4 * i;
        if (e0 + 4 <= nt) {
          float4 t4 = reinterpret_cast<float4*>(a.t)[i], m4 = reinterpret_cast<float4*>(a.m_t)[i], v4 = reinterpret_cast<float4*>(a.v_t)[i];
          const float4 g4 = reinterpret_cast<const float4*>(a.gT)[i];
          adam_update(t4.x, m4.x, v4.x, g4.x, a.ga.h);
          adam_update(t4.y, m4.y, v4.y, g4.y, a.ga.h);
          adam_update(t4.z, m4.z, v4.z, g4.z, a.ga.h);
          adam_update(t4.w, m4.w, v4.w, g4.w, a.ga.h);
          reinterpret_cast<float4*>(a.t)[i] = t4;
          reinterpret_cast<float4*>(a.m_t)[i] = m4;
          reinterpret_cast<float4*>(a.v_t)[i] = v4;
        } else {
          for (int64_t e = e0; e < nt; ++e) adam_update(a.t[e], a.m_t[e], a.v_t[e], a.gT[e], a.ga.h);
        }
      }
    } else if (i - nt4 < a.N) {
      const int64_t n = i - nt4;
      float g = a.gU[n];
      if (a.defer_yv) {
        double yv = 0.0;
        for (int cb = 0; cb < a.nCB; ++cb) yv += (double)a.rowpart[(int64_t)cb * a.N + n];
        g = (float)((double)g + yv);
        a.gU[n] = g;
        if (a.YV) a.YV[n] = (float)yv;
      }
      if (a.ga.h.apply) adam_update(a.U[n], a.m_U[n], a.v_U[n], g, a.ga.h);
    }
  } else if (threadIdx.x == 0) {
    const ScalarAdamArgs& q = a.sa;
    const double chi = exp((double)q.chi_raw[0]);
    q.g_chi[0] = (float)(-0.5 * chi * q.wsq[0] + 0.5 * (double)q.G + 1.0 - chi);
    double mx = -1e300;
    for (int c = 0; c < q.C; ++c) mx = fmax(mx, (double)q.u[c]);
    double z = 0.0;
    for (int c = 0; c < q.C; ++c) z += exp((double)q.u[c] - mx);
    double rsum = 0.0;
    for (int c = 0; c < q.C; ++c) {
      const double al = exp((double)q.u[c] - mx) / z;
      rsum += al / (al + 1e-3);
    }
    for (int c = 0; c < q.C; ++c) {
      const double al = exp((double)q.u[c] - mx) / z;
      const double r = al / (al + 1e-3);
      q.g_u[c] = (float)((double)q.gsum[c] - q.n_total * al + (1.0 / q.C - 1.0) * (r - al * rsum));
    }
    if (q.h.apply) {
      adam_update(q.chi_raw[0], q.m_chi[0], q.v_chi[0], q.g_chi[0], q.h);
      for (int c = 0; c < q.C; ++c) adam_update(q.u[c], q.m_u[c], q.v_u[c], q.g_u[c], q.h);
      a.state->adam_t += 1;                     // read only by the next k_prologue (stream order)
    }
  }
}

// variant DEFER, ELBO evaluations: sum_n psi_n (YW)_n from the Y-pass row partials, one fixed-order partial per block
__global__ void __launch_bounds__(256) k_yv_dot(int64_t N, int nCB, const float* __restrict__ rowpart, const float* __restrict__ U,
                                                float* __restrict__ YV, double* __restrict__ part_out) {
  __shared__ double scratch[32];
  const int64_t chunk = (N + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk < N ? beg + chunk : N;
  double acc = 0.0;
  for (int64_t n = beg + threadIdx.x; n < end; n += blockDim.x) {
    double yv = 0.0;
    for (int cb = 0; cb < nCB; ++cb) yv += (double)rowpart[(int64_t)cb * N + n];
    if (YV) YV[n] = (float)yv;
    acc += (double)U[n] * yv;
  }
  const double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) part_out[blockIdx.x] = t;
}

}  // namespace ca
