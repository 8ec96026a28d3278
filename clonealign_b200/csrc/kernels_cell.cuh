// Variant CELL2 of the interp path (with EPI2 + LEAN + DEFER, S <= 8): the per-cell and per-gene kernels re-laid-out from the
// round-2 ncu source counters of k_cell_fused (kernels_fused.cuh): 1 110 warp instructions per cell at config 3, of which only
// 276 were the Clenshaw recurrences -- the rest was the lane = column mapping paying for itself: 38 shuffles, a shared-memory
// round trip of log Z to regroup (sample, clone) columns by clone, 12 of 32 lanes active in the clone softmax, an IEEE fp64
// division, two library logarithms, 117 IMAD of address arithmetic.
//
//   * lane = (cell, clone): a warp works on 32 / WC cells at once (WC = lanes per cell, the power of two >= C), every lane
//     owns ALL S samples of its clone.  sum_s log Z_scn, gamma_nc, R_scn and sum_s R Z' are then in-lane; the only
//     cross-lane steps are the clone softmax and two sums over the WC lanes of a cell (log2 WC shuffle levels each).
//   * the interpolants are evaluated in the MONOMIAL basis of the panel variable (Horner: one DFMA per coefficient instead
//     of DADD + DFMA for Clenshaw; k_interp_coeffs3 converts the Chebyshev coefficients, kernels_interp.cuh), coefficients
//     staged as (a_2k, a_2k+1) pairs: one 16-byte shared-memory load per two Horner steps, at compile-time offsets from
//     one per-lane base (table padded to [panel][pair][SB samples][WC lanes]).
//   * the w-weighted column Z'_sc = sum_g w_g M e^(psi w_g - m) is NOT interpolated: with m = psi w_ref linear on a side,
//     Z' = dZ/dpsi + w_ref Z, and dZ/dpsi is the derivative of the interpolant, carried in the same Horner pass (each
//     coefficient is loaded once and feeds p and dp/dt).  Half the node sums, half the table, half the shared-memory
//     traffic of the kernel (which was bound by it: 4 wavefronts per 16-byte load of a warp); the same holds for
//     dM' = d(dM)/dw in the gene kernel.  Accuracy: interp_make_plan(..., wide) in kernels_interp.cuh.
//   * sum_s log Z_s = log prod_s Z_s: one split-exponent logarithm per 4 samples on the fp64 product instead of one per
//     sample (absolute error 3.6e-7 per logarithm, see log_pos_f32 in kernels_fused.cuh).
//   * gamma = e / sum e through a Newton-refined hardware reciprocal (2 DFMA + 2 DMUL, relative error < 1e-15) instead of
//     the IEEE division subroutine; exponentials in fp32, normalisation in fp64 as before (1 - gamma_max keeps its digits).
//   * the per-cell inputs of the NEXT group of cells are requested before this group's recurrences start; the TF1-Adam
//     update of the gamma logits is applied by the lane that computed the gradient.
// Outputs: Rx [N][Jn] (Jn = S*C rounded up to even: R only), gT or the updated logits, gU without the Y-linear term, shift,
// per-block partial sums in a fixed order; the backward node kernel and the optimiser are shared with the other kernel sets.
// Math: SURVEY.md App. A.2/A.3; reference graph nodes R/inference-tflow.R:272-273,288-308,318-319,332-340.
#pragma once
#include "common.cuh"
#include "kernels_interp.cuh"
#include "kernels_fused.cuh"

namespace ca {

constexpr int kCell2MaxS = 8;            // samples a lane keeps in registers

struct Cell2Args {
  int64_t N;
  int C, S, SC, J, Jn, smem_panels;      // Jn: S*C rounded up to even = row pitch of Rx and of the coefficient table
  const InterpPlan* plan;
  const double2* coef2;                  // monomial pairs of the S*C normaliser columns [panel][kIP / 2][Jn] (k_interp_coeffs3)
  const float* mm;                       // (w_min, w_max)
  const float *U, *Bm, *vA, *s, *log_alpha;
  float* t;                              // gamma_logits (written in INIT mode)
  float *gT, *Rx, *gU, *Fout, *shift;
  float* Zx;                             // optional inspection copy of (Z | Z') [N][J], nullptr in the timed path
  double *elbo_part, *gsum_part;         // one partial per block
  // TRAIN mode with apply_t != 0: the TF1-Adam update of the gamma logits happens HERE, by the lane that has just computed the
  // gradient (d t depends on nothing outside this cell), instead of writing g_t for k_adam_all to read back: 1.2 M of the 1.3 M
  // elements that kernel updates at config 3, off the critical tail behind the Y-pass join.  The per-cell kernel is bound by
  // shared-memory latency, not by issue slots (ncu of round 2), so the update is close to free here (it was not in k_cell_fused).
  // lr_t of the step comes from StepState (k_prologue).
  int apply_t;
  float *m_t, *v_t;
  const StepState* state;
};

inline size_t cell2_panel_bytes(int WC, int SB) { return (size_t)(kIP / 2) * SB * WC * sizeof(double2); }
inline size_t cell2_smem_bytes(int WC, int SB, int C, int smem_panels, int warps) {
  return (size_t)smem_panels * cell2_panel_bytes(WC, SB) + ((size_t)warps + (size_t)warps * C) * sizeof(double) + 16;
}
inline int cell2_smem_panels(int WC, int SB, int C, size_t budget, int warps) {
  const size_t fixed = cell2_smem_bytes(WC, SB, C, 0, warps);
  if (fixed >= budget) return 0;
  const size_t n = (budget - fixed) / cell2_panel_bytes(WC, SB);
  return (int)(n > (size_t)kIMaxPanF ? (size_t)kIMaxPanF : n);
}
inline int cell2_pick_wc(int C) { return C <= 4 ? 4 : (C <= 8 ? 8 : (C <= 16 ? 16 : 32)); }
inline int cell2_pick_sb(int S) { return S <= 1 ? 1 : (S <= 2 ? 2 : (S <= 4 ? 4 : 8)); }

// log of a positive double: exponent and mantissa are split, the mantissa (rounded to fp32, in [0.707, 1.414]) goes through
// the hardware lg2, the exponent is added in fp64 (log_pos_f32 of kernels_fused.cuh for a double argument: the products
// of up to four normaliser values leave the fp32 range).  Zero, negative, tiny, huge and NaN arguments take the library path.
__device__ __forceinline__ double log_pos_f64(double z) {
  if (!(z >= 1e-290 && z <= 1e290)) return log(z);
  const int hi = __double2hiint(z);
  int e = ((hi >> 20) & 0x7ff) - 1023;
  float m = (float)__hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(z));
  if (m > 1.41421356f) { m *= 0.5f; e += 1; }
  return (double)__logf(m) + (double)e * 0.69314718055994530942;
}

// sums / maxima over the WC lanes of one cell (xor butterflies stay inside the aligned lane group; every lane of the group
// ends up with the same bits: each level adds the same two operands on both partners)
template <int WC, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = WC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(CA_FULL, v, o);
  return v;
}
template <int WC, typename T>
__device__ __forceinline__ T group_max(T v) {
#pragma unroll
  for (int o = WC / 2; o > 0; o >>= 1) {
    const T w = __shfl_xor_sync(CA_FULL, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// Horner evaluation of SBB samples' interpolants of one clone -- value p and, when DERIV, derivative dp / dt in the same pass
// (dp <- dp t + p; p <- p t + a_m: every coefficient is loaded once and feeds both recurrences) -- from the staged coefficient
// pairs (a_2k, a_2k+1), table [pair][SB][WC]: every offset is a compile-time constant.  Samples beyond S repeat the last one.
template <bool DERIV, int WC, int SB, int SBB>
__device__ __forceinline__ void cell2_horner(const double2* __restrict__ tb, double tt, int s0, int S, double (&p)[SBB], double (&dp)[SBB]) {
  constexpr int KP = kIP / 2;
  int so[SBB];
#pragma unroll
  for (int i = 0; i < SBB; ++i) so[i] = (s0 + i < S ? s0 + i : S - 1) * WC;
#pragma unroll
  for (int kp = KP - 1; kp >= 0; --kp) {
#pragma unroll
    for (int i = 0; i < SBB; ++i) {
      const double2 cz = tb[kp * SB * WC + so[i]];
      if (kp == KP - 1) {
        if (DERIV) dp[i] = cz.y;                                  // p = a_15 -> dp = a_15, p = a_15 t + a_14
        p[i] = fma(cz.y, tt, cz.x);
      } else {
        if (DERIV) dp[i] = fma(dp[i], tt, p[i]);
        p[i] = fma(p[i], tt, cz.y);
        if (DERIV) dp[i] = fma(dp[i], tt, p[i]);
        p[i] = fma(p[i], tt, cz.x);
      }
    }
  }
}

template <int MODE, int WC, int SB>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) k_cell_fused2(Cell2Args a) {
  CA_DYNAMIC_SMEM(double2, sm2);
  constexpr int CPW = 32 / WC;                  // cells per warp
  constexpr int SBB = SB < 4 ? SB : 4;          // samples per Horner batch (2 SBB independent chains per lane)
  constexpr int KP = kIP / 2;                   // coefficient pairs
  constexpr int kPanelStride = KP * SB * WC;       // double2 per staged panel
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / WC, cl = lane % WC;
  const int C = a.C, S = a.S, SC = a.SC, J = a.J;
  const bool cok = cl < C;
  const int c = cok ? cl : C - 1;               // idle lanes shadow the last clone (reads stay in range, results are masked)
  double* blkE = reinterpret_cast<double*>(sm2 + (size_t)a.smem_panels * kPanelStride);   // [warps]
  double* blkG = blkE + nwarps;                                                          // [warps][C]
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
  const float la = a.log_alpha[c];
  const float wmin = a.mm[0], wmax = a.mm[1];
  const double invS = 1.0 / (double)S;
  const float lr_t = (MODE == EPI_TRAIN && a.apply_t) ? a.state->lr_t : 0.f;

  double elbo_w = 0.0, gacc = 0.0;
  const int64_t chunk = (a.N + gridDim.x - 1) / gridDim.x;
  const int64_t ibeg = (int64_t)blockIdx.x * chunk;
  const int64_t iend = ibeg + chunk < a.N ? ibeg + chunk : a.N;
  if (ibeg >= iend) {                            // more blocks than cells: the partials of this block are zero
    if (MODE != EPI_INIT) {
      if (threadIdx.x == 0) a.elbo_part[blockIdx.x] = 0.0;
      if (MODE == EPI_TRAIN && threadIdx.x < C) a.gsum_part[(int64_t)blockIdx.x * C + threadIdx.x] = 0.0;
    }
    return;
  }
  // per-cell inputs, requested one group of cells ahead
  float psi_nx, s_nx, bm_nx, va_nx, t_nx = 0.f;
  auto request = [&](int64_t base) {
    int64_t n = base + sub;
    n = n < iend ? n : iend - 1;
    psi_nx = a.U[n];
    s_nx = a.s[n];
    bm_nx = a.Bm[n * C + c];
    va_nx = a.vA[n * C + c];
    if (MODE != EPI_INIT) t_nx = a.t[n * C + c];
  };
  const int64_t first = ibeg + (int64_t)wid * CPW, stride = (int64_t)nwarps * CPW;
  // The coefficient tables of a.smem_panels (>= 1) panels are staged at a time; in the common case (2 - 4 active panels)
  // that is all of them and the loop below runs once.  Otherwise the block walks its cells once per subset of panels and
  // a cell is worked on in the round that holds its panel.
  const int cap = a.smem_panels;
  for (int p0 = 0; p0 < (npan > 0 ? npan : 1); p0 += cap) {
  if (p0 > 0) __syncthreads();                  // every warp is done with the previous subset
  {   // [panel][pair][S*C] -> [panel][pair][sample][lane of the cell]
    const int per = KP * SC;                    // pairs per panel that exist
    const int np = npan - p0 < cap ? npan - p0 : cap;
    for (int i = threadIdx.x; i < np * per; i += blockDim.x) {
      const int pan = i / per, r = i - pan * per;
      const int kp = r / SC, sc = r - kp * SC;                    // sc = s * C + c
      const int s = sc / C, cc = sc - s * C;
      sm2[(size_t)pan * kPanelStride + (kp * SB + s) * WC + cc] = a.coef2[((int64_t)(p0 + pan) * KP + kp) * a.Jn + sc];
    }
    __syncthreads();
  }
  if (first < iend) request(first);
  for (int64_t base = first; base < iend; base += stride) {
    const int64_t n_raw = base + sub;
    bool nok = n_raw < iend;
    const int64_t n = nok ? n_raw : iend - 1;
    const float psif = psi_nx, tv_in = t_nx;
    const double sn = (double)s_nx, bv = (double)bm_nx + (double)va_nx;
    if (base + stride < iend) request(base + stride);
    const float mf = fmaxf(psif * wmin, psif * wmax);
    const double x = (double)psif, m = (double)mf;
    // ---- panel and panel variable: tt = (x - lo) * 2 / width - 1 ----
    int panel;
    double tt, ih, wref;                         // Z' = (2 / width) dp/dt + w_ref p: the shift m = psi w_ref is linear on a side
    if (x < 0.0) {
      int pf = (int)((x - pmin) * ih_neg * 0.5);
      pf = pf < 0 ? 0 : (pf >= nf_neg ? nf_neg - 1 : pf);
      tt = (x - (pmin + pf * w_neg)) * ih_neg - 1.0;
      panel = pf; ih = ih_neg; wref = (double)wmin;
    } else {
      int pf = (int)(x * ih_pos * 0.5);
      pf = pf >= nf_pos ? nf_pos - 1 : pf;
      tt = w_pos > 0.0 ? (x - pf * w_pos) * ih_pos - 1.0 : 0.0;
      panel = nf_neg + pf; ih = ih_pos; wref = (double)wmax;
    }
    // a NaN psi (diverged fit) must yield NaN results, not an out-of-range table index
    panel = panel < 0 ? 0 : (panel >= npan ? (npan > 0 ? npan - 1 : 0) : panel);
    nok = nok && panel >= p0 && panel < p0 + cap;                             // this round's cells
    if (!__any_sync(CA_FULL, nok)) continue;
    const bool act = nok && cok;
    const double2* tb = sm2 + (nok ? panel - p0 : 0) * kPanelStride + c;
    // ---- Z (and Z') of this lane's clone for every sample: Horner on coefficient pairs ----
    float rz[SB];                                // 1 / Z_s
    double L = 0.0, u = 0.0;                     // sum_s log Z_s, sum_s Z'_s / Z_s
#pragma unroll
    for (int s0 = 0; s0 < SB; s0 += SBB) {
      if (s0 < S) {                              // warp-uniform
        double p[SBB], dp[SBB];
        cell2_horner<MODE == EPI_TRAIN, WC, SB, SBB>(tb, tt, s0, S, p, dp);
        double pr = 1.0;
#pragma unroll
        for (int i = 0; i < SBB; ++i) {
          if (s0 + i < S) {
            pr *= p[i];
            const float zf = (float)p[i];
            if (MODE == EPI_TRAIN) {
              const double q = fma(ih, dp[i], wref * p[i]);        // Z'_s
              rz[s0 + i] = rcp_approx(zf);
              u = fma((double)rz[s0 + i], q, u);
              if (a.Zx && act) a.Zx[n * J + SC + (s0 + i) * C + c] = (float)q;
            }
            if (a.Zx && act) a.Zx[n * J + (s0 + i) * C + c] = zf;
          }
        }
        L += log_pos_f64(pr);
      }
    }
    if (cl == 0 && nok) a.shift[n] = mf;
    if (MODE == EPI_INIT) {
      // gamma_init: t <- F - logsumexp_c F   (sum over s, R/inference-tflow.R:338-340); runs once per fit: fp64
      const double F = (double)S * bv - sn * (L + (double)S * m);
      const double mx = group_max<WC>(cok ? F : -1e300);
      const double z = group_sum<WC>(cok ? exp(F - mx) : 0.0);
      if (act) a.t[n * C + c] = (float)(F - (mx + log(z)));
      continue;
    }
    const double F = bv - sn * (L * invS + m);
    // ---- gamma = softmax(t) over the lanes of the cell: exponentials in fp32, normalisation in fp64 ----
    const float tv = cok ? tv_in : -3.0e38f;
    const float mx = group_max<WC>(tv);
    const double ex = cok ? (double)expf(tv - mx) : 0.0;
    const double zs = group_sum<WC>(ex);            // in [1, C]
    const float lg = tv - (mx + __logf((float)zs));      // zs in [1, C]: absolute error < 1e-6
    double rcp = (double)rcp_approx((float)zs);
    rcp = rcp * fma(-zs, rcp, 2.0);
    rcp = rcp * fma(-zs, rcp, 2.0);
    const double g = ex * rcp;
    const double H = F + (double)la - (double)lg;
    const double gh = (g == 0.0) ? 0.0 : g * H;     // tf$where(gamma == 0, 0, gamma * log gamma), :333
    const double sumGH = group_sum<WC>(gh);
    if (a.Fout && act) a.Fout[n * C + c] = (float)F;
    // the Y-linear term psi_n (YW)_n joins the ELBO in k_yv_dot and (YW)_n joins d psi_n in k_adam_all (variant DEFER)
    if (cl == 0 && nok) elbo_w += sumGH - 0.5 * x * x - 0.5 * kLog2Pi;
    if (MODE == EPI_TRAIN) {
      if (act) {
        gacc += g;
        const float gt = (g == 0.0) ? 0.f : (float)(g * (H - sumGH));
        if (a.apply_t) {
          const int64_t o = n * C + c;
          float tnew = tv_in, mm_ = a.m_t[o], vv_ = a.v_t[o];
          adam_update_tf1(tnew, mm_, vv_, gt, lr_t);
          a.t[o] = tnew; a.m_t[o] = mm_; a.v_t[o] = vv_;
        } else {
          a.gT[n * C + c] = gt;
        }
      }
      // R_scn = gamma_nc s_n / (S Z_scn) and d psi_n = (YW)_n - sum_sc R Z' - psi_n
      const float gs = (float)(g * sn * invS);
      float* rx = a.Rx + n * a.Jn + c;              // [N][Jn]: the psi-scaled copy of R is not needed (k_gene_fused2)
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (s < S && act) { *rx = gs * rz[s]; rx += C; }
      }
      const double gu = group_sum<WC>(cok ? (double)gs * u : 0.0);
      if (cl == 0 && nok) a.gU[n] = (float)(-gu - x);
    }
  }
  }   // rounds over subsets of panels
  if (MODE == EPI_INIT) return;
  // ---- per-block partials in a fixed order: cells of a warp, then warps ----
  {
    double e = elbo_w, gsum = gacc;
#pragma unroll
    for (int k = 1; k < CPW; ++k) {
      e += __shfl_sync(CA_FULL, elbo_w, k * WC);
      gsum += __shfl_sync(CA_FULL, gacc, (k * WC + cl) & 31);
    }
    if (lane == 0) blkE[wid] = e;
    if (MODE == EPI_TRAIN && lane < C) blkG[(size_t)wid * C + lane] = gsum;
  }
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

}  // namespace ca

namespace ca {

// =====================================================================================================================
// Gene kernel of the CELL2 set: the backward interpolants dMx[g][j] = H_j(w_g) evaluated in the same layout -- lane =
// (gene, clone), all S samples of the clone in the lane, Horner on staged coefficient pairs -- and consumed in place by
// the gene-gradient reductions (k_gene_fused, kernels_fused.cuh: 900 warp instructions per gene, 41 us at config 3).
// The column partials of the Y pass are NOT read here (k_colpart_add adds them after the join), so the kernel runs before /
// next to the Y pass instead of behind it; the last block reduces the sum-gamma partials into the allreduce buffer.
//   ar[g]         = d/d loc_g      : - sum_sc sigma_sg L_gc dM_scg
//   ar[G + g]     = d/d log_sd_g   : - sum_sc sigma_sg L_gc dM_scg sd_g eps_sg
//   ar[2 G + g]   = d/d w_g        : - sum_sc mu_sg L_gc dM'_scg          (+ (Y^T psi)_g: k_colpart_add)
// Reference graph nodes: R/inference-tflow.R:288-296 (reverse mode of the log-normaliser), 345-346.
// =====================================================================================================================
struct Gene2Args {
  int G, C, S, SC, J, Jn, smem_panels;
  const InterpPlan* plan;
  const double2* coef2;      // backward monomial pairs of the S*C columns [panel][kIP / 2][SC]
  const float *Vm, *mu, *sig, *eps, *lsd, *L;
  float *ar, *dM_out;        // dM_out: inspection copy [G][J] or nullptr
  const double* gsum_part;   // [n_parts][C]
  int64_t n_parts;
};
constexpr int kGene2Warps = 16;
inline size_t gene2_smem_bytes(int WC, int SB, int smem_panels) { return (size_t)smem_panels * cell2_panel_bytes(WC, SB) + 16; }
inline int gene2_smem_panels(int WC, int SB, size_t budget) {
  const size_t n = budget / cell2_panel_bytes(WC, SB);
  return (int)(n > (size_t)kIMaxPanB ? (size_t)kIMaxPanB : n);
}

template <int WC, int SB>
__global__ void __launch_bounds__(kGene2Warps * 32, 2) k_gene_fused2(Gene2Args a) {
  CA_DYNAMIC_SMEM(double2, sm2);
  __shared__ double scratch[32];
  constexpr int GPW = 32 / WC;                  // genes per warp
  constexpr int SBB = SB < 4 ? SB : 4;
  constexpr int KP = kIP / 2;
  constexpr int kPanelStride = KP * SB * WC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / WC, cl = lane % WC;
  const int C = a.C, S = a.S, SC = a.SC, J = a.J, G = a.G;
  if (blockIdx.x == gridDim.x - 1) {   // role: sum_n gamma_nc partials -> float slots of the allreduce buffer
    float* out = a.ar + 3 * (int64_t)G;
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int64_t i = threadIdx.x; i < a.n_parts; i += blockDim.x) s += a.gsum_part[i * C + c];
      const double t = block_sum(s, scratch);
      if (threadIdx.x == 0) out[c] = (float)t;
    }
    return;
  }
  const bool cok = cl < C;
  const int c = cok ? cl : C - 1;
  const InterpPlan pl = *a.plan;
  const int npan = pl.nb;
  const double ih = pl.b_w > 0.0 ? 2.0 / pl.b_w : 0.0;
  const int nblk = gridDim.x - 1;
  const int cap = a.smem_panels;
  const int first = (blockIdx.x * nwarps + wid) * GPW, stride = nblk * nwarps * GPW;
  for (int p0 = 0; p0 < (npan > 0 ? npan : 1); p0 += cap) {
    if (p0 > 0) __syncthreads();
    {
      const int per = KP * SC;
      const int np = npan - p0 < cap ? npan - p0 : cap;
      for (int i = threadIdx.x; i < np * per; i += blockDim.x) {
        const int pan = i / per, r = i - pan * per;
        const int kp = r / SC, sc = r - kp * SC;
        const int s = sc / C, cc = sc - s * C;
        sm2[(size_t)pan * kPanelStride + (kp * SB + s) * WC + cc] = a.coef2[((int64_t)(p0 + pan) * KP + kp) * a.Jn + sc];
      }
      __syncthreads();
    }
    for (int gb = first; gb < G; gb += stride) {
      const int g_raw = gb + sub;
      bool gok = g_raw < G;
      const int g = gok ? g_raw : G - 1;
      const float wf = a.Vm[g];
      const float lsdf = a.lsd[g];
      const float lc = a.L[(int64_t)g * C + c];
      const double x = (double)wf;
      int pb = (int)((x - pl.wmin) * ih * 0.5);
      pb = pb < 0 ? 0 : (pb >= npan ? (npan > 0 ? npan - 1 : 0) : pb);
      const double tt = pl.b_w > 0.0 ? (x - (pl.wmin + pb * pl.b_w)) * ih - 1.0 : 0.0;
      gok = gok && pb >= p0 && pb < p0 + cap;
      if (!__any_sync(CA_FULL, gok)) continue;
      const bool act = gok && cok;
      const double2* tb = sm2 + (gok ? pb - p0 : 0) * kPanelStride + c;
      const double sd = (double)expf(lsdf);
      double aloc = 0.0, alsd = 0.0, gv = 0.0;     // per clone: sums over the samples, times L_gc below
#pragma unroll
      for (int s0 = 0; s0 < SB; s0 += SBB) {
        if (s0 < S) {
          float sg[SBB], ep[SBB], mu[SBB];        // requested before the recurrences, consumed behind them
#pragma unroll
          for (int i = 0; i < SBB; ++i) {
            const int64_t o = (int64_t)(s0 + i < S ? s0 + i : S - 1) * G + g;
            sg[i] = a.sig[o]; ep[i] = a.eps[o]; mu[i] = a.mu[o];
          }
          double d[SBB], dd[SBB];
          cell2_horner<true, WC, SB, SBB>(tb, tt, s0, S, d, dd);
#pragma unroll
          for (int i = 0; i < SBB; ++i) {
            if (s0 + i < S) {
              // dM' = sum_n psi_n R exp(psi_n w - m_n) = d(dM)/dw: the derivative of the interpolant (the shifts m_n do not depend on w)
              const float df = (float)d[i], d2f = (float)(ih * dd[i]);   // the unfused path rounds dMx to fp32: keep its numerics
              if (a.dM_out && act) {
                a.dM_out[(int64_t)g * J + (s0 + i) * C + c] = df;
                a.dM_out[(int64_t)g * J + SC + (s0 + i) * C + c] = d2f;
              }
              const double dx = (double)sg[i] * (double)df;
              aloc += dx;
              alsd = fma(dx, (double)ep[i], alsd);
              gv = fma((double)mu[i], (double)d2f, gv);
            }
          }
        }
      }
      const double lcd = cok ? -(double)lc : 0.0;
      aloc = group_sum<WC>(aloc * lcd);
      alsd = group_sum<WC>(alsd * lcd) * sd;
      gv = group_sum<WC>(gv * lcd);
      if (cl == 0 && gok) {
        a.ar[g] = (float)aloc;
        a.ar[G + g] = (float)alsd;
        a.ar[2 * (int64_t)G + g] = (float)gv;
      }
    }
  }
}

// (Y^T psi)_g = sum over the row blocks of the Y pass' column partials [nRB][G], in a fixed order, added to the w-gradient
// slot of the allreduce buffer once the Y pass has been joined: 4 slices of row blocks per gene, 8 loads in flight per thread.
constexpr int kColAddGenes = 64, kColAddSlices = 4;
__global__ void __launch_bounds__(kColAddGenes * kColAddSlices)
k_colpart_add(int G, int nRB, const float* __restrict__ colpart, float* __restrict__ ar_w /*ar + 2 G*/, float* __restrict__ YtU) {
  __shared__ double part[kColAddSlices][kColAddGenes];
  const int gi = threadIdx.x % kColAddGenes, sl = threadIdx.x / kColAddGenes;
  const int g = blockIdx.x * kColAddGenes + gi;
  double acc = 0.0;
  if (g < G) {
    const float* cp = colpart + g;
    int rb = sl;
    for (; rb + 7 * kColAddSlices < nRB; rb += 8 * kColAddSlices) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = cp[(int64_t)(rb + u * kColAddSlices) * G];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += (double)v[u];
    }
    for (; rb < nRB; rb += kColAddSlices) acc += (double)cp[(int64_t)rb * G];
  }
  part[sl][gi] = acc;
  __syncthreads();
  if (sl == 0 && g < G) {
    double t = part[0][gi];
#pragma unroll
    for (int z = 1; z < kColAddSlices; ++z) t += part[z][gi];
    YtU[g] = (float)t;
    ar_w[g] = (float)((double)ar_w[g] + t);
  }
}

}  // namespace ca
