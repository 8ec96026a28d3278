// "interp" path (K = 1, P = 0): both contractions collapse to UNIVARIATE functions.
//
// With a single latent dimension eta_ng = psi_n w_g, so
//     Zx[n][j]  = sum_g Mx[g][j] exp(psi_n w_g - m(psi_n)) = F_j(psi_n)     (normaliser, R/inference-tflow.R:288-290)
//     dMx[g][j] = sum_n Rx[n][j] exp(psi_n w_g - m_n)       = H_j(w_g)       (its reverse-mode gradient)
// where F_j, H_j are sums of exponentials: entire functions of ONE real variable.  Piecewise Chebyshev
// interpolation on panels whose exponent half-range is <= kAmax converges spectrally (24 nodes: ~1e-14 relative, 16 nodes: ~6e-9;
// scripts/interp_prototype.py, tests/test_interp_model.py), so the (N x G) x (G x J) contraction is replaced by
//     nodes:  (n_nodes x G) x (G x J)   with n_nodes ~ 50..200 instead of N = 100 000        (k_interp_nodes2<FWD>)
//     DCT  :  node values -> Chebyshev coefficients per panel                                  (k_interp_coeffs2)
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
constexpr int kIMaxPanF = 32;      // forward panels (both signs of psi together): exponent range 16 x 8 = 128 per side
constexpr int kIMaxPanB = 32;      // backward panels over [w_min, w_max]
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
// Chebyshev nodes cos(pi (p + 1/2) / 16) and the DCT matrix cos(pi k (q + 1/2) / 16) as correctly rounded literals (generated with
// Python's math.cos; kIP == 16): the launches that need them are latency-bound, a double-precision cos() per thread at their start
// is microseconds on the step's critical path
static_assert(kIP == 16, "regenerate kChebNodeTable / kDctTable for another node count");
static __device__ const double kChebNodeTable[kIP] = {0.9951847266721969, 0.9569403357322088, 0.881921264348355, 0.773010453362737, 0.6343932841636455, 0.4713967368259978, 0.29028467725446233, 0.09801714032956077, -0.09801714032956065, -0.29028467725446216, -0.4713967368259977, -0.6343932841636454, -0.773010453362737, -0.8819212643483549, -0.9569403357322088, -0.9951847266721968};
static __device__ const double kDctTable[kIP][kIP] = {
  {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0},
  {0.9951847266721969, 0.9569403357322088, 0.881921264348355, 0.773010453362737, 0.6343932841636455, 0.4713967368259978, 0.29028467725446233, 0.09801714032956077, -0.09801714032956065, -0.29028467725446216, -0.4713967368259977, -0.6343932841636454, -0.773010453362737, -0.8819212643483549, -0.9569403357322088, -0.9951847266721968},
  {0.9807852804032304, 0.8314696123025452, 0.5555702330196023, 0.19509032201612833, -0.1950903220161282, -0.555570233019602, -0.8314696123025453, -0.9807852804032304, -0.9807852804032304, -0.8314696123025455, -0.5555702330196022, -0.19509032201612866, 0.1950903220161283, 0.5555702330196018, 0.8314696123025452, 0.9807852804032303},
  {0.9569403357322088, 0.6343932841636455, 0.09801714032956077, -0.4713967368259977, -0.8819212643483549, -0.9951847266721969, -0.7730104533627371, -0.29028467725446244, 0.29028467725446205, 0.7730104533627367, 0.9951847266721969, 0.881921264348355, 0.471396736825998, -0.09801714032955997, -0.6343932841636448, -0.9569403357322085},
  {0.9238795325112867, 0.38268343236508984, -0.3826834323650897, -0.9238795325112867, -0.9238795325112868, -0.38268343236509034, 0.38268343236509, 0.9238795325112865, 0.9238795325112867, 0.38268343236509045, -0.3826834323650899, -0.9238795325112864, -0.9238795325112867, -0.38268343236509056, 0.3826834323650898, 0.9238795325112864},
  {0.881921264348355, 0.09801714032956077, -0.773010453362737, -0.9569403357322089, -0.29028467725446244, 0.6343932841636456, 0.9951847266721969, 0.471396736825998, -0.4713967368259975, -0.9951847266721969, -0.6343932841636454, 0.29028467725446266, 0.9569403357322085, 0.7730104533627377, -0.09801714032955972, -0.8819212643483547},
  {0.8314696123025452, -0.1950903220161282, -0.9807852804032304, -0.5555702330196022, 0.5555702330196018, 0.9807852804032304, 0.19509032201612878, -0.8314696123025451, -0.8314696123025456, 0.1950903220161272, 0.9807852804032304, 0.5555702330196025, -0.5555702330196015, -0.9807852804032307, -0.19509032201613, 0.831469612302544},
  {0.773010453362737, -0.4713967368259977, -0.9569403357322089, 0.09801714032956009, 0.9951847266721969, 0.29028467725446255, -0.8819212643483548, -0.6343932841636454, 0.6343932841636447, 0.8819212643483553, -0.29028467725446255, -0.9951847266721969, -0.0980171403295627, 0.9569403357322089, 0.4713967368259984, -0.7730104533627357},
  {0.7071067811865476, -0.7071067811865475, -0.7071067811865477, 0.7071067811865474, 0.7071067811865477, -0.7071067811865467, -0.7071067811865471, 0.7071067811865466, 0.7071067811865472, -0.7071067811865465, -0.7071067811865474, 0.7071067811865464, 0.7071067811865475, -0.7071067811865464, -0.7071067811865476, 0.7071067811865462},
  {0.6343932841636455, -0.8819212643483549, -0.29028467725446244, 0.9951847266721969, -0.09801714032955997, -0.9569403357322087, 0.47139673682599736, 0.7730104533627377, -0.773010453362737, -0.4713967368259983, 0.9569403357322089, 0.09801714032956282, -0.995184726672197, 0.2902846772544622, 0.8819212643483563, -0.6343932841636443},
  {0.5555702330196023, -0.9807852804032304, 0.1950903220161283, 0.8314696123025455, -0.8314696123025451, -0.19509032201612803, 0.9807852804032307, -0.5555702330196015, -0.5555702330196026, 0.9807852804032304, -0.19509032201612858, -0.8314696123025449, 0.8314696123025438, 0.19509032201613036, -0.9807852804032308, 0.5555702330196011},
  {0.4713967368259978, -0.9951847266721969, 0.6343932841636449, 0.2902846772544634, -0.9569403357322093, 0.7730104533627359, 0.09801714032956259, -0.8819212643483562, 0.8819212643483538, -0.0980171403295577, -0.773010453362739, 0.9569403357322078, -0.2902846772544587, -0.6343932841636487, 0.9951847266721965, -0.4713967368259935},
  {0.38268343236508984, -0.9238795325112868, 0.9238795325112865, -0.3826834323650899, -0.38268343236509056, 0.9238795325112867, -0.9238795325112864, 0.38268343236508956, 0.3826834323650909, -0.9238795325112876, 0.9238795325112868, -0.3826834323650892, -0.3826834323650912, 0.9238795325112877, -0.9238795325112854, 0.38268343236508556},
  {0.29028467725446233, -0.7730104533627369, 0.9951847266721968, -0.8819212643483556, 0.47139673682599736, 0.09801714032955905, -0.6343932841636456, 0.9569403357322084, -0.9569403357322089, 0.634393284163647, -0.09801714032956099, -0.47139673682599564, 0.8819212643483547, -0.9951847266721968, 0.7730104533627398, -0.29028467725446505},
  {0.19509032201612833, -0.5555702330196022, 0.8314696123025455, -0.9807852804032307, 0.9807852804032304, -0.831469612302545, 0.5555702330196015, -0.19509032201612858, -0.19509032201613025, 0.5555702330196028, -0.831469612302545, 0.9807852804032309, -0.9807852804032297, 0.8314696123025456, -0.5555702330196007, 0.19509032201612425},
  {0.09801714032956077, -0.2902846772544633, 0.471396736825998, -0.6343932841636467, 0.7730104533627377, -0.8819212643483562, 0.9569403357322094, -0.995184726672197, 0.9951847266721965, -0.9569403357322078, 0.8819212643483536, -0.7730104533627331, 0.6343932841636439, -0.47139673682599326, 0.2902846772544547, -0.09801714032955673}};
__device__ __forceinline__ double cheb_node(int p) { return kChebNodeTable[p]; }

// Chebyshev -> monomial basis of the panel variable t: T_k(t) = sum_m kT2M.v[k][m] t^m (integers up to 2^14 * 3.3, exact in
// fp64; T_{k+1} = 2 t T_k - T_{k-1}).  The per-cell and per-gene kernels evaluate the interpolants by Horner's rule (one
// DFMA per coefficient instead of Clenshaw's DADD + DFMA).  Conditioning: the panels keep the exponent half-range <= kIAmax,
// the interpolated sums of exponentials have monomial coefficients bounded by sum_k A^k / k! = e^A times their maximum,
// against values >= e^-2A of it: evaluation error <= e^(3A) 1.1e-16 ~ 2e-11 relative at the worst point of a worst-case
// panel, below the 6e-9 of the interpolation itself (tests/test_interp_model.py).
struct ChebToMono { double v[kIP][kIP]; };
constexpr ChebToMono make_cheb_to_mono() {
  ChebToMono t{};
  t.v[0][0] = 1.0;
  t.v[1][1] = 1.0;
  for (int k = 2; k < kIP; ++k)
    for (int m = 0; m <= k; ++m) t.v[k][m] = (m > 0 ? 2.0 * t.v[k - 1][m - 1] : 0.0) - t.v[k - 2][m];
  return t;
}
static __device__ const ChebToMono kT2M = make_cheb_to_mono();   // global memory: indexed per thread (staged to shared memory by its users)

// panel structure from the current ranges of psi and w
// `wide` (CELL2 set): the derivative columns Z' = dZ/dpsi + w_ref Z and dM' = d(dM)/dw are taken from the DERIVATIVE of the
// interpolants instead of being interpolated themselves (half the node sums, half the tables).  Differentiating amplifies the
// fp32-class noise of the node values (~2e-8 relative) by ~n^2 / (panel half-width), so a single panel that would be narrower
// than an exponent half-range of kIAmin = 2 is widened to it (it then reaches beyond the data: harmless, the function is entire):
// |d/dx| error <= ~1e-6 of (scale x value) for every range of psi and W, including W = 0 and psi = 0 at the start of a fit.
constexpr double kIAmin = 2.0;
__device__ __forceinline__ InterpPlan interp_make_plan(double wmin, double wmax, double pmin, double pmax, bool wide = false) {
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
  if (wide) {
    const double wf = fmin(2.0 * kIAmin / fmax(D, 1e-30), 1e3), wb = fmin(2.0 * kIAmin / fmax(A, 1e-30), 1e3);
    if (pl.nf_neg == 1 && pl.f_neg_w < wf) { pl.f_neg_w = wf; pl.pmin = -wf; }     // panel [-wf, 0)
    if (pl.nf_pos == 1 && pl.f_pos_w < wf) pl.f_pos_w = wf;                         // panel [0, wf]
    if (pl.nb == 1 && pl.b_w < wb) pl.b_w = wb;                                     // panel [wmin, wmin + wb]
  }
  return pl;
}
// one thread
__global__ void k_interp_plan(const float* __restrict__ mm_w, const float* __restrict__ mm_psi, InterpPlan* __restrict__ plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *plan = interp_make_plan((double)mm_w[0], (double)mm_w[1], (double)mm_psi[0], (double)mm_psi[1]);
}

// =====================================================================================================================
// Node sums (k_interp_nodes2 / k_interp_coeffs2).  The round-1 kernel kept 48 fp64 accumulators per lane, fed by one SHFL
// + one F2F per 6 DFMA, 8 warps per SM, every group of 8 nodes re-reading its slice of B: 0.24 ms for 3e8 MACs at config 3,
// issue slots 11 % busy (ncu of round 2, profiles/r02a_ncu_full_c3_start_of_round.md).  The node sums are a plain dense
// contraction  V[node][j] = sum_r E[r][node] B[r][j]  with a generated A operand (16 nodes per panel, J <= 256 columns,
// reduction over r = genes (FWD) or cells (BWD)), so they are tiled like one:
//   * work item = (panel, slice of the reduction index); 128-thread blocks stride over the items (the panel count is
//     device-side data), 3 blocks per SM;
//   * B rows arrive in 32-row chunks by cp.async (16 bytes per request, double buffered: the next chunk is in flight
//     while this one is consumed; a chunk of rows is ONE contiguous range of B);
//   * the exponentials of a chunk (32 rows x 16 nodes: 4 per thread, fp64 argument, fp32 expf -- as before) go to shared
//     memory as duplicated pairs (e, e), so that a thread's 4 nodes x TJ columns are fed by 2 + TJ/4 vector loads and
//     updated by 2 TJ packed fma.rn.f32x2 per row: 16 FFMA2 per 4 LDS (TJ = 8);
//   * fp32 accumulation runs over ONE chunk only (<= 32 terms), then is flushed into fp64 registers; slices are summed
//     in fp64 in a fixed order by k_interp_coeffs2: same accuracy class as the fp64 kernel (products are exact in fp64
//     there, rounded to fp32 here: 6e-8 relative per term, averaged over the sum), deterministic.
// Columns beyond J (padding of the last column group) read the neighbouring row / slack and are never stored.
// =====================================================================================================================
constexpr int kN2Threads = 128;
constexpr int kN2Chunk = 32;                 // rows of the reduction index per staged chunk
constexpr int kN2EPitch = 2 * kIP + 8;       // floats per row of the (e, e) tile: 16 pairs + 8 floats of bank shift
constexpr int kN2BlocksPerSM = 4;                // 128 registers: two blocks fit next to the two persistent Y-pass CTAs of an SM

inline int n2_pick_tj(int J) {               // columns per thread: the choice that pads fewer columns
  auto padded = [](int J_, int tj) { int ncg = (J_ + tj - 1) / tj, p = 1; while (p < ncg) p <<= 1; return p * tj; };
  if (J > 6 * 32) return 8;                  // 6 x 32 lanes cover at most 192 columns (J <= 256 = 8 x 32 always fits)
  return padded(J, 6) <= padded(J, 8) ? 6 : 8;
}
inline int n2_ncg_pow2(int J, int tj) { int ncg = (J + tj - 1) / tj, p = 1; while (p < ncg) p <<= 1; return p; }
inline size_t n2_smem_bytes(int J, int tj) {
  const size_t tile = (size_t)kN2Chunk * J + (size_t)32 * tj;          // + slack for the padded columns of the last row
  return (2 * tile + 2 * (size_t)kN2Chunk * kN2EPitch) * sizeof(float) + kIP * sizeof(double) + 16;
}

template <bool FWD, int TJ>
__global__ void __launch_bounds__(kN2Threads, kN2BlocksPerSM)
k_interp_nodes2(const InterpPlan* __restrict__ plan, const float* __restrict__ rv /*FWD: w[G]  BWD: psi[N]*/,
                const float* __restrict__ shift /*BWD: m[N]*/, const float* __restrict__ B /*[R][J]*/, int64_t R, int J,
                int ncgp /*column groups, power of two <= 32*/, int nsplit, int max_pan, double* __restrict__ vals) {
  CA_DYNAMIC_SMEM(float, smf);
  const InterpPlan pl = *plan;
  const int npan = FWD ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const int tid = threadIdx.x, lane = tid & 31, qg = tid >> 5;          // qg: nodes 4 qg .. 4 qg + 3 of the panel
  const int cg = lane & (ncgp - 1), ks = lane / ncgp, nks = 32 / ncgp;   // column group / row slice of this lane
  const size_t tile = (size_t)kN2Chunk * J + (size_t)32 * TJ;
  float* Bs[2] = {smf, smf + tile};
  float* Es[2] = {smf + 2 * tile, smf + 2 * tile + (size_t)kN2Chunk * kN2EPitch};
  double* xq = reinterpret_cast<double*>(smf + 2 * tile + 2 * (size_t)kN2Chunk * kN2EPitch + 2);   // 8-byte aligned: see n2_smem_bytes
  xq = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(xq) + 7) & ~(uintptr_t)7);
  const int64_t per = ((R + nsplit - 1) / nsplit + 3) & ~(int64_t)3;     // rows per slice, multiple of 4 (16-byte chunk starts)
  const bool vec16 = (J % 4) == 0;
  const int64_t nodes_total = (int64_t)max_pan * kIP;
  const int e_row = tid >> 2, e_q = (tid & 3) * 4;                        // exponentials: row of the chunk, first of 4 nodes

  for (int item = blockIdx.x; item < npan * nsplit; item += gridDim.x) {
    const int panel = item % npan, split = item / npan;
    const int64_t rbeg = (int64_t)split * per;
    const int64_t rend = rbeg + per < R ? rbeg + per : R;
    double mid, half, wref = 0.0;
    if (FWD) fwd_panel(pl, panel, mid, half, wref);
    else bwd_panel(pl, panel, mid, half);
    __syncthreads();                                   // the previous item's readers of xq / tiles are done
    if (tid < kIP) xq[tid] = mid + half * cheb_node(tid);
    double acc64[4][TJ];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < TJ; ++j) acc64[q][j] = 0.0;
    const int nchunks = rend > rbeg ? (int)((rend - rbeg + kN2Chunk - 1) / kN2Chunk) : 0;
    auto issue = [&](int k) {                           // chunk k -> Bs[k & 1]: one contiguous range of B
      const int64_t r0 = rbeg + (int64_t)k * kN2Chunk;
      const int rows = (int)((rend - r0) < kN2Chunk ? (rend - r0) : kN2Chunk);
      const float* src = B + r0 * J;
      float* dst = Bs[k & 1];
      const int nfl = rows * J;
      if (vec16) { for (int i = tid * 4; i < nfl; i += kN2Threads * 4) cp_async16(dst + i, src + i); }
      else { for (int i = tid * 2; i < nfl; i += kN2Threads * 2) cp_async8(dst + i, src + i); }
      cp_async_commit();
    };
    // row scalars of the chunk this thread generates exponentials for (prefetched one chunk ahead)
    auto load_rv = [&](int k, float& v, float& sh) {
      const int64_t r = rbeg + (int64_t)k * kN2Chunk + e_row;
      v = r < rend ? rv[r] : 0.f;
      sh = (!FWD && r < rend) ? shift[r] : 0.f;
    };
    float v_cur = 0.f, sh_cur = 0.f, v_nxt = 0.f, sh_nxt = 0.f;
    if (nchunks > 0) { issue(0); load_rv(0, v_cur, sh_cur); }
    __syncthreads();                                   // xq visible
    for (int k = 0; k < nchunks; ++k) {
      if (k + 1 < nchunks) load_rv(k + 1, v_nxt, sh_nxt);
      {   // exponentials of chunk k: row e_row, nodes e_q .. e_q + 3, stored as (e, e) pairs
        float* er = Es[k & 1] + (size_t)e_row * kN2EPitch + 2 * e_q;
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double x = xq[e_q + q];
          e[q] = FWD ? expf((float)(x * ((double)v_cur - wref))) : expf((float)((double)v_cur * x - (double)sh_cur));
        }
        reinterpret_cast<float4*>(er)[0] = make_float4(e[0], e[0], e[1], e[1]);
        reinterpret_cast<float4*>(er)[1] = make_float4(e[2], e[2], e[3], e[3]);
      }
      cp_async_wait<0>();
      __syncthreads();                                 // chunk k (every thread's copies) and its exponentials are visible;
                                                       // everyone has finished consuming chunk k - 1
      if (k + 1 < nchunks) issue(k + 1);               // overwrites the buffer chunk k - 1 was read from
      const int64_t r0 = rbeg + (int64_t)k * kN2Chunk;
      const int rows = (int)((rend - r0) < kN2Chunk ? (rend - r0) : kN2Chunk);
      float2 acc[4][TJ / 2];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < TJ / 2; ++j) acc[q][j] = make_float2(0.f, 0.f);
      const float* bt = Bs[k & 1] + cg * TJ;
      const float* et = Es[k & 1] + 8 * qg;
#pragma unroll 4
      for (int c = ks; c < rows; c += nks) {
        float2 b[TJ / 2];
        if (TJ == 8 && vec16) {                 // rows of B are 16-byte aligned: two 16-byte loads
          const float4 b0 = *reinterpret_cast<const float4*>(bt + (size_t)c * J);
          const float4 b1 = *reinterpret_cast<const float4*>(bt + (size_t)c * J + 4);
          b[0] = make_float2(b0.x, b0.y); b[1] = make_float2(b0.z, b0.w);
          b[TJ / 2 - 2] = make_float2(b1.x, b1.y); b[TJ / 2 - 1] = make_float2(b1.z, b1.w);
        } else {                                // J is even: 8-byte alignment always holds
#pragma unroll
          for (int j = 0; j < TJ / 2; ++j) b[j] = *reinterpret_cast<const float2*>(bt + (size_t)c * J + 2 * j);
        }
        const float4 e01 = *reinterpret_cast<const float4*>(et + (size_t)c * kN2EPitch);
        const float4 e23 = *reinterpret_cast<const float4*>(et + (size_t)c * kN2EPitch + 4);
        const float2 ee[4] = {make_float2(e01.x, e01.y), make_float2(e01.z, e01.w), make_float2(e23.x, e23.y), make_float2(e23.z, e23.w)};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < TJ / 2; ++j) acc[q][j] = __ffma2_rn(ee[q], b[j], acc[q][j]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < TJ / 2; ++j) {
          acc64[q][2 * j] += (double)acc[q][j].x;
          acc64[q][2 * j + 1] += (double)acc[q][j].y;
        }
      v_cur = v_nxt; sh_cur = sh_nxt;
    }
    // row slices of a warp (ks): combined in slice order through shuffles (fixed order), then lanes with ks == 0 store
    if (nks > 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          double t = acc64[q][j];
          for (int s2 = 1; s2 < nks; ++s2) {
            const double o = __shfl_sync(CA_FULL, acc64[q][j], cg + s2 * ncgp);
            t += o;
          }
          acc64[q][j] = t;
        }
    }
    if (ks == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double* out = vals + ((int64_t)split * nodes_total + panel * kIP + 4 * qg + q) * J;
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          const int col = cg * TJ + j;
          if (col < J) out[col] = acc64[q][j];
        }
      }
    }
  }
}

// Coefficients from the slice partials of k_interp_nodes2: block = (4 columns, panel); thread (slice lane z, node p, column)
// sums every 8th slice -- 8 independent loads in flight per thread: the slices of one (node, column) are megabytes apart,
// so a serial chain of loads costs one L2 / DRAM latency per slice (the first version: 47 us for 296 slices at config 3,
// issue slots 6 % busy) -- the 8 slice lanes are combined in a fixed order through shared memory, then thread (k = p, column)
// applies the DCT  c_k = (2/P) sum_p f(x_p) cos(pi k (p + 1/2) / P),  c_0 halved.  Blocks of inactive panels exit at once.
constexpr int kC2Cols = 4, kC2Lanes = 8, kC2PanelsY = 8;
__global__ void __launch_bounds__(kIP * kC2Cols * kC2Lanes)
k_interp_coeffs2(const InterpPlan* __restrict__ plan, const double* __restrict__ vals, int nsplit, int max_pan, int J,
                 int fwd, double* __restrict__ coeff, double* __restrict__ coef2 /*monomial pairs [panel][kIP/2][J][2] or nullptr*/) {
  __shared__ double part[kC2Lanes][kIP][kC2Cols];
  __shared__ double cheb[kIP][kC2Cols];
  __shared__ double t2m[kIP][kIP];
  __shared__ double ct[kIP][kIP];
  const InterpPlan pl = *plan;
  const int npan = fwd ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const int c4 = threadIdx.x % kC2Cols, p = (threadIdx.x / kC2Cols) % kIP, z = threadIdx.x / (kC2Cols * kIP);
  const int j = blockIdx.x * kC2Cols + c4;
  const int64_t nodes_total = (int64_t)max_pan * kIP;
  if (threadIdx.x < kIP * kIP) {
    const int k = threadIdx.x / kIP, q = threadIdx.x % kIP;
    ct[k][q] = kDctTable[k][q];
    t2m[k][q] = kT2M.v[k][q];
  }
  for (int panel = blockIdx.y; panel < npan; panel += gridDim.y) {     // grid.y = kC2PanelsY: blocks stride over the active panels
  double acc = 0.0;
  if (j < J) {
    const double* base = vals + ((int64_t)panel * kIP + p) * J + j;
    const int64_t stride = nodes_total * J;                     // doubles between consecutive slices
    int s = z;
    for (; s + 7 * kC2Lanes < nsplit; s += 8 * kC2Lanes) {     // 8 loads in flight, summed in slice order
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = base[(int64_t)(s + u * kC2Lanes) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; s < nsplit; s += kC2Lanes) acc += base[(int64_t)s * stride];
  }
  part[z][p][c4] = acc;
  __syncthreads();
  if (z == 0) {
    double f = part[0][p][c4];
#pragma unroll
    for (int zz = 1; zz < kC2Lanes; ++zz) f += part[zz][p][c4];
    part[0][p][c4] = f;
  }
  __syncthreads();
  if (z == 0) {
    const int k = p;
    double c = 0.0;
#pragma unroll
    for (int q = 0; q < kIP; ++q) c += part[0][q][c4] * ct[k][q];
    c *= 2.0 / kIP;
    if (k == 0) c *= 0.5;
    if (j < J) coeff[((int64_t)panel * kIP + k) * J + j] = c;
    cheb[k][c4] = c;
  }
  __syncthreads();   // part is rewritten for the next panel
  if (coef2 && z == 0 && j < J) {   // monomial coefficient a_m, m = p: smallest Chebyshev terms first
    const int mo = p;
    double am = 0.0;
#pragma unroll
    for (int k = kIP - 1; k >= 0; --k) am += t2m[k][mo] * cheb[k][c4];   // entries with k < m are zero
    coef2[(((int64_t)panel * (kIP / 2) + (mo >> 1)) * J + j) * 2 + (mo & 1)] = am;
  }
  }
}

// k_interp_coeffs3: the same result as k_interp_coeffs2 with every block busy.  ncu of round 2 (config 3, backward pass):
// 22 us for 10.9 MB of slice partials -- with one or two active panels only 48 - 96 of the 384 blocks had work, and each of
// their threads walked 55 slices that lie 786 KB apart.  Here the slices are split over grid.y = kC3Groups blocks per
// column group as well: a thread has at most 8 loads, all in flight at once; the block combines its slice lanes through
// shared memory and publishes one partial per (panel, node, column); the block of a column group that arrives LAST (ticket)
// adds the kC3Groups partials in group order and applies the DCT and the monomial conversion.  Fixed summation order:
// slices s = 64 i + 8 z + gy -> over i per thread, over z per block, over gy by the last block.
constexpr int kC3Groups = 8;
__global__ void __launch_bounds__(kIP * kC2Cols * kC2Lanes)
k_interp_coeffs3(const InterpPlan* __restrict__ plan, const double* __restrict__ vals, int nsplit, int max_pan, int J, int fwd,
                 double* __restrict__ partial2 /*[kC3Groups][max_pan][kIP][J]*/, unsigned* __restrict__ tickets /*[grid.x], zero*/,
                 double* __restrict__ coeff, double* __restrict__ coef2) {
  __shared__ double part[kC2Lanes][kIP][kC2Cols];
  __shared__ double cheb[kIP][kC2Cols];
  __shared__ double t2m[kIP][kIP];
  __shared__ double ct[kIP][kIP];
  __shared__ int is_last;
  const InterpPlan pl = *plan;
  const int npan = fwd ? (pl.nf_neg + pl.nf_pos) : pl.nb;
  const int c4 = threadIdx.x % kC2Cols, p = (threadIdx.x / kC2Cols) % kIP, z = threadIdx.x / (kC2Cols * kIP);
  const int j = blockIdx.x * kC2Cols + c4;
  const int gy = blockIdx.y;
  const int64_t nodes_total = (int64_t)max_pan * kIP;
  if (threadIdx.x < kIP * kIP) {
    const int k = threadIdx.x / kIP, q = threadIdx.x % kIP;
    ct[k][q] = kDctTable[k][q];
    t2m[k][q] = kT2M.v[k][q];
  }
  constexpr int kStep = kC3Groups * kC2Lanes;                       // slices between two loads of a thread
  for (int panel = 0; panel < npan; ++panel) {
    double acc = 0.0;
    if (j < J) {
      const double* base = vals + ((int64_t)panel * kIP + p) * J + j;
      const int64_t stride = nodes_total * J;                       // doubles between consecutive slices
      int s = z * kC3Groups + gy;
      for (; s + 7 * kStep < nsplit; s += 8 * kStep) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = base[(int64_t)(s + u * kStep) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (s + u * kStep < nsplit) ? base[(int64_t)(s + u * kStep) * stride] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
    }
    part[z][p][c4] = acc;
    __syncthreads();
    if (z == 0 && j < J) {
      double f = part[0][p][c4];
#pragma unroll
      for (int zz = 1; zz < kC2Lanes; ++zz) f += part[zz][p][c4];
      partial2[(((int64_t)gy * max_pan + panel) * kIP + p) * J + j] = f;
    }
    __syncthreads();   // part is rewritten for the next panel
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;                   // ready for the next launch (stream order)
  for (int panel = 0; panel < npan; ++panel) {
    if (z == 0) {
      double f = 0.0;
      if (j < J) {
        double v[kC3Groups];
#pragma unroll
        for (int g = 0; g < kC3Groups; ++g) v[g] = __ldcv(partial2 + (((int64_t)g * max_pan + panel) * kIP + p) * J + j);
#pragma unroll
        for (int g = 0; g < kC3Groups; ++g) f += v[g];
      }
      part[0][p][c4] = f;
    }
    __syncthreads();
    if (z == 0) {
      const int k = p;
      double c = 0.0;
#pragma unroll
      for (int q = 0; q < kIP; ++q) c += part[0][q][c4] * ct[k][q];
      c *= 2.0 / kIP;
      if (k == 0) c *= 0.5;
      if (j < J) coeff[((int64_t)panel * kIP + k) * J + j] = c;
      cheb[k][c4] = c;
    }
    __syncthreads();
    if (coef2 && z == 0 && j < J) {   // monomial coefficient a_m, m = p: smallest Chebyshev terms first
      const int mo = p;
      double am = 0.0;
#pragma unroll
      for (int k = kIP - 1; k >= 0; --k) am += t2m[k][mo] * cheb[k][c4];   // entries with k < m are zero
      coef2[(((int64_t)panel * (kIP / 2) + (mo >> 1)) * J + j) * 2 + (mo & 1)] = am;
    }
    __syncthreads();
  }
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
