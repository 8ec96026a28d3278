// Runs the kernels of clonealign_b200/csrc/kernels_interp.cuh on the CPU emulation (tests/cuda_emul/cuda_emul.h) in the
// same order and with the same arguments as run_forward / run_train in core.cu, on inputs read from a binary file.
//   usage: run_interp in.bin out.bin
//   in : int32 N, G, J, smem_panels; float psi[N], w[G], Mx[G*J], Rx[N*J]
//   out: int32 nf_neg, nf_pos, nb; float Zx[N*J], dMx[G*J], Zh[N*J] (forward monomial table of k_interp_coeffs2, Horner on the host),
//        dMh[G*J] (backward monomial table of k_interp_coeffs3)
#include <cstdio>
#include <vector>

#include "../../clonealign_b200/csrc/common.cuh"
#include "../../clonealign_b200/csrc/kernels_interp.cuh"

using namespace ca;

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  int hdr[4];
  if (!f || fread(hdr, 4, 4, f) != 4) return 2;
  const int N = hdr[0], G = hdr[1], J = hdr[2], smem_panels = hdr[3];
  std::vector<float> psi(N), w(G), Mx((size_t)G * J), Rx((size_t)N * J), shift(N);
  if (fread(psi.data(), 4, N, f) != (size_t)N || fread(w.data(), 4, G, f) != (size_t)G ||
      fread(Mx.data(), 4, (size_t)G * J, f) != (size_t)G * J || fread(Rx.data(), 4, (size_t)N * J, f) != (size_t)N * J)
    return 3;
  fclose(f);
  float mm_w[2] = {w[0], w[0]}, mm_psi[2] = {psi[0], psi[0]};
  for (float v : w) { mm_w[0] = fminf(mm_w[0], v); mm_w[1] = fmaxf(mm_w[1], v); }
  for (float v : psi) { mm_psi[0] = fminf(mm_psi[0], v); mm_psi[1] = fmaxf(mm_psi[1], v); }
  for (int n = 0; n < N; ++n) shift[n] = fmaxf(psi[n] * mm_w[0], psi[n] * mm_w[1]);   // k_shift_k1

  InterpPlan plan;
  ca_emul::launch(k_interp_plan, dim3(1), dim3(32), 0, (const float*)mm_w, (const float*)mm_psi, &plan);
  const int npf = plan.nf_neg + plan.nf_pos;
  const int tj = n2_pick_tj(J), ncgp = n2_ncg_pow2(J, tj);
  const int split_f = 5, split_b = 7;                       // odd slice counts, ragged slices
  std::vector<double> vals((size_t)std::max(split_f * kIMaxPanF, split_b * kIMaxPanB) * kIP * J, 0.0);
  std::vector<double> coef((size_t)std::max(kIMaxPanF, kIMaxPanB) * kIP * J, 0.0);
  std::vector<double> coef2f(coef.size(), 0.0), coef2b(coef.size(), 0.0), part2((size_t)kC3Groups * std::max(kIMaxPanF, kIMaxPanB) * kIP * J, 0.0);
  std::vector<unsigned> tickets((J + kC2Cols - 1) / kC2Cols, 0u);
  std::vector<float> Zx((size_t)N * J, -1.f), dMx((size_t)G * J, -1.f), Zh((size_t)N * J, -1.f), dMh((size_t)G * J, -1.f);
  const size_t eval_smem = (size_t)smem_panels * kIP * J * sizeof(double);
  const size_t n2_smem = n2_smem_bytes(J, tj);

  // forward: nodes -> coefficients -> evaluation per cell (3 blocks stride over the (panel, slice) work items)
  if (tj == 8)
    ca_emul::launch(k_interp_nodes2<true, 8>, dim3(3), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)w.data(),
                    (const float*)nullptr, (const float*)Mx.data(), (int64_t)G, J, ncgp, split_f, kIMaxPanF, vals.data());
  else
    ca_emul::launch(k_interp_nodes2<true, 6>, dim3(3), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)w.data(),
                    (const float*)nullptr, (const float*)Mx.data(), (int64_t)G, J, ncgp, split_f, kIMaxPanF, vals.data());
  ca_emul::launch(k_interp_coeffs2, dim3((J + kC2Cols - 1) / kC2Cols, 3), dim3(kIP * kC2Cols * kC2Lanes), 0,
                  (const InterpPlan*)&plan, (const double*)vals.data(), split_f, kIMaxPanF, J, 1, coef.data(), coef2f.data());
  ca_emul::launch(k_interp_eval<true>, dim3(3), dim3(kIEvalWarps * 32), eval_smem, (const InterpPlan*)&plan,
                  (const double*)coef.data(), (const float*)psi.data(), (int64_t)N, J, Zx.data(), smem_panels);

  // backward: nodes over w (reduction over cells) -> coefficients -> evaluation per gene
  if (tj == 8)
    ca_emul::launch(k_interp_nodes2<false, 8>, dim3(4), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)psi.data(),
                    (const float*)shift.data(), (const float*)Rx.data(), (int64_t)N, J, ncgp, split_b, kIMaxPanB, vals.data());
  else
    ca_emul::launch(k_interp_nodes2<false, 6>, dim3(4), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)psi.data(),
                    (const float*)shift.data(), (const float*)Rx.data(), (int64_t)N, J, ncgp, split_b, kIMaxPanB, vals.data());
  ca_emul::launch(k_interp_coeffs3, dim3((J + kC2Cols - 1) / kC2Cols, kC3Groups), dim3(kIP * kC2Cols * kC2Lanes), 0,
                  (const InterpPlan*)&plan, (const double*)vals.data(), split_b, kIMaxPanB, J, 0, part2.data(), tickets.data(), coef.data(),
                  coef2b.data());
  for (unsigned t : tickets) if (t != 0u) return 4;         // the last block of every column group re-arms its ticket
  ca_emul::launch(k_interp_eval<false>, dim3(2), dim3(kIEvalWarps * 32), eval_smem, (const InterpPlan*)&plan,
                  (const double*)coef.data(), (const float*)w.data(), (int64_t)G, J, dMx.data(), smem_panels);

  // Horner evaluation of the monomial pair tables [panel][kIP / 2][J][2] on the host (what k_cell_fused2 / k_gene_fused2 do)
  auto horner = [&](const std::vector<double>& c2, int panel, int j, double t) {
    double p = 0.0;
    for (int m = kIP - 1; m >= 0; --m) p = p * t + c2[(((size_t)panel * (kIP / 2) + (m >> 1)) * J + j) * 2 + (m & 1)];
    return p;
  };
  for (int n = 0; n < N; ++n) {
    const double x = psi[n];
    int panel; double t;
    if (x < 0.0) {
      int pf = (int)((x - plan.pmin) / plan.f_neg_w);
      pf = pf < 0 ? 0 : (pf >= plan.nf_neg ? plan.nf_neg - 1 : pf);
      t = (x - (plan.pmin + pf * plan.f_neg_w)) * 2.0 / plan.f_neg_w - 1.0;
      panel = pf;
    } else {
      int pf = plan.f_pos_w > 0.0 ? (int)(x / plan.f_pos_w) : 0;
      pf = pf >= plan.nf_pos ? plan.nf_pos - 1 : pf;
      t = plan.f_pos_w > 0.0 ? (x - pf * plan.f_pos_w) * 2.0 / plan.f_pos_w - 1.0 : 0.0;
      panel = plan.nf_neg + pf;
    }
    for (int j = 0; j < J; ++j) Zh[(size_t)n * J + j] = (float)horner(coef2f, panel, j, t);
  }
  for (int g = 0; g < G; ++g) {
    const double x = w[g];
    int pb = plan.b_w > 0.0 ? (int)((x - plan.wmin) / plan.b_w) : 0;
    pb = pb < 0 ? 0 : (pb >= plan.nb ? plan.nb - 1 : pb);
    const double t = plan.b_w > 0.0 ? (x - (plan.wmin + pb * plan.b_w)) * 2.0 / plan.b_w - 1.0 : 0.0;
    for (int j = 0; j < J; ++j) dMh[(size_t)g * J + j] = (float)horner(coef2b, pb, j, t);
  }

  FILE* o = fopen(argv[2], "wb");
  int oh[3] = {plan.nf_neg, plan.nf_pos, plan.nb};
  fwrite(oh, 4, 3, o);
  fwrite(Zx.data(), 4, Zx.size(), o);
  fwrite(dMx.data(), 4, dMx.size(), o);
  fwrite(Zh.data(), 4, Zh.size(), o);
  fwrite(dMh.data(), 4, dMh.size(), o);
  fclose(o);
  return 0;
}
