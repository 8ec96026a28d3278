// Runs the kernels of clonealign_b200/csrc/kernels_interp.cuh on the CPU emulation (tests/cuda_emul/cuda_emul.h) in the
// same order and with the same arguments as run_forward / run_train in core.cu, on inputs read from a binary file.
//   usage: run_interp in.bin out.bin
//   in : int32 N, G, J, smem_panels; float psi[N], w[G], Mx[G*J], Rx[N*J]
//   out: int32 nf_neg, nf_pos, nb; float Zx[N*J], dMx[G*J]
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
  std::vector<float> Zx((size_t)N * J, -1.f), dMx((size_t)G * J, -1.f);
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
                  (const InterpPlan*)&plan, (const double*)vals.data(), split_f, kIMaxPanF, J, 1, coef.data());
  ca_emul::launch(k_interp_eval<true>, dim3(3), dim3(kIEvalWarps * 32), eval_smem, (const InterpPlan*)&plan,
                  (const double*)coef.data(), (const float*)psi.data(), (int64_t)N, J, Zx.data(), smem_panels);

  // backward: nodes over w (reduction over cells) -> coefficients -> evaluation per gene
  if (tj == 8)
    ca_emul::launch(k_interp_nodes2<false, 8>, dim3(4), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)psi.data(),
                    (const float*)shift.data(), (const float*)Rx.data(), (int64_t)N, J, ncgp, split_b, kIMaxPanB, vals.data());
  else
    ca_emul::launch(k_interp_nodes2<false, 6>, dim3(4), dim3(kN2Threads), n2_smem, (const InterpPlan*)&plan, (const float*)psi.data(),
                    (const float*)shift.data(), (const float*)Rx.data(), (int64_t)N, J, ncgp, split_b, kIMaxPanB, vals.data());
  ca_emul::launch(k_interp_coeffs2, dim3((J + kC2Cols - 1) / kC2Cols, 2), dim3(kIP * kC2Cols * kC2Lanes), 0,
                  (const InterpPlan*)&plan, (const double*)vals.data(), split_b, kIMaxPanB, J, 0, coef.data());
  ca_emul::launch(k_interp_eval<false>, dim3(2), dim3(kIEvalWarps * 32), eval_smem, (const InterpPlan*)&plan,
                  (const double*)coef.data(), (const float*)w.data(), (int64_t)G, J, dMx.data(), smem_panels);

  FILE* o = fopen(argv[2], "wb");
  int oh[3] = {plan.nf_neg, plan.nf_pos, plan.nb};
  fwrite(oh, 4, 3, o);
  fwrite(Zx.data(), 4, Zx.size(), o);
  fwrite(dMx.data(), 4, dMx.size(), o);
  fclose(o);
  return 0;
}
