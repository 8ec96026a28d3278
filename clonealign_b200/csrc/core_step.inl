// core.cu, part 3: launch sequences of a forward pass / train step / ELBO evaluation and their CUDA-graph replay.
// Part of the single translation unit core.cu (included from there; not compiled on its own).
namespace {

struct LaunchScope {
  ca_handle* h;
  bool on;
  size_t idx = 0;
  cudaStream_t st;
  LaunchScope(ca_handle* h_, const char* name, int n_kernels = 1, cudaStream_t st_ = nullptr) : h(h_), on(h_->prof_on), st(st_ ? st_ : h_->stream) {
    h->launches_last_step += n_kernels;
    if (on) {
      Prof p;
      p.name = name;
      CUDA_OK(cudaEventCreate(&p.a));
      CUDA_OK(cudaEventCreate(&p.b));
      CUDA_OK(cudaEventRecord(p.a, st));
      h->prof.push_back(p);
      idx = h->prof.size() - 1;
    }
  }
  ~LaunchScope() {
    if (on) cudaEventRecord(h->prof[idx].b, st);
  }
};
#define KCHECK() CUDA_OK(cudaGetLastError())
// timing ablation (results INVALID): CLONEALIGN_B200_DBG_SKIP = bit mask of launches of the default train step that are not issued
// (1 prologue, 2 forward node sums + coefficients, 4 per-cell kernel, 8 backward node sums + coefficients, 16 gene kernel, 32 optimiser);
// read at every non-replayed step: ca_core_profile_step shows what the Y pass costs next to each of the others (bench.py: CA_BENCH_ABLATE)
inline unsigned dbg_skip() { const char* e = getenv("CLONEALIGN_B200_DBG_SKIP"); return e ? (unsigned)atoi(e) : 0u; }

template <typename F>
void dispatch_y(ca_handle* h, F&& f) {
  switch (h->ystore) {
    case CA_STORE_F32: f((const float*)h->Y); break;
    case CA_STORE_U16: f((const uint16_t*)h->Y); break;
    case CA_STORE_U8: f((const uint8_t*)h->Y); break;
    default: fail("bad y_store");
  }
}

AdamHyper adam_hyper(ca_handle* h, bool apply) {
  AdamHyper a;
  int t = h->adam_t + 1;
  double b1 = 0.9, b2 = 0.999;
  a.lr_t = (float)(h->cfg.learning_rate * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t)));
  a.b1 = 0.9f;
  a.b2 = 0.999f;
  a.eps = 1e-8f;
  a.apply = apply ? 1 : 0;
  return a;
}

// ---- the Y pass (K3) ---------------------------------------------------------------------------
void run_ypass(ca_handle* h, cudaStream_t st) {
  if (h->KP == 0 || !h->ydirty) return;
  dispatch_y(h, [&](auto* Yp) {
    using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
    if (h->KP == 1) {
      LaunchScope ls(h, "ypass", 1, st);
      dim3 grid(h->nCB, h->nRB);
      if (h->ypass5 && std::is_same<T, uint8_t>::value) {
        const int64_t tiles = (int64_t)h->nCB * h->nRB;
        const unsigned g5 = (unsigned)std::min<int64_t>(tiles, (int64_t)h->num_sms);
        if (h->y7) {
          y7_launch(h->y7plan, (unsigned)h->num_sms, st, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart, h->colpart);
        } else if (h->y5_spec)
          CA_LAUNCH(k_ypass_k1_v6, g5, kY6Threads, ypass6_smem_bytes(), st)((const uint8_t*)(const void*)Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm,
                                                                           h->rowpart, h->colpart);
        else if (h->y5_warps == 8)
          CA_LAUNCH(k_ypass_k1_v5<8>, g5, 256, ypass5_smem_bytes(), st)((const uint8_t*)(const void*)Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm,
                                                                       h->rowpart, h->colpart);
        else
          CA_LAUNCH(k_ypass_k1_v5<16>, g5, 512, ypass5_smem_bytes(), st)((const uint8_t*)(const void*)Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm,
                                                                        h->rowpart, h->colpart);
      } else if (h->variants & CA_VAR_YPASS4) {
        const int64_t tiles = (int64_t)h->nCB * h->nRB;
        // persistent grid: 2 CTAs per SM (64 KB rings); fp32 storage has 128 KB rings: one per SM
        const int per_sm = std::is_same<T, float>::value ? 1 : 2;
        const unsigned g4 = (unsigned)std::min<int64_t>(tiles, per_sm * (int64_t)h->num_sms);
        if (h->y4_minb == 3) {
          auto k = k_ypass_k1_v4<T, 3>;
          CA_LAUNCH(k, g4, 256, ypass4_smem_bytes<T>(), st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart, h->colpart);
        } else {
          auto k = k_ypass_k1_v4<T, 4>;
          CA_LAUNCH(k, g4, 256, ypass4_smem_bytes<T>(), st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart, h->colpart);
        }
      } else if (h->variants & CA_VAR_YPASS3) {
        CA_LAUNCH(k_ypass_k1_v3<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else if (h->variants & CA_VAR_YPASS2) {
        CA_LAUNCH(k_ypass_k1_v2<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else if (st == h->stream && !getenv("CLONEALIGN_B200_YPASS_LIGHT")) {
        CA_LAUNCH(k_ypass_k1<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else {
        int64_t tiles = (int64_t)h->nCB * h->nRB;
        unsigned g = (unsigned)std::min<int64_t>(tiles, 2 * (int64_t)h->num_sms);
        CA_LAUNCH(k_ypass_k1_persistent<T>, g, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart,
                                                    h->colpart);
      }
      KCHECK();
    } else {
      {
        LaunchScope ls(h, "ypass_rows");
        CA_LAUNCH(k_ypass_rows_generic<T>, (unsigned)ceil_div64(h->N, 8), 256, 0, st)(Yp, h->ldY, h->N, h->G, h->KP, h->Vm,
                                                                                 h->rowpart);
        KCHECK();
      }
      {
        LaunchScope ls(h, "ypass_cols");
        dim3 grid((h->G + 127) / 128, h->nRB);
        CA_LAUNCH(k_ypass_cols_generic<T>, grid, 128, 0, st)(Yp, h->ldY, h->N, h->G, h->KP, h->RB, h->U, h->colpart);
        KCHECK();
      }
    }
  });
  h->ydirty = false;
}

// ---- forward: eps -> mu, Mx -> shift -> Zx -> (Y pass) -> epilogue -------------------------------
void stage_eps(ca_handle* h, const float** eps_in) {
  *eps_in = nullptr;
  size_t per = (size_t)h->S * h->G;
  if ((size_t)h->eps_q_head * per < h->eps_queue.size()) {
    CUDA_OK(cudaMemcpyAsync(h->eps_in, h->eps_queue.data() + (size_t)h->eps_q_head * per, per * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // source is pageable host memory owned by the queue
    h->eps_q_head++;
    if ((size_t)h->eps_q_head * per >= h->eps_queue.size()) {
      h->eps_queue.clear();
      h->eps_q_head = 0;
    }
    *eps_in = h->eps_in;
  }
}

// node sums + coefficients of the interp path (kernels_interp.cuh, k_interp_nodes2 / k_interp_coeffs2)
template <bool FWD>
void launch_interp_nodes(ca_handle* h, const float* rv, const float* shift, const float* B, int64_t R) {
  const int nsplit = FWD ? h->n2_split_f : h->n2_split_b;
  const int max_pan = FWD ? kIMaxPanF : kIMaxPanB;
  const int Jn = h->Jn;   // node-sum columns = row pitch of B (CELL2 set: the S*C normaliser columns only)
  const unsigned grid = (unsigned)std::min<int64_t>((int64_t)max_pan * nsplit, (int64_t)h->n2_blocks_per_sm * h->num_sms);
  if (h->n2_tj == 8) {
    auto k = k_interp_nodes2<FWD, 8>;
    CA_LAUNCH(k, grid, kN2Threads, h->n2_smem, h->stream)(h->iplan, rv, shift, B, R, Jn, h->n2_ncgp, nsplit, max_pan, h->ivals);
  } else {
    auto k = k_interp_nodes2<FWD, 6>;
    CA_LAUNCH(k, grid, kN2Threads, h->n2_smem, h->stream)(h->iplan, rv, shift, B, R, Jn, h->n2_ncgp, nsplit, max_pan, h->ivals);
  }
  if (h->cell2)   // CELL2 set: slices split over grid.y as well, the last block of a column group finishes (k_interp_coeffs3)
    CA_LAUNCH(k_interp_coeffs3, dim3((Jn + kC2Cols - 1) / kC2Cols, kC3Groups), kIP * kC2Cols * kC2Lanes, 0, h->stream)(
        h->iplan, h->ivals, nsplit, max_pan, Jn, FWD ? 1 : 0, h->ipart2, h->itickets, h->icoef, h->icoef2);
  else
  CA_LAUNCH(k_interp_coeffs2, dim3((Jn + kC2Cols - 1) / kC2Cols, kC2PanelsY), kIP * kC2Cols * kC2Lanes, 0, h->stream)(
      h->iplan, h->ivals, nsplit, max_pan, Jn, FWD ? 1 : 0, h->icoef, nullptr);
}

// the partial sums of the Y pass are needed from here on: wait for the pass forked onto stream2, or run it now
void join_ypass(ca_handle* h, int mode) {
  if (h->pending_join) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->pending_join = false;
  } else if (mode != EPI_INIT) {
    run_ypass(h, h->stream);
  }
}

template <int MODE>
void launch_fused_mode(ca_handle* h, const FusedArgs& a) {
  const unsigned grid = (unsigned)h->n_cell_parts;
  switch (h->fused_nj) {
    case 1: { auto k = k_cell_fused<MODE, 1>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 2: { auto k = k_cell_fused<MODE, 2>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 3: { auto k = k_cell_fused<MODE, 3>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 4: { auto k = k_cell_fused<MODE, 4>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    default: fail("fused per-cell kernel: unsupported S*C");
  }
}
void launch_fused(ca_handle* h, int mode, const FusedArgs& a) {
  if (mode == EPI_TRAIN) launch_fused_mode<EPI_TRAIN>(h, a);
  else if (mode == EPI_EVAL) launch_fused_mode<EPI_EVAL>(h, a);
  else launch_fused_mode<EPI_INIT>(h, a);
}
template <int NJ>
void fused_set_smem(size_t smem) {
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_TRAIN, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_EVAL, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_INIT, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

// (lanes per cell, samples in registers) -> compile-time constants of k_cell_fused2
template <typename F>
void cell2_dispatch(int wc, int sb, F&& f) {
  auto with_sb = [&](auto wcc) {
    switch (sb) {
      case 1: f(wcc, std::integral_constant<int, 1>{}); break;
      case 2: f(wcc, std::integral_constant<int, 2>{}); break;
      case 4: f(wcc, std::integral_constant<int, 4>{}); break;
      case 8: f(wcc, std::integral_constant<int, 8>{}); break;
      default: fail("per-cell kernel: unsupported sample count");
    }
  };
  switch (wc) {
    case 4: with_sb(std::integral_constant<int, 4>{}); break;
    case 8: with_sb(std::integral_constant<int, 8>{}); break;
    case 16: with_sb(std::integral_constant<int, 16>{}); break;
    case 32: with_sb(std::integral_constant<int, 32>{}); break;
    default: fail("per-cell kernel: unsupported clone count");
  }
}
void launch_cell2(ca_handle* h, int mode, const Cell2Args& a) {
  const unsigned grid = (unsigned)h->n_cell_parts, block = (unsigned)h->fused_warps * 32;
  cell2_dispatch(h->cell2_wc, h->cell2_sb, [&](auto wc, auto sb) {
    constexpr int WC = decltype(wc)::value, SB = decltype(sb)::value;
    if (mode == EPI_TRAIN) { auto k = k_cell_fused2<EPI_TRAIN, WC, SB>; CA_LAUNCH(k, grid, block, h->cell2_smem, h->stream)(a); }
    else if (mode == EPI_EVAL) { auto k = k_cell_fused2<EPI_EVAL, WC, SB>; CA_LAUNCH(k, grid, block, h->cell2_smem, h->stream)(a); }
    else { auto k = k_cell_fused2<EPI_INIT, WC, SB>; CA_LAUNCH(k, grid, block, h->cell2_smem, h->stream)(a); }
  });
}

void run_forward(ca_handle* h, int mode) {
  const float* eps_in;
  stage_eps(h, &eps_in);
  // The Y stream (HBM-bound, touches only Y, psi, W) is independent of the forward contraction (tensor / MUFU
  // bound): fork it onto a second stream so both run on the SMs at once; joined before the per-cell epilogue.
  bool joined_later = false;
  // (per-kernel profiling serialises the step: no fork -- unless the co-scheduled timeline itself is asked for, ca_core_profile_step)
  const bool want_fork = mode != EPI_INIT && (h->overlap || h->cosched) && (!h->prof_on || h->prof_overlap) && h->ydirty && h->KP > 0;
  auto fork_ypass = [&]() {
    CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    run_ypass(h, h->stream2);
    CUDA_OK(cudaEventRecord(h->ev_join, h->stream2));
    joined_later = true;
  };
  // Variant DEFER forks later, right before the per-cell kernel: two Y-pass CTAs take the whole register file of an SM, so a
  // pass started here would only push the short gene-level launches (prologue, node sums, coefficients: the head of the
  // step's critical path) behind its first wave; started together with the per-cell kernel (half a register file per
  // CTA) it shares every SM with it instead.
  if (want_fork && (!h->defer || h->cosched)) fork_ypass();   // cosched: the persistent 2-CTA-per-SM pass goes first, everything else fits next to it
  SampleMuArgs sm;
  sm.G = h->G; sm.C = h->C; sm.S = h->S; sm.K = h->K; sm.KP = h->KP; sm.SCp = h->SCp; sm.J = h->J; sm.Gld = h->Gld;
  sm.loc = h->loc; sm.lsd = h->lsd; sm.Vm = h->Vm; sm.L = h->L; sm.colsum = h->colsum; sm.chi_raw = h->chi_raw;
  sm.eps_in = eps_in; sm.seed = h->cfg.seed; sm.draw = h->draw++;
  sm.eps_out = h->eps; sm.mu = h->mu; sm.logmu = h->logmu; sm.sig = h->sig;
  sm.Mx = h->tc ? nullptr : h->Mx; sm.MxT_hi = h->tc ? h->MxT_hi : nullptr; sm.MxT_lo = h->tc ? h->MxT_lo : nullptr;
  sm.gene_part = h->gene_part; sm.no_w_half = h->cell2 ? 1 : 0; sm.Jm = h->Jn;
  if (h->lean) {
    LaunchScope ls(h, "prologue");
    PrologueArgs a;
    a.N = h->N; a.G = h->G; a.C = h->C; a.K = h->K;
    a.u = h->u; a.chi_raw = h->chi_raw; a.Vm = h->Vm; a.U = h->U;
    a.log_alpha = h->log_alpha; a.mm = h->mm; a.mm_psi = h->mm_psi; a.chi_cur = h->chi_cur;
    a.scal_elbo = h->scal_elbo; a.wsq = h->wsq; a.pmm_part = h->pmm_part; a.ticket = h->ticket; a.plan = h->iplan;
    a.dirichlet_const = (double)h->C * lgamma(1.0 / h->C) - lgamma(1.0);
    a.state = h->dstate; a.lr = h->cfg.learning_rate;
    a.mu = sm;
    a.mu_vec4 = (h->C % 4 == 0) ? 1 : 0;
    a.wide_panels = h->cell2 ? 1 : 0;
    if (!(dbg_skip() & 1u)) CA_LAUNCH(k_prologue, 2 + kProPsiBlocks + h->n_gene_blocks, kProThreads, 0, h->stream)(a);
    KCHECK();
  } else {
    LaunchScope ls(h, "alpha");
    CA_LAUNCH(k_alpha, 1, 32, 0, h->stream)(h->u, h->C, h->chi_raw, h->K, h->log_alpha, h->scal_elbo);
    KCHECK();
  }
  if (!h->lean) {
    LaunchScope ls(h, "sample_mu");
    CA_LAUNCH(k_sample_mu, h->n_gene_blocks, 256, 0, h->stream)(sm);
    KCHECK();
  }
  if (h->KP == 0) {
    CUDA_OK(cudaMemsetAsync(h->shift, 0, sizeof(float) * h->N, h->stream));
  } else if (h->lean) {
    // W range (and sum of squares) come from k_prologue, m_n from the fused per-cell kernel
  } else if (h->K == 1 && h->P == 0) {
    LaunchScope ls(h, "shift", h->epi2 ? 1 : 2);
    CA_LAUNCH(k_minmax, 1, 1024, 0, h->stream)(h->Vm, h->G, h->mm);
    KCHECK();
    if (!h->epi2) {   // EPI2 computes m_n inside the fused per-cell kernel
      CA_LAUNCH(k_shift_k1, (unsigned)ceil_div64(h->N, 256), 256, 0, h->stream)(h->U, h->mm, h->N, h->shift);
      KCHECK();
    }
  } else {
    LaunchScope ls(h, "shift");
    CA_LAUNCH(k_shift_general, (unsigned)ceil_div64(h->N, 8), 256, 0, h->stream)(h->U, h->Vm, h->N, h->G, h->KP, h->shift);
    KCHECK();
  }
  {
    LaunchScope ls(h, "lse_fwd", h->interp ? (h->lean ? 2 : (h->epi2 ? 4 : 5)) : 1);
    if (h->interp) {
      // K = 1: Zx[n][j] = F_j(psi_n) by piecewise Chebyshev interpolation (kernels_interp.cuh)
      if (!h->lean) {
        CA_LAUNCH(k_minmax, 1, 1024, 0, h->stream)(h->U, (int)h->N, h->mm_psi);
        CA_LAUNCH(k_interp_plan, 1, 32, 0, h->stream)(h->mm, h->mm_psi, h->iplan);
      }
      if (!(dbg_skip() & 2u)) launch_interp_nodes<true>(h, h->Vm, nullptr, h->Mx, h->G);
      if (!h->epi2)
        CA_LAUNCH(k_interp_eval<true>, h->num_sms, kIEvalWarps * 32, h->ieval_smem, h->stream)(h->iplan, h->icoef, h->U, h->N, h->J, h->Zx,
                                                                                    h->ieval_panels);
    } else if (h->tc) {
      tc_launch_fwd(h->tcplan, h->U, h->Vm, h->shift, h->Zx, h->stream);
    } else {
      int Jc = (mode == EPI_TRAIN) ? h->J : h->SC;   // ELBO-only passes need just Z
      dim3 grid((Jc + 63) / 64, (unsigned)ceil_div64(h->N, 64));
      CA_LAUNCH(k_expgemm<true>, grid, 256, 0, h->stream)(h->U, h->Vm, h->shift, h->Mx, h->Zx, h->N, h->G, Jc, h->J, h->KP);
    }
    KCHECK();
  }
  h->pending_join = joined_later;
  if (!h->defer) join_ypass(h, mode);
  if (h->epi2) {
    LaunchScope ls(h, mode == EPI_TRAIN ? "cell_epilogue" : (mode == EPI_EVAL ? "cell_epilogue_eval" : "gamma_init"));
    FusedArgs a;
    a.N = h->N; a.C = h->C; a.S = h->S; a.SC = h->SC; a.J = h->J; a.nCB = h->nCB; a.smem_panels = h->fused_panels;
    a.plan = h->iplan; a.coeff = h->icoef; a.mm = h->mm;
    a.U = h->U; a.Bm = h->Bm; a.vA = h->vA; a.s = h->s; a.log_alpha = h->log_alpha; a.rowpart = h->rowpart;
    a.t = h->t; a.gT = h->g_t; a.Rx = h->Rx; a.gU = h->g_U; a.YV = h->YV; a.shift = h->shift;
    a.Fout = h->inspect ? h->Fout : nullptr;     // inspection copies (ca_core_grads): not written by the timed path
    a.Zx = h->inspect ? h->Zx : nullptr;
    a.elbo_part = h->elbo_part; a.gsum_part = h->gsum_part;
    a.defer_yv = h->defer ? 1 : 0;
    // DEFER + OVERLAP: the pass may start once everything before the per-cell kernel is done (event recorded here), but
    // it is handed to the device AFTER the per-cell kernel, whose 148 persistent CTAs should be placed first
    if (want_fork && h->defer && !h->cosched) CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
    if (h->cell2) {
      Cell2Args b;
      b.N = a.N; b.C = a.C; b.S = a.S; b.SC = a.SC; b.J = a.J; b.Jn = h->Jn; b.smem_panels = h->cell2_panels;
      b.plan = a.plan; b.coef2 = reinterpret_cast<const double2*>(h->icoef2); b.mm = a.mm;
      b.U = a.U; b.Bm = a.Bm; b.vA = a.vA; b.s = a.s; b.log_alpha = a.log_alpha;
      b.t = a.t; b.gT = a.gT; b.Rx = a.Rx; b.gU = a.gU; b.Fout = a.Fout; b.shift = a.shift; b.Zx = a.Zx;
      b.elbo_part = a.elbo_part; b.gsum_part = a.gsum_part;
      b.apply_t = (mode == EPI_TRAIN && h->apply_now && !getenv("CLONEALIGN_B200_NO_CELL_ADAM")) ? 1 : 0;
      b.m_t = h->m_t; b.v_t = h->v_t; b.state = h->dstate;
      h->t_done = b.apply_t != 0;
      if (!(dbg_skip() & 4u)) launch_cell2(h, mode, b);
    } else
    launch_fused(h, mode, a);
    KCHECK();
    if (want_fork && h->defer && !h->cosched) {
      CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
      run_ypass(h, h->stream2);
      CUDA_OK(cudaEventRecord(h->ev_join, h->stream2));
      h->pending_join = true;
    }
    if (h->defer && mode == EPI_EVAL) {   // the ELBO needs sum_n psi_n (YW)_n now; a train step joins before k_gene_fused
      join_ypass(h, mode);
      LaunchScope ls2(h, "yv_dot");
      CA_LAUNCH(k_yv_dot, h->n_yv_blocks, 256, 0, h->stream)(h->N, h->nCB, h->rowpart, h->U, h->YV, h->elbo_part + h->n_cell_parts);
      KCHECK();
    }
  } else {
    LaunchScope ls(h, mode == EPI_TRAIN ? "cell_epilogue" : (mode == EPI_EVAL ? "cell_epilogue_eval" : "gamma_init"));
    EpiArgs a;
    a.N = h->N; a.Nld = h->Nld; a.C = h->C; a.S = h->S; a.SCp = h->SCp; a.J = h->J; a.K = h->K; a.KP = h->KP; a.nCB = h->nCB;
    a.fsplit = h->tc ? h->tcplan.fsplit : 1;
    a.Zx = h->Zx; a.Bm = h->Bm; a.vA = h->vA; a.s = h->s; a.shift = h->shift; a.log_alpha = h->log_alpha; a.U = h->U;
    a.rowpart = h->rowpart; a.t = h->t; a.gT = h->g_t; a.Rx = h->tc ? nullptr : h->Rx; a.gU = h->g_U; a.YV = h->YV;
    a.Fout = h->Fout; a.RxT = h->tc ? h->RxT : nullptr; a.shift_bwd = h->shift_bwd; a.elbo_part = h->elbo_part; a.gsum_part = h->gsum_part;
    size_t smem = epi_smem_bytes(h->SCp, h->C, h->J, h->tc);
    unsigned grid = (unsigned)h->n_epi_blocks;
    if (mode == EPI_TRAIN) CA_LAUNCH(k_cell_epilogue<EPI_TRAIN>, grid, kEpiWarps * 32, smem, h->stream)(a);
    else if (mode == EPI_EVAL) CA_LAUNCH(k_cell_epilogue<EPI_EVAL>, grid, kEpiWarps * 32, smem, h->stream)(a);
    else CA_LAUNCH(k_cell_epilogue<EPI_INIT>, grid, kEpiWarps * 32, smem, h->stream)(a);
    KCHECK();
  }
}

void run_train(ca_handle* h, bool apply) {
  h->launches_last_step = 0;
  h->apply_now = apply;
  h->t_done = false;
  run_forward(h, EPI_TRAIN);
  {
    LaunchScope ls(h, "lse_bwd", h->interp ? (h->lean ? 2 : 3) : 1);
    if (h->interp) {
      // K = 1: dMx[g][j] = H_j(w_g); the plan of this step's forward pass is still valid (psi, W unchanged)
      if (!(dbg_skip() & 8u)) launch_interp_nodes<false>(h, h->U, h->shift, h->Rx, h->N);
      if (!h->lean)
        CA_LAUNCH(k_interp_eval<false>, h->num_sms, kIEvalWarps * 32, h->ieval_smem, h->stream)(h->iplan, h->icoef, h->Vm, h->G, h->J, h->dMx,
                                                                                     h->ieval_panels);
    } else if (h->tc) {
      tc_launch_bwd(h->tcplan, h->U, h->Vm, h->shift_bwd, h->dMx, h->stream);
    } else {
      dim3 grid((h->J + 63) / 64, (h->G + 63) / 64);
      CA_LAUNCH(k_expgemm<false>, grid, 256, 0, h->stream)(h->Vm, h->U, h->shift, h->Rx, h->dMx, h->G, h->N, h->J, h->J, h->KP);
    }
    KCHECK();
  }
  if (h->cell2) {
    // CELL2 set: the gene kernel does not read the Y-pass partials -- it runs before the join, next to the Y pass; the
    // column partials are added to the w-gradient slots by a short launch behind the join
    {
      LaunchScope ls(h, "gene_grads", 1);
      Gene2Args a;
      a.G = h->G; a.C = h->C; a.S = h->S; a.SC = h->SC; a.J = h->J; a.Jn = h->Jn; a.smem_panels = h->gene2_panels;
      a.plan = h->iplan; a.coef2 = reinterpret_cast<const double2*>(h->icoef2);
      a.Vm = h->Vm; a.mu = h->mu; a.sig = h->sig; a.eps = h->eps; a.lsd = h->lsd; a.L = h->L;
      a.ar = h->ar; a.dM_out = h->inspect ? h->dM_sum : nullptr;
      a.gsum_part = h->gsum_part; a.n_parts = h->n_cell_parts;
      cell2_dispatch(h->cell2_wc, h->cell2_sb, [&](auto wc, auto sb) {
        auto k = k_gene_fused2<decltype(wc)::value, decltype(sb)::value>;
        if (!(dbg_skip() & 16u)) CA_LAUNCH(k, 2 * h->num_sms + 1, kGene2Warps * 32, h->gene2_smem, h->stream)(a);
      });
      KCHECK();
    }
    join_ypass(h, EPI_TRAIN);                // colpart and rowpart (d psi in k_adam_all) are needed from here on
    if (h->cfg.world > 1) {                  // (an unsharded fit adds the column partials inside k_adam_all)
      LaunchScope ls(h, "colpart_add", 1);
      CA_LAUNCH(k_colpart_add, (h->G + kColAddGenes - 1) / kColAddGenes, kColAddGenes * kColAddSlices, 0, h->stream)(
          h->G, h->nRB, h->colpart, h->ar + 2 * (int64_t)h->G, h->YtU);
      KCHECK();
    }
  } else {
  if (h->defer) join_ypass(h, EPI_TRAIN);   // colpart (gene gradients) and rowpart (d psi in k_adam_all) are needed from here on
  if (h->lean) {
    LaunchScope ls(h, "gene_grads", 1);
    GeneFusedArgs a;
    a.G = h->G; a.C = h->C; a.S = h->S; a.SC = h->SC; a.J = h->J; a.nRB = h->nRB; a.smem_panels = h->gene_panels;
    a.plan = h->iplan; a.coeff = h->icoef;
    a.Vm = h->Vm; a.colpart = h->colpart; a.mu = h->mu; a.sig = h->sig; a.eps = h->eps; a.lsd = h->lsd; a.L = h->L;
    a.ar = h->ar; a.YtU = h->YtU; a.dM_out = h->inspect ? h->dM_sum : nullptr;
    a.gsum_part = h->gsum_part; a.n_parts = h->n_cell_parts;
    // two 512-thread blocks per SM (64 registers, <= 78 KB of coefficients each): 32 warps keep the fp64 recurrences fed
    switch ((h->SC + 31) / 32) {
      case 1: { auto k = k_gene_fused<1>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      case 2: { auto k = k_gene_fused<2>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      case 3: { auto k = k_gene_fused<3>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      default: { auto k = k_gene_fused<4>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
    }
    KCHECK();
  } else {
    LaunchScope ls(h, "gene_grads", 2);
    GeneGradArgs a;
    a.G = h->G; a.C = h->C; a.S = h->S; a.K = h->K; a.KP = h->KP; a.SCp = h->SCp; a.J = h->J; a.nsplit = h->nsplit; a.nRB = h->nRB;
    a.dMx = h->dMx; a.colpart = h->colpart; a.mu = h->mu; a.sig = h->sig; a.eps = h->eps; a.lsd = h->lsd; a.L = h->L;
    a.ar = h->ar; a.YtU = h->YtU; a.dM_out = h->dM_sum;
    CA_LAUNCH(k_gene_grads_warp, (h->G + 7) / 8, 256, 0, h->stream)(a);
    KCHECK();
    CA_LAUNCH(k_reduce_gsum, 1, 1024, 0, h->stream)(h->gsum_part, h->n_cell_parts, h->C, h->ar + (int64_t)h->G * (2 + h->KP));
    KCHECK();
  }
  }   // !cell2
  if (h->cfg.world > 1 && h->p2p) {
    if (!h->p2p_ready) fail("variant p2p: ca_core_p2p_connect has not been called");
    LaunchScope ls(h, "allreduce");
    P2PArgs a;
    a.world = h->cfg.world; a.rank = h->cfg.rank; a.cnt = h->p2p_cnt; a.cnt_pad = h->p2p_cnt_pad; a.step_ctr = &h->dstate->p2p_step;
    a.src = h->ar; a.dst = h->ar; a.ticket = h->p2p_ticket; a.error = h->p2p_err;
    for (int r = 0; r < kP2PMaxWorld; ++r) { a.slots[r] = h->p2p_slots[r]; a.flags[r] = h->p2p_flags[r]; }
    CA_LAUNCH(k_p2p_allreduce, std::min(h->num_sms, 64), kP2PThreads, 0, h->stream)(a);
    KCHECK();
  } else if (h->cfg.world > 1) {
    LaunchScope ls(h, "allreduce");
    size_t cnt = (size_t)h->G * (2 + h->KP) + h->C;
    NCCL_OK(nccl().AllReduce(h->ar, h->ar, cnt, kNcclFloat32, kNcclSum, h->comm, h->stream));
  }
  {
    LaunchScope ls(h, "adam", h->lean ? 1 : (apply ? 4 : 3));
    AdamHyper hy = adam_hyper(h, apply);
    if (!h->lean) {
      CA_LAUNCH(k_wsq, 1, 1024, 0, h->stream)(h->Vm, h->G, h->K, h->KP, h->wsq);
      KCHECK();
    }
    ScalarAdamArgs sa;
    sa.G = h->G; sa.C = h->C; sa.K = h->K; sa.n_total = (double)h->Ntot; sa.wsq = h->wsq;
    sa.gsum = h->ar + (int64_t)h->G * (2 + h->KP);
    sa.chi_raw = h->chi_raw; sa.m_chi = h->m_chi; sa.v_chi = h->v_chi; sa.g_chi = h->g_chi;
    sa.u = h->u; sa.m_u = h->m_u; sa.v_u = h->v_u; sa.g_u = h->g_u; sa.h = hy;
    GeneAdamArgs ga;
    ga.G = h->G; ga.S = h->S; ga.K = h->K; ga.KP = h->KP; ga.ar = h->ar; ga.mu = h->mu; ga.logmu = h->logmu; ga.sig = h->sig;
    ga.eps = h->eps; ga.colsum = h->colsum; ga.chi_raw = h->chi_raw; ga.loc = h->loc; ga.lsd = h->lsd; ga.Vm = h->Vm;
    ga.m_loc = h->m_loc; ga.v_loc = h->v_loc; ga.m_lsd = h->m_lsd; ga.v_lsd = h->v_lsd; ga.m_V = h->m_V; ga.v_V = h->v_V;
    ga.g_loc = h->g_loc; ga.g_lsd = h->g_lsd; ga.g_V = h->g_V; ga.h = hy;
    if (h->lean) {
      AdamAllArgs aa;
      aa.ga = ga; aa.chi_cur = h->chi_cur; aa.sa = sa; aa.N = h->N; aa.C = h->C;
      aa.t = h->t; aa.m_t = h->m_t; aa.v_t = h->v_t; aa.U = h->U; aa.m_U = h->m_U; aa.v_U = h->v_U; aa.gT = h->g_t; aa.gU = h->g_U;
      aa.n_gene_blocks = (h->G + 255) / 256;
      aa.t_done = h->t_done ? 1 : 0;
      aa.n_cell_blocks = (apply || h->defer) ? ceil_div64((aa.t_done ? 0 : ceil_div64(h->N * h->C, 4)) + h->N, 256) : 0;
      aa.defer_yv = h->defer ? 1 : 0; aa.nCB = h->nCB; aa.rowpart = h->rowpart; aa.YV = h->YV;
      aa.state = h->dstate;
      const bool adam_adds_colpart = h->cell2 && h->cfg.world == 1;
      aa.colpart = adam_adds_colpart ? h->colpart : nullptr; aa.nRB = h->nRB; aa.YtU = h->YtU;
      if (!(dbg_skip() & 32u)) CA_LAUNCH(k_adam_all, (unsigned)(aa.n_gene_blocks + aa.n_cell_blocks + 1), 256, 0, h->stream)(aa);
      KCHECK();
    } else {
    // gene kernel reads chi_raw (old) -> must precede the scalar update
    CA_LAUNCH(k_gene_adam, (h->G + 127) / 128, 128, 0, h->stream)(ga);
    KCHECK();
    CA_LAUNCH(k_scalar_adam, 1, 32, 0, h->stream)(sa);
    KCHECK();
    }
    if (apply && !h->lean) {
      int64_t tot = h->N * h->C + h->N * h->KP;
      CA_LAUNCH(k_cell_adam, (unsigned)ceil_div64(tot, 256), 256, 0, h->stream)(h->N, h->C, h->K, h->KP, h->t, h->m_t, h->v_t, h->g_t,
                                                                       h->U, h->m_U, h->v_U, h->g_U, hy);
      KCHECK();
    }
  }
  if (apply) {
    h->adam_t++;
    h->ydirty = true;
  }
}

void run_elbo_async(ca_handle* h) {
  h->launches_last_step = 0;
  run_forward(h, EPI_EVAL);
  LaunchScope ls(h, "elbo_reduce", 2);
  CA_LAUNCH(k_reduce_partials, 1, 1024, 0, h->stream)(h->elbo_part, h->n_cell_parts + h->n_yv_blocks, 1, h->cell_sum, h->const_sum);
  KCHECK();
  if (h->cfg.world > 1) NCCL_OK(nccl().AllReduce(h->cell_sum, h->cell_sum, 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
  CA_LAUNCH(k_elbo_final, 1, 256, 0, h->stream)(h->cell_sum, h->gene_part, h->n_gene_blocks, h->scal_elbo, h->poison, h->elbo_dev);
  KCHECK();
}

// ---- CUDA-graph replay of the train step / the ELBO evaluation --------------------------------------
// The fused (lean) kernel set keeps everything that changes from step to step in device memory (StepState), so the
// launches of a step have constant arguments: the sequence is captured once per (kind, "Y pass needed") and replayed with
// one cudaGraphLaunch -- 7-9 launches, the fork / join of the Y-pass stream and the all-reduce of a sharded fit included.
// Not used with host-fed draws (test hook), per-kernel profiling or inspection copies; CLONEALIGN_B200_NO_GRAPH=1 disables it.
bool graph_ok(ca_handle* h) {
  return kGraphsAvailable && h->use_graph && h->lean && !h->prof_on && !h->inspect && h->eps_queue.empty();
}
template <typename F>
void capture_or_replay(ca_handle* h, cudaGraphExec_t& exec, F&& body, const std::function<void()>& host_effects) {
  if (!exec) {
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    try {
      body();                                  // also applies the host-side bookkeeping once
    } catch (...) {
      cudaStreamEndCapture(h->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_OK(cudaStreamEndCapture(h->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OK(e);
  } else {
    host_effects();
  }
  CUDA_OK(cudaGraphLaunch(exec, h->stream));
}
void train_step(ca_handle* h) {
  if (graph_ok(h)) {
    const int key = h->ydirty ? 1 : 0;
    const int n_launch = h->launches_last_step;
    capture_or_replay(h, h->g_train[key], [&] { run_train(h, true); },
                      [&] { h->draw++; h->adam_t++; h->ydirty = true; h->pending_join = false; (void)n_launch; });
    return;
  }
  run_train(h, true);
}
void eval_step(ca_handle* h) {
  if (graph_ok(h)) {
    const int key = h->ydirty ? 1 : 0;
    capture_or_replay(h, h->g_eval[key], [&] { run_elbo_async(h); },
                      [&] { h->draw++; if (h->KP > 0) h->ydirty = false; h->pending_join = false; });
    return;
  }
  run_elbo_async(h);
}

}  // namespace
