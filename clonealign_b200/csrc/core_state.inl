// core.cu, part 2: ca_data / ca_handle (device state of one session).
// Part of the single translation unit core.cu (included from there; not compiled on its own).
// Device-resident inputs of a fit that do not depend on the restart: the count matrix as stored and
// everything derived from it once (library sizes, B = Y log L, multinomial constants, column sums, allele term).
// Sessions created with ca_core_create_shared read them in place (read-only), so the restarts of run_clonealign
// (R/clonealign.R:50-56) upload and preprocess Y once per device instead of once per fit.
struct ca_data {
  int dev = 0;
  int64_t N = 0, ldY = 0;
  int G = 0, C = 0, V = 0, ystore = CA_STORE_F32, poison = 0;
  double const_sum = 0.0;
  void* Y = nullptr;
  float *L = nullptr, *Bm = nullptr, *vA = nullptr, *s = nullptr, *colsum = nullptr, *snv = nullptr;
  std::vector<void*> allocs;
  std::atomic<int> refs{0};   // sessions created from it (host threads of concurrent restarts create / destroy them)
};

struct ca_handle {
  ca_config cfg{};
  ca_data* shared = nullptr;       // inputs owned by a ca_data (ca_core_create_shared), else by this handle
  bool data_only = false;          // ca_core_data_create: stop after the Y-derived part of build()
  int dev = 0, num_sms = 148;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;
  int64_t N = 0, Ntot = 0, ldY = 0, Gld = 0, Nld = 0;
  int G = 0, C = 0, S = 0, K = 0, P = 0, KP = 0, SC = 0, SCp = 0, J = 0, V = 0;
  bool tc = false;
  bool interp = false;             // K = 1 univariate-interpolation path (kernels_interp.cuh)
  uint32_t variants = 0;           // enum ca_variant bits
  bool epi2 = false;               // interp path: fused Clenshaw + per-cell epilogue (kernels_fused.cuh)
  bool lean = false;               // with epi2: k_prologue / k_gene_fused / k_adam_all
  bool defer = false;              // with lean: Y-linear terms added after the per-cell kernel (late join of the Y pass)
  int y4_minb = 4;                 // k_ypass_k1_v4 register budget: sized for 4 (64 registers) or 3 (80) CTAs per SM
  bool y7 = false;                 // CLONEALIGN_B200_Y5_SPEC=2: k_ypass_k1_v7 (stages as 2-D tensor copies); 6 consumers, 1 536-column tiles
  Y7Plan y7plan;
  bool y7_auto = false;            // chosen by path = auto (core_build.inl)
  bool y5_spec = false;            // k_ypass_k1_v6: the same arithmetic, stages handed over through mbarriers (producer warp + 8 consumer warps)
  int y5_warps = 8;                // warps of a k_ypass_k1_v5 CTA (8 x 128 registers or 16 x 64)
  bool ypass5 = false;             // with ypass4, counts stored as u8: integer tensor-pipe Y pass (k_ypass_k1_v5), one persistent CTA per SM
  bool cosched = false;            // with defer + ypass4: the Y pass starts first in the step, next to everything up to the gene kernel
  bool pending_join = false;       // a Y pass forked onto stream2 has not been joined yet
  int n_yv_blocks = 0;             // ELBO partials written by k_yv_dot (behind the per-cell kernel's in elbo_part)
  double* chi_cur = nullptr;
  float* pmm_part = nullptr;
  unsigned* ticket = nullptr;
  int gene_panels = 0;
  size_t gene_smem = 0;
  int fused_nj = 0, fused_panels = 0, fused_warps = kFusedWarps;
  size_t fused_smem = 0;
  int Jn = 0;                      // columns of the node sums: J, or S*C for the CELL2 set (derivative columns from the interpolants)
  bool cell2 = false;              // with epi2 + lean + defer, S <= 8: k_cell_fused2 (kernels_cell.cuh) is the per-cell kernel
  int cell2_wc = 16, cell2_sb = 8, cell2_panels = 0;
  size_t cell2_smem = 0;
  double* icoef2 = nullptr;        // monomial coefficient pairs [panel][kIP / 2][J][2] (k_interp_coeffs2)
  double* ipart2 = nullptr;        // k_interp_coeffs3: one partial per (slice group, panel, node, column)
  unsigned* itickets = nullptr;    // ... and one arrival counter per column group
  int gene2_panels = 0;            // k_gene_fused2: panels staged at a time, dynamic shared memory
  size_t gene2_smem = 0;
  int64_t n_cell_parts = 0;        // per-block ELBO / sum-gamma partials written by the per-cell kernel in use
  InterpPlan* iplan = nullptr;
  int n2_tj = 8, n2_ncgp = 32, n2_split_f = 1, n2_split_b = 1, n2_blocks_per_sm = kN2BlocksPerSM;   // k_interp_nodes2 launch geometry
  size_t n2_smem = 0;
  float* mm_psi = nullptr;
  double *ivals = nullptr, *icoef = nullptr;
  size_t ieval_smem = 0;
  int ieval_panels = 0;
  int ystore = CA_STORE_F32;
  int poison = 0;
  std::vector<void*> allocs;

  void* Y = nullptr;
  float *L = nullptr, *Bm = nullptr, *vA = nullptr, *s = nullptr, *colsum = nullptr, *snv = nullptr;
  double const_sum = 0.0;
  // trainable + Adam state + gradients
  float *U = nullptr, *Vm = nullptr, *chi_raw = nullptr, *u = nullptr, *loc = nullptr, *lsd = nullptr, *t = nullptr;
  float *m_U = nullptr, *v_U = nullptr, *m_V = nullptr, *v_V = nullptr, *m_chi = nullptr, *v_chi = nullptr;
  float *m_u = nullptr, *v_u = nullptr, *m_loc = nullptr, *v_loc = nullptr, *m_lsd = nullptr, *v_lsd = nullptr;
  float *m_t = nullptr, *v_t = nullptr;
  float *g_U = nullptr, *g_V = nullptr, *g_chi = nullptr, *g_u = nullptr, *g_loc = nullptr, *g_lsd = nullptr, *g_t = nullptr;
  // per-iteration scratch
  float *eps_in = nullptr, *eps = nullptr, *mu = nullptr, *logmu = nullptr, *sig = nullptr;
  float *Mx = nullptr, *shift = nullptr, *mm = nullptr, *Zx = nullptr, *Rx = nullptr, *dMx = nullptr, *dM_sum = nullptr;
  __nv_bfloat16 *MxT_hi = nullptr, *MxT_lo = nullptr;
  __half* RxT = nullptr;
  float* shift_bwd = nullptr;
  float *rowpart = nullptr, *colpart = nullptr, *YV = nullptr, *YtU = nullptr, *Fout = nullptr, *log_alpha = nullptr;
  float* ar = nullptr;
  double *gsum_part = nullptr, *elbo_part = nullptr, *gene_part = nullptr, *scal_elbo = nullptr, *cell_sum = nullptr,
         *wsq = nullptr, *elbo_dev = nullptr;
  int nCB = 1, nRB = 1, RB = 512, n_gene_blocks = 0, nsplit = 1;
  int64_t n_epi_blocks = 0;
  bool ydirty = true;
  bool t_done = false;             // the per-cell kernel of the step in flight has updated the gamma logits itself
  double* cp_scratch = nullptr;    // ca_core_params: clone_probs [N][C] in fp64 before the download
  bool apply_now = false;          // the train step in flight applies its updates (false: ca_core_grads)
  bool inspect = false;            // test hook (ca_core_grads): also write inspection-only arrays (Z of the fused kernel)
  TcPlan tcplan;

  std::vector<float> eps_queue;   // host-fed draws, S*G floats each
  int64_t eps_q_head = 0;         // next draw to consume
  uint64_t draw = 0;
  int adam_t = 0;

  StepState* dstate = nullptr;     // device-side counters (draw, adam_t, lr_t, p2p_step): constant launch arguments
  cudaGraphExec_t g_train[2] = {nullptr, nullptr}, g_eval[2] = {nullptr, nullptr};   // replayable step / evaluation, by ydirty
  bool use_graph = false;
  void* comm = nullptr;
  // variant P2P: exchange buffer of this rank, the peers' mappings, step counter
  bool p2p = false, p2p_ready = false;
  float* p2p_buf = nullptr;
  int64_t p2p_cnt = 0, p2p_cnt_pad = 0;
  float* p2p_slots[kP2PMaxWorld] = {};
  unsigned* p2p_flags[kP2PMaxWorld] = {};
  void* p2p_mapped[kP2PMaxWorld] = {};
  unsigned* p2p_ticket = nullptr;
  int* p2p_err = nullptr;
  unsigned p2p_step = 0;
  bool prof_on = false;
  bool prof_overlap = false;       // CLONEALIGN_B200_PROF_OVERLAP=1: profile the step as it runs (Y pass on its own stream), with start offsets
  std::vector<Prof> prof;
  int launches_last_step = 0;

  template <typename T> T* alloc(size_t n, bool zero = true) {
    void* p = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    CUDA_OK(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    if (zero) CUDA_OK(cudaMemsetAsync(p, 0, bytes, stream));
    return (T*)p;
  }
  void release(void* p) {
    for (auto& q : allocs)
      if (q == p) { cudaFree(p); q = nullptr; }
  }
};
