// core.cu, part 4: host <-> device helpers, ingest of Y, construction / destruction of a session, array lookup.
// Part of the single translation unit core.cu (included from there; not compiled on its own).
namespace {

// ---- host <-> device helpers ---------------------------------------------------------------------
void upload_colmajor(ca_handle* h, const double* src, int64_t rows, int cols, float* dst, int ld_dst, int col_off) {
  if (!src || rows * cols == 0) return;
  double* tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, sizeof(double) * rows * cols));
  CUDA_OK(cudaMemcpyAsync(tmp, src, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, h->stream));
  CA_LAUNCH(k_colmajor_to_rowmajor_f, (unsigned)ceil_div64(rows * cols, 256), 256, 0, h->stream)(tmp, rows, cols, dst, ld_dst, col_off);
  KCHECK();
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaFree(tmp));
}

// device row-major float [rows][ld] (columns col_off..col_off+cols) -> host column-major double
void download_colmajor(ca_handle* h, const float* src, int64_t rows, int cols, int ld, int col_off, double* out) {
  if (!out || rows * cols == 0) return;
  std::vector<float> tmp((size_t)rows * ld);
  CUDA_OK(cudaMemcpyAsync(tmp.data(), src, sizeof(float) * rows * ld, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < cols; ++c)
    for (int64_t r = 0; r < rows; ++r) out[(int64_t)c * rows + r] = (double)tmp[(size_t)r * ld + col_off + c];
}

template <typename Tin>
void ingest_y(ca_handle* h, const Tin* Ysrc, float* Yf) {
  const ca_config& c = h->cfg;
  const int64_t N = h->N;
  const int G = h->G;
  const bool on_dev = c.y_mem == CA_Y_DEVICE;
  if (c.y_layout == CA_Y_CSR) {
    if (on_dev) fail("CSR input must be in host memory");
    if (!c.y_indptr || !c.y_indices) fail("CSR input needs y_indptr and y_indices");
    const int32_t* ip = c.y_indptr;
    if (ip[0] < 0) fail("bad CSR row offsets");
    for (int64_t r = 0; r < N; ++r)
      if (ip[r + 1] < ip[r]) fail("bad CSR row offsets");
    int *d_ip = nullptr, *d_idx = nullptr, *d_bad = nullptr;
    Tin* d_val = nullptr;
    const int64_t cap = std::max<int64_t>(1, (int64_t)(128ll << 20) / (int64_t)(sizeof(Tin) + sizeof(int)));   // entries per chunk
    CUDA_OK(cudaMalloc(&d_ip, sizeof(int) * (N + 1)));
    CUDA_OK(cudaMalloc(&d_bad, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
    CUDA_OK(cudaMemcpyAsync(d_ip, ip, sizeof(int) * (N + 1), cudaMemcpyHostToDevice, h->stream));
    int64_t r0 = 0;
    int64_t cur_cap = 0;
    while (r0 < N) {
      // rows [r0, r1) whose entries fit in one chunk (a single row longer than the chunk gets a chunk of its own)
      int64_t r1 = r0 + 1;
      while (r1 < N && (int64_t)ip[r1 + 1] - ip[r0] <= cap) ++r1;
      const int64_t base = ip[r0], cnt = (int64_t)ip[r1] - base;
      if (cnt > cur_cap) {
        if (d_idx) { CUDA_OK(cudaFree(d_idx)); CUDA_OK(cudaFree(d_val)); }
        cur_cap = std::max(cnt, cap);
        CUDA_OK(cudaMalloc(&d_idx, sizeof(int) * cur_cap));
        CUDA_OK(cudaMalloc(&d_val, sizeof(Tin) * cur_cap));
      }
      if (cnt > 0) {
        CUDA_OK(cudaMemcpyAsync(d_idx, c.y_indices + base, sizeof(int) * cnt, cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaMemcpyAsync(d_val, Ysrc + base, sizeof(Tin) * cnt, cudaMemcpyHostToDevice, h->stream));
        CA_LAUNCH(k_ingest_csr<Tin>, (unsigned)ceil_div64(r1 - r0, 8), 256, 0, h->stream)(d_ip, d_idx, d_val, base, r0, r1 - r0, G, Yf,
                                                                                          h->ldY, d_bad);
        KCHECK();
        CUDA_OK(cudaStreamSynchronize(h->stream));
      }
      r0 = r1;
    }
    int hbad = 0;
    CUDA_OK(cudaMemcpyAsync(&hbad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d_ip); cudaFree(d_bad);
    if (d_idx) { cudaFree(d_idx); cudaFree(d_val); }
    if (hbad) fail("CSR input has a gene index outside [0, G)");
  } else if (c.y_layout == CA_Y_COLMAJOR) {
    int64_t ld = c.y_ld ? c.y_ld : N;
    int gchunk = (int)std::max<int64_t>(1, std::min<int64_t>(G, (int64_t)(256ll << 20) / (int64_t)(sizeof(Tin) * N)));
    Tin* stage = nullptr;
    if (!on_dev) CUDA_OK(cudaMalloc(&stage, sizeof(Tin) * (size_t)gchunk * N));
    for (int g0 = 0; g0 < G; g0 += gchunk) {
      int gc = std::min(gchunk, G - g0);
      const Tin* src;
      int64_t ld_in;
      if (on_dev) {
        src = Ysrc + (int64_t)g0 * ld;
        ld_in = ld;
      } else {
        CUDA_OK(cudaMemcpy2DAsync(stage, sizeof(Tin) * N, Ysrc + (int64_t)g0 * ld, sizeof(Tin) * ld, sizeof(Tin) * N, gc,
                                  cudaMemcpyHostToDevice, h->stream));
        src = stage;
        ld_in = N;
      }
      dim3 grid((unsigned)ceil_div64(N, 32), (gc + 31) / 32), blk(32, 8);
      CA_LAUNCH(k_ingest_colmajor<Tin>, grid, blk, 0, h->stream)(src, ld_in, N, g0, gc, Yf, h->ldY);
      KCHECK();
      CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (stage) CUDA_OK(cudaFree(stage));
  } else {
    int64_t ld = c.y_ld ? c.y_ld : G;
    int64_t rchunk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)(256ll << 20) / (int64_t)(sizeof(Tin) * ld)));
    rchunk = std::min<int64_t>(rchunk, 65535);
    Tin* stage = nullptr;
    if (!on_dev) CUDA_OK(cudaMalloc(&stage, sizeof(Tin) * (size_t)rchunk * ld));
    for (int64_t r0 = 0; r0 < N; r0 += rchunk) {
      int64_t rc = std::min(rchunk, N - r0);
      const Tin* src;
      if (on_dev) {
        src = Ysrc + r0 * ld;
      } else {
        CUDA_OK(cudaMemcpyAsync(stage, Ysrc + r0 * ld, sizeof(Tin) * (size_t)rc * ld, cudaMemcpyHostToDevice, h->stream));
        src = stage;
      }
      dim3 grid(std::min((G + 255) / 256, 64), (unsigned)rc);
      CA_LAUNCH(k_ingest_rowmajor<Tin>, grid, 256, 0, h->stream)(src, ld, rc, G, Yf + r0 * h->ldY, h->ldY);
      KCHECK();
      CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (stage) CUDA_OK(cudaFree(stage));
  }
}

void destroy(ca_handle* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) comm_release(h->cfg.world, h->cfg.rank, h->dev, h->comm);   // parked for the next session of this shape
  for (int r = 0; r < kP2PMaxWorld; ++r)
    if (h->p2p_mapped[r]) cudaIpcCloseMemHandle(h->p2p_mapped[r]);
  for (auto* g : {&h->g_train[0], &h->g_train[1], &h->g_eval[0], &h->g_eval[1]})
    if (*g) { cudaGraphExecDestroy(*g); *g = nullptr; }
  tc_plan_destroy(h->tcplan);
  for (auto& p : h->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (void* p : h->allocs)
    if (p) cudaFree(p);
  if (h->shared) h->shared->refs--;
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

void build(ca_handle* h, const void* Y, const double* L, const double* psi_init, const double* loc_init, const double* X,
           const double* colsum_total, const double* clone_allele, const double* alt, const double* cov) {
  const ca_config& c = h->cfg;
  if (c.N <= 0 || c.G <= 0 || c.C <= 0 || c.S <= 0 || c.K < 0 || c.P < 0) fail("bad dimensions");
  if (c.K + c.P > kMaxKP) fail("K + P = %d exceeds the supported maximum of %d", c.K + c.P, kMaxKP);
  if (c.world < 1 || c.rank < 0 || c.rank >= c.world) fail("bad rank/world");
  if (c.world > 1 && !c.nccl_id) fail("world > 1 requires cfg.nccl_id");
  if (!h->shared && (!Y || !L)) fail("missing input pointer");
  if (!h->data_only && (!loc_init || (c.K > 0 && !psi_init) || (c.P > 0 && !X))) fail("missing input pointer");
  if (!h->shared && c.V > 0 && (!clone_allele || !alt || !cov)) fail("V > 0 requires clone_allele, alt and cov");
  if (h->shared) {
    const ca_data* d = h->shared;
    if (c.world != 1) fail("shared inputs are for single-shard sessions (world == 1)");
    if (c.N != d->N || c.G != d->G || c.C != d->C || c.V != d->V || c.device != d->dev)
      fail("session dimensions / device do not match the shared inputs (N %lld G %d C %d V %d device %d)", (long long)d->N, d->G, d->C,
           d->V, d->dev);
  }
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (c.device < 0 || c.device >= ndev) fail("CUDA device %d not available (%d devices)", c.device, ndev);
  h->dev = c.device;
  CUDA_OK(cudaSetDevice(h->dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, h->dev));
  if (prop.major != 10) fail("clonealign_b200 kernels are built for sm_100a only; device %d is sm_%d%d", h->dev, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (c.world > 1) h->comm = comm_acquire(c.world, c.rank, h->dev, c.nccl_id);   // collective (first session of this shape)
  // Measured on B200 (profiles/r01_notes.md): co-scheduling the Y stream with the forward contraction does not pay
  // yet (the register-light Y kernel is slower than the saved time), so the fork is opt-in.
  h->overlap = getenv("CLONEALIGN_B200_OVERLAP") != nullptr || (c.variants & CA_VAR_OVERLAP);
  h->use_graph = getenv("CLONEALIGN_B200_NO_GRAPH") == nullptr;

  h->N = c.N; h->Ntot = c.N_total > 0 ? c.N_total : c.N; h->G = c.G; h->C = c.C; h->S = c.S; h->K = c.K; h->P = c.P;
  h->KP = c.K + c.P; h->SC = c.S * c.C; h->V = c.V;
  bool tc_ok = kTcAvailable && (c.K == 1 && c.P == 0 && round_up64(h->SC, 16) <= 128);
  if (c.path == CA_PATH_TENSOR && !tc_ok) fail("tensor path needs K == 1, P == 0 and S*C <= 128");
  if (c.path == CA_PATH_INTERP && !(c.K == 1 && c.P == 0)) fail("interp path needs K == 1 and P == 0");
  // path = auto: the reference's default model (K = 1, no covariates; K is forced to 1 at R/clonealign.R:226-232) runs the
  // univariate-interpolation kernel set that round 2 validated on hardware (profiles/r02_notes.md): interp + bulk-copy Y
  // pass on the stored integers, co-scheduled with the rest of the step + fused per-cell kernel + fused gene-level
  // launches + late join of the Y pass.  Explicit variant bits of the caller are kept (ypass2 / ypass3, overlap, p2p).  Other shapes: tcgen05 contractions
  // (K = 1, S*C <= 128) or the CUDA-core kernels (any K + P <= 8).
  const bool interp_ok = c.K == 1 && c.P == 0 && c.C <= kFusedMaxC && c.S * c.C <= 32 * kFusedMaxNJ;
  if (c.path == CA_PATH_AUTO && interp_ok) {
    h->cfg.path = CA_PATH_INTERP;
    if (!(h->cfg.variants & (CA_VAR_YPASS2 | CA_VAR_YPASS3 | CA_VAR_YPASS4))) h->cfg.variants |= CA_VAR_YPASS4;
    h->cfg.variants |= CA_VAR_EPI2 | CA_VAR_LEAN | CA_VAR_DEFER;
    if (h->cfg.variants & CA_VAR_YPASS4) h->cfg.variants |= CA_VAR_COSCHED;
    if (c.S <= kCell2MaxS && !getenv("CLONEALIGN_B200_NO_CELL2")) h->cfg.variants |= CA_VAR_CELL2;
    // stored bytes + cell2 + co-scheduling: the integer tensor-pipe Y pass on 2-D tensor copies (k_ypass_k1_v7) where the matrix is large
    // enough to keep one persistent CTA per SM busy (CLONEALIGN_B200_Y7 = 0 / 1 overrides the size rule)
    if (kY7Available && (h->cfg.variants & CA_VAR_YPASS4) && (h->cfg.variants & CA_VAR_CELL2) && c.N > 0 && y7_wanted((int64_t)c.N * c.G)) {
      h->cfg.variants |= CA_VAR_YPASS5;
      h->y7_auto = true;
    }
  }
  h->interp = (c.path == CA_PATH_INTERP);
  h->variants = c.variants;
  if (c.variants & ~(uint32_t)(CA_VAR_YPASS2 | CA_VAR_EPI2 | CA_VAR_LEAN | CA_VAR_P2P | CA_VAR_OVERLAP | CA_VAR_YPASS3 | CA_VAR_DEFER | CA_VAR_YPASS4 | CA_VAR_COSCHED | CA_VAR_CELL2 | CA_VAR_YPASS5))
    fail("unknown kernel variant bits 0x%x", c.variants);
  if ((c.variants & CA_VAR_P2P) && c.world > kP2PMaxWorld) fail("variant p2p supports at most %d ranks", kP2PMaxWorld);
  h->p2p = (c.variants & CA_VAR_P2P) && c.world > 1;
  if ((c.variants & CA_VAR_LEAN) && !(c.variants & CA_VAR_EPI2)) fail("variant lean needs variant epi2");
  h->lean = (c.variants & CA_VAR_LEAN) != 0;
  if ((c.variants & CA_VAR_DEFER) && !(c.variants & CA_VAR_LEAN)) fail("variant defer needs variants epi2 and lean");
  h->defer = (c.variants & CA_VAR_DEFER) != 0;
  if ((c.variants & CA_VAR_YPASS2) && c.K + c.P != 1) fail("variant ypass2 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS3) && c.K + c.P != 1) fail("variant ypass3 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS3) && (c.variants & CA_VAR_YPASS2)) fail("variants ypass2 and ypass3 are alternatives");
  if ((c.variants & CA_VAR_YPASS4) && c.K + c.P != 1) fail("variant ypass4 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS4) && (c.variants & (CA_VAR_YPASS2 | CA_VAR_YPASS3))) fail("variants ypass2, ypass3 and ypass4 are alternatives");
  if ((c.variants & CA_VAR_COSCHED) && !((c.variants & CA_VAR_DEFER) && (c.variants & CA_VAR_YPASS4)))
    fail("variant cosched needs variants defer and ypass4");
  h->cosched = (c.variants & CA_VAR_COSCHED) != 0;
  if ((c.variants & CA_VAR_CELL2) && !(c.variants & CA_VAR_DEFER)) fail("variant cell2 needs variants epi2, lean and defer");
  h->cell2 = (c.variants & CA_VAR_CELL2) != 0 && c.S <= kCell2MaxS;   // more samples than a lane keeps in registers: k_cell_fused
  if (c.variants & CA_VAR_EPI2) {
    if (!h->interp) fail("variant epi2 belongs to the interp path (path = interp)");
    if (c.C > kFusedMaxC || c.S * c.C > 32 * kFusedMaxNJ) fail("variant epi2 needs C <= %d and S*C <= %d", kFusedMaxC, 32 * kFusedMaxNJ);
    h->epi2 = true;
  }
  h->tc = !h->interp && ((c.path == CA_PATH_TENSOR) || (c.path == CA_PATH_AUTO && tc_ok));
  h->SCp = h->tc ? (int)round_up64(h->SC, 16) : h->SC;
  h->J = h->SCp * (1 + h->KP);
  h->ldY = round_up64(h->G, 16);
  h->Gld = round_up64(h->G, 64);
  h->Nld = round_up64(h->N, 64);
  const int64_t N = h->N;
  const int G = h->G, C = h->C, S = h->S, K = h->K, KP = h->KP, J = h->J;

  if (h->shared) {
    const ca_data* d = h->shared;
    h->Y = d->Y; h->ystore = d->ystore; h->L = d->L; h->Bm = d->Bm; h->vA = d->vA; h->s = d->s; h->colsum = d->colsum;
    h->snv = d->snv; h->const_sum = d->const_sum; h->poison = d->poison;
    if (h->ldY != d->ldY) fail("shared inputs: leading dimension mismatch");
  } else {
  // Row-major u8 counts that are to be stored as u8 go straight into their final array: one 2-D copy, no fp32 staging matrix
  // (4 x the bytes, an 8 GB allocation at config 3), no widening / scan / narrowing passes; the set-up reductions read the bytes.
  const bool compact_u8 = c.y_dtype == CA_Y_U8 && c.y_layout == CA_Y_ROWMAJOR && (c.y_store == CA_STORE_AUTO || c.y_store == CA_STORE_U8);
  uint8_t* Yc = nullptr;
  float* Yf = nullptr;
  if (compact_u8) {
    Yc = h->alloc<uint8_t>((size_t)N * h->ldY, h->ldY != G);          // (the padding columns must read as zero)
    const int64_t ld = c.y_ld ? c.y_ld : G;
    CUDA_OK(cudaMemcpy2DAsync(Yc, (size_t)h->ldY, Y, (size_t)ld, (size_t)G, (size_t)N,
                              c.y_mem == CA_Y_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    h->ystore = CA_STORE_U8;
    h->Y = Yc;
  } else {
  // ---- Y -> device fp32 [N][ldY] ----
  Yf = h->alloc<float>((size_t)N * h->ldY);
  switch (c.y_dtype) {
    case CA_Y_F64: ingest_y<double>(h, (const double*)Y, Yf); break;
    case CA_Y_F32: ingest_y<float>(h, (const float*)Y, Yf); break;
    case CA_Y_I32: ingest_y<int>(h, (const int*)Y, Yf); break;
    case CA_Y_U8: ingest_y<uint8_t>(h, (const uint8_t*)Y, Yf); break;
    case CA_Y_U16: ingest_y<uint16_t>(h, (const uint16_t*)Y, Yf); break;
    default: fail("bad y_dtype");
  }
  // ---- narrow storage if exact ----
  int* flags = h->alloc<int>(1);
  {
    dim3 grid(std::min((G + 255) / 256, 64), 1);
    // grid.y is limited to 65535: loop over row chunks
    for (int64_t r0 = 0; r0 < N; r0 += 65535) {
      grid.y = (unsigned)std::min<int64_t>(65535, N - r0);
      CA_LAUNCH(k_scan_y, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, G, flags);
      KCHECK();
    }
  }
  int hflags = 0;
  CUDA_OK(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  int want = c.y_store;
  if (want == CA_STORE_AUTO) want = (hflags & 1) ? CA_STORE_F32 : ((hflags & 2) ? ((hflags & 4) ? CA_STORE_F32 : CA_STORE_U16) : CA_STORE_U8);
  if (want == CA_STORE_U8 && (hflags & 3)) fail("y_store = u8 requested but Y has non-integer, negative or > 255 entries");
  if (want == CA_STORE_U16 && (hflags & 5)) fail("y_store = u16 requested but Y has non-integer, negative or > 65535 entries");
  h->ystore = want;
  h->Y = Yf;
  }   // !compact_u8

  // ---- small inputs ----
  h->L = h->alloc<float>((size_t)G * C);
  upload_colmajor(h, L, G, C, h->L, C, 0);
  std::vector<float> logL((size_t)G * C);
  for (int g = 0; g < G; ++g)
    for (int cc = 0; cc < C; ++cc) {
      double l = L[(size_t)cc * G + g];
      if (!(l > 0.0)) h->poison = 1;   // copy number 0 => 0*log(0) = NaN in the reference
      logL[(size_t)g * C + cc] = l > 0.0 ? (float)log(l) : 0.f;
    }
  float* d_logL = h->alloc<float>((size_t)G * C);
  CUDA_OK(cudaMemcpyAsync(d_logL, logL.data(), sizeof(float) * G * C, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));

  h->Bm = h->alloc<float>((size_t)N * C);
  h->vA = h->alloc<float>((size_t)N * C);
  h->s = h->alloc<float>(N);
  h->colsum = h->alloc<float>(G);
  double* cst = h->alloc<double>(N);
  if (compact_u8) CA_LAUNCH(k_setup_rows<uint8_t>, (unsigned)N, 256, 0, h->stream)(Yc, h->ldY, N, G, C, d_logL, h->s, cst, h->Bm);
  else CA_LAUNCH(k_setup_rows<float>, (unsigned)N, 256, 0, h->stream)(Yf, h->ldY, N, G, C, d_logL, h->s, cst, h->Bm);
  KCHECK();
  {
    double* csum = h->alloc<double>(1);
    CA_LAUNCH(k_reduce_partials, 1, 1024, 0, h->stream)(cst, N, 1, csum, 0.0);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(&h->const_sum, csum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(csum);
  }
  h->release(cst);
  h->release(d_logL);
  if (colsum_total) {
    std::vector<float> cs(G);
    for (int g = 0; g < G; ++g) cs[g] = (float)colsum_total[g];
    CUDA_OK(cudaMemcpyAsync(h->colsum, cs.data(), sizeof(float) * G, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  } else {
    // colSums(Y) (R/inference-tflow.R:117) over this shard in fp64; under cell sharding the shards' sums are added
    // with one all-reduce (collective: every rank passes colsum_total == NULL or none does)
    const int RS = 64;
    double* part = h->alloc<double>((size_t)RS * G);
    double* tot = h->alloc<double>(G);
    dim3 grid((G + 127) / 128, RS);
    if (compact_u8) CA_LAUNCH(k_colsum_part<uint8_t>, grid, 128, 0, h->stream)(Yc, h->ldY, N, G, RS, part);
    else CA_LAUNCH(k_colsum_part<float>, grid, 128, 0, h->stream)(Yf, h->ldY, N, G, RS, part);
    KCHECK();
    CA_LAUNCH(k_colsum_final, (G + 127) / 128, 128, 0, h->stream)(part, RS, G, h->colsum, tot);
    KCHECK();
    if (c.world > 1) {
      NCCL_OK(nccl().AllReduce(tot, tot, (size_t)G, kNcclFloat64, kNcclSum, h->comm, h->stream));
      CA_LAUNCH(k_colsum_final, (G + 127) / 128, 128, 0, h->stream)(tot, 1, G, h->colsum, nullptr);
      KCHECK();
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(part);
    h->release(tot);
  }
  if (c.V > 0) {
    int V = c.V;
    float* d_alt = h->alloc<float>((size_t)N * V);
    float* d_cov = h->alloc<float>((size_t)N * V);
    float* d_cn = h->alloc<float>((size_t)V * C);
    upload_colmajor(h, alt, N, V, d_alt, V, 0);
    upload_colmajor(h, cov, N, V, d_cov, V, 0);
    upload_colmajor(h, clone_allele, V, C, d_cn, C, 0);
    CA_LAUNCH(k_allele, (unsigned)N, 128, 0, h->stream)(d_alt, d_cov, d_cn, N, V, C, h->vA);
    KCHECK();
    h->snv = h->alloc<float>((size_t)N * C);
    CA_LAUNCH(k_softmax_rows, (unsigned)ceil_div64(N, 128), 128, 0, h->stream)(h->vA, N, C, h->snv);
    KCHECK();
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(d_alt);
    h->release(d_cov);
    h->release(d_cn);
  }
  // narrow Y after the setup passes that read it as fp32
  if (!compact_u8 && h->ystore != CA_STORE_F32) {
    dim3 grid(std::min<int64_t>((h->ldY + 255) / 256, 64), 1);
    void* Yn = nullptr;
    if (h->ystore == CA_STORE_U16) Yn = h->alloc<uint16_t>((size_t)N * h->ldY, false);
    else Yn = h->alloc<uint8_t>((size_t)N * h->ldY, false);
    for (int64_t r0 = 0; r0 < N; r0 += 65535) {
      grid.y = (unsigned)std::min<int64_t>(65535, N - r0);
      if (h->ystore == CA_STORE_U16) CA_LAUNCH(k_narrow_y<uint16_t>, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, (uint16_t*)Yn + r0 * h->ldY);
      else CA_LAUNCH(k_narrow_y<uint8_t>, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, (uint8_t*)Yn + r0 * h->ldY);
      KCHECK();
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(Yf);
    h->Y = Yn;
  }
  }   // !shared
  if (h->data_only) {
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return;
  }

  // ---- parameters (R/inference-tflow.R:240-272) ----
  h->dstate = h->alloc<StepState>(1);
  auto z = [&](size_t n) { return h->alloc<float>(n); };
  h->U = z((size_t)N * KP + 64); h->m_U = z((size_t)N * KP); h->v_U = z((size_t)N * KP); h->g_U = z((size_t)N * KP);
  h->Vm = z((size_t)G * KP + 64); h->m_V = z((size_t)G * KP); h->v_V = z((size_t)G * KP); h->g_V = z((size_t)G * KP);
  h->chi_raw = z(K); h->m_chi = z(K); h->v_chi = z(K); h->g_chi = z(K);
  h->u = z(C); h->m_u = z(C); h->v_u = z(C); h->g_u = z(C);
  h->loc = z(G); h->m_loc = z(G); h->v_loc = z(G); h->g_loc = z(G);
  h->lsd = z(G); h->m_lsd = z(G); h->v_lsd = z(G); h->g_lsd = z(G);
  h->t = z((size_t)N * C); h->m_t = z((size_t)N * C); h->v_t = z((size_t)N * C); h->g_t = z((size_t)N * C);
  if (K > 0) upload_colmajor(h, psi_init, N, K, h->U, KP, 0);
  if (c.P > 0) upload_colmajor(h, X, N, c.P, h->U, KP, K);
  {
    std::vector<float> lf(G);
    for (int g = 0; g < G; ++g) lf[g] = (float)loc_init[g];
    CUDA_OK(cudaMemcpyAsync(h->loc, lf.data(), sizeof(float) * G, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }

  // ---- scratch ----
  h->eps_in = z((size_t)S * G); h->eps = z((size_t)S * G); h->mu = z((size_t)S * G); h->logmu = z((size_t)S * G); h->sig = z((size_t)S * G);
  h->shift = z(N + 64); h->mm = z(2); h->log_alpha = z(C);
  h->YV = z((size_t)N * std::max(KP, 1)); h->YtU = z((size_t)G * std::max(KP, 1)); h->Fout = z((size_t)N * C);
  h->dM_sum = z((size_t)G * J);
  h->n_gene_blocks = (int)ceil_div64((int64_t)G * S, h->lean ? kProThreads : 256);   // one thread per (sample, gene) pair
  // fused prologue: two 512-thread blocks fit an SM; 32 of the slots go to its scalar / range blocks, the gene blocks stride
  if (h->lean) h->n_gene_blocks = std::min(h->n_gene_blocks, std::max(1, 2 * h->num_sms - 2 - kProPsiBlocks));
  h->n_epi_blocks = ceil_div64(N, kEpiWarps);
  h->n_cell_parts = h->epi2 ? (int64_t)h->num_sms : h->n_epi_blocks;
  h->gene_part = h->alloc<double>(h->n_gene_blocks);
  h->n_yv_blocks = h->defer ? h->num_sms : 0;
  h->elbo_part = h->alloc<double>(h->n_cell_parts + h->n_yv_blocks);
  h->gsum_part = h->alloc<double>((size_t)h->n_cell_parts * C);
  h->scal_elbo = h->alloc<double>(1); h->cell_sum = h->alloc<double>(1); h->wsq = h->alloc<double>(std::max(K, 1));
  h->elbo_dev = h->alloc<double>(1);
  h->ar = z((size_t)G * (2 + KP) + C + 4);
  h->ypass5 = kImmaAvailable && (h->variants & CA_VAR_YPASS5) && (h->variants & CA_VAR_YPASS4) && h->ystore == CA_STORE_U8 && KP == 1;
  if (KP == 1) {
    int tile_cols = kYCB;
    if (h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4))   // column tile of k_ypass_k1_v3 / v4: 256 threads x the columns a thread owns for this storage type
      tile_cols = h->ystore == CA_STORE_U8 ? ypass3_tile_cols<uint8_t>() : (h->ystore == CA_STORE_U16 ? ypass3_tile_cols<uint16_t>() : ypass3_tile_cols<float>());
    h->y5_spec = h->ypass5 && (h->y7_auto || getenv("CLONEALIGN_B200_Y5_SPEC") != nullptr);
    h->y7 = h->y5_spec && kY7Available && (h->y7_auto || atoi(getenv("CLONEALIGN_B200_Y5_SPEC")) == 2);
    if (h->ypass5) tile_cols = h->y5_spec ? kY6Cols : kY5Cols;
    h->nCB = (int)ceil_div64(h->ldY, tile_cols);
    h->RB = 512;
    if (h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4)) {
      // Size the row blocks so that the grid is (just under) a whole number of waves of the 2 CTAs an SM holds: with
      // 512-row blocks config 3 gives 5 x 196 = 980 CTAs = 3.31 waves of 296, i.e. a last wave that is one third full
      // on the kernel that bounds the step; 432-row blocks give 5 x 232 = 1160 CTAs = 3.92 waves.  Small problems get
      // enough row blocks to cover every SM (10k x 5k: 250 CTAs instead of 40).  Any multiple of 16 rows works (vector
      // loads of psi, 8 / 16 rows in flight).
      // The persistent pass (ypass4) walks its tiles itself: two tiles per CTA are enough to balance the grid, and longer
      // row blocks halve the column partials the gene kernel has to gather (and the pipeline refills at tile starts).
      const int64_t slots = 2 * (int64_t)h->num_sms;
      int64_t waves = std::max<int64_t>(1, ceil_div64((int64_t)h->nCB * ceil_div64(N, 512), slots));
      if (h->variants & CA_VAR_YPASS4) waves = std::min<int64_t>(waves, h->cell2 ? 1 : 2);   // cell2: half the column partials to add behind the join
      const int64_t nrb = std::max<int64_t>(1, waves * slots / h->nCB);
      h->RB = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(16, round_up64(ceil_div64(N, nrb), 16)));
    }
    if (h->ypass5) {
      // one persistent CTA per SM; a whole number of waves of (kY5Cols columns x RB rows) tiles, RB a multiple of the 32-row stage
      // and at most kY5MaxRows (the psi digits of a tile live in shared memory)
      const int64_t slots = h->num_sms;
      const int64_t waves = std::max<int64_t>(1, ceil_div64((int64_t)h->nCB * ceil_div64(N, kY5MaxRows), slots));
      const int64_t nrb = std::max<int64_t>(1, waves * slots / h->nCB);
      h->RB = (int)std::min<int64_t>(kY5MaxRows, std::max<int64_t>(kY5StageRows, round_up64(ceil_div64(N, nrb), kY5StageRows)));
      if (h->y7) y7_plan_tiles(h->y7plan, N, h->ldY, h->num_sms, h->RB);   // per-CTA tile lists
    }
  } else {
    h->nCB = 1;
    h->RB = 1024;
  }
  h->nRB = (int)ceil_div64(N, h->RB);
  h->rowpart = z((size_t)h->nCB * N * std::max(KP, 1));
  h->colpart = z((size_t)h->nRB * G * std::max(KP, 1));
  if (h->tc) {
    h->MxT_hi = h->alloc<__nv_bfloat16>((size_t)J * h->Gld);
    h->MxT_lo = h->alloc<__nv_bfloat16>((size_t)J * h->Gld);
    h->RxT = h->alloc<__half>((size_t)J * h->Nld);
    h->shift_bwd = z((size_t)h->Nld);
    tc_plan_create(h->tcplan, h->dev, N, h->Nld, G, h->Gld, h->SCp, J, h->MxT_hi, h->MxT_lo, h->RxT);
    // The overlapped Y pass must be able to share an SM with a contraction CTA (211 KB of shared memory): give it
    // the same (maximum) shared-memory carveout, otherwise the SM has to drain before it can be reconfigured.
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->nsplit = h->tcplan.nsplit;
    h->Zx = z((size_t)h->tcplan.fsplit * N * J);
    h->dMx = z((size_t)h->nsplit * G * J);
  } else {
    h->Mx = z((size_t)G * J);
    h->Zx = z((size_t)N * J);
    h->Rx = z((size_t)N * J);
    h->dMx = z((size_t)G * J);
    h->nsplit = 1;
  }
  if (h->interp) {
    h->iplan = h->alloc<InterpPlan>(1);
    h->mm_psi = z(2);
    // node-sum kernel: columns per thread, column groups per warp, slices of the reduction index (>= 4 staged chunks per
    // work item, at most three items per SM and panel: with one active panel every resident block still has work), dynamic shared memory
    h->Jn = h->cell2 ? (int)round_up64(h->SC, 2) : J;   // even: 8-byte row starts (cp.async, float2 reads of the node kernel)
    h->n2_tj = n2_pick_tj(h->Jn);
    h->n2_ncgp = n2_ncg_pow2(h->Jn, h->n2_tj);
    auto n2_split = [&](int64_t R) {
      return (int)std::max<int64_t>(1, std::min<int64_t>(kN2BlocksPerSM * (int64_t)h->num_sms, ceil_div64(ceil_div64(R, kN2Chunk), 4)));
    };
    h->n2_split_f = n2_split(G);
    h->n2_split_b = n2_split(N);
    h->n2_smem = n2_smem_bytes(h->Jn, h->n2_tj);
    h->n2_blocks_per_sm = (int)std::max<size_t>(1, std::min<size_t>(kN2BlocksPerSM, (220 * 1024) / h->n2_smem));
    if (h->n2_smem > 48 * 1024) {
      if (h->n2_tj == 8) {
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
      } else {
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
      }
    }
    const size_t nodes_f = (size_t)h->n2_split_f * kIMaxPanF * kIP, nodes_b = (size_t)h->n2_split_b * kIMaxPanB * kIP;
    h->ivals = h->alloc<double>(std::max(nodes_f, nodes_b) * J, false);
    h->icoef = h->alloc<double>((size_t)std::max(kIMaxPanF, kIMaxPanB) * kIP * J);
    if (h->cell2) {
      h->icoef2 = h->alloc<double>((size_t)std::max(kIMaxPanF, kIMaxPanB) * kIP * J);
      h->ipart2 = h->alloc<double>((size_t)kC3Groups * std::max(kIMaxPanF, kIMaxPanB) * kIP * J, false);
      h->itickets = h->alloc<unsigned>((size_t)(J + kC2Cols - 1) / kC2Cols);
    }
    const size_t per_panel = (size_t)kIP * J * sizeof(double);
    h->ieval_panels = (int)std::min<size_t>(16, (200 * 1024) / per_panel);
    h->ieval_smem = (size_t)h->ieval_panels * per_panel;
    if (h->ieval_smem > 48 * 1024) {
      CUDA_OK(cudaFuncSetAttribute(k_interp_eval<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ieval_smem));
      CUDA_OK(cudaFuncSetAttribute(k_interp_eval<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ieval_smem));
    }
  }
  if (h->ypass5) {
    if (const char* e = getenv("CLONEALIGN_B200_Y5_WARPS")) h->y5_warps = atoi(e) == 16 ? 16 : 8;
    if (h->y7) {
      y7_plan_create(h->y7plan, h->Y, N, h->ldY);
      int* dt = h->alloc<int>(h->y7plan.tiles.size());
      int* dof = h->alloc<int>(h->y7plan.offs.size());
      CUDA_OK(cudaMemcpyAsync(dt, h->y7plan.tiles.data(), sizeof(int) * h->y7plan.tiles.size(), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(cudaMemcpyAsync(dof, h->y7plan.offs.data(), sizeof(int) * h->y7plan.offs.size(), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      h->y7plan.d_tiles = dt; h->y7plan.d_offs = dof;
      CUDA_OK(y7_set_attributes());
    }
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ypass6_smem_bytes()));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v6, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v5<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ypass5_smem_bytes()));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v5<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v5<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ypass5_smem_bytes()));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v5<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  if (h->variants & CA_VAR_YPASS4) {
    if (const char* e = getenv("CLONEALIGN_B200_Y4_MINB")) h->y4_minb = atoi(e) == 3 ? 3 : 4;
    auto set4 = [&](auto kern, size_t bytes) { CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); };
    set4(k_ypass_k1_v4<float, 3>, ypass4_smem_bytes<float>()); set4(k_ypass_k1_v4<float, 4>, ypass4_smem_bytes<float>());
    set4(k_ypass_k1_v4<uint16_t, 3>, ypass4_smem_bytes<uint16_t>()); set4(k_ypass_k1_v4<uint16_t, 4>, ypass4_smem_bytes<uint16_t>());
    set4(k_ypass_k1_v4<uint8_t, 3>, ypass4_smem_bytes<uint8_t>()); set4(k_ypass_k1_v4<uint8_t, 4>, ypass4_smem_bytes<uint8_t>());
  }
  if (h->lean) {
    h->chi_cur = h->alloc<double>(std::max(K, 1));
    h->pmm_part = h->alloc<float>(2 * kProPsiBlocks);
    h->ticket = h->alloc<unsigned>(1);
    h->gene_panels = gene_fused_smem_panels(J);
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS")) h->gene_panels = std::max(0, std::min(h->gene_panels, atoi(e)));
    h->gene_smem = gene_fused_smem_bytes(J, h->gene_panels);
    if (h->gene_smem > 48 * 1024)
    {
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
    }
  }
  if (h->epi2) {
    h->fused_nj = (h->SC + 31) / 32;
    // defer + overlap: 16 warps x 64 registers = half of the register file, so that one Y-pass CTA (256 threads x 128
    // registers, the other half) can be resident on the same SM while the per-cell kernel runs
    h->fused_warps = (h->defer && (c.variants & (CA_VAR_OVERLAP | CA_VAR_COSCHED))) ? kFusedWarps / 2 : kFusedWarps;
    if (h->cosched && h->y4_minb == 3) h->fused_warps = 12;     // 2 x 256 x 80 registers for the stream leave 24 K of the 64 K
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_WARPS")) h->fused_warps = std::max(1, std::min(kFusedWarps, atoi(e)));
    if (h->fused_warps != kFusedWarps) {
      // the Y-pass CTA must fit next to ~200 KB of shared memory: ask for the maximum shared-memory carveout, otherwise the
      // SM would have to drain before it can be reconfigured (measured in round 1 for the contraction kernels)
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    // cosched: two persistent Y-pass CTAs (64 KB rings) stay resident on every SM; the per-cell CTA gets what is left
    const size_t fused_budget = h->cosched ? 92 * 1024 : 200 * 1024;
    h->fused_panels = fused_smem_panels(h->SC, C, J, fused_budget, h->fused_warps);
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS"))   // test hook: force the coefficients-through-L2 branch
      h->fused_panels = std::max(0, std::min(h->fused_panels, atoi(e)));
    h->fused_smem = fused_smem_bytes(h->SC, C, J, h->fused_panels, h->fused_warps);
    if (h->fused_smem > 48 * 1024) {
      switch (h->fused_nj) {
        case 1: fused_set_smem<1>(h->fused_smem); break;
        case 2: fused_set_smem<2>(h->fused_smem); break;
        case 3: fused_set_smem<3>(h->fused_smem); break;
        default: fused_set_smem<4>(h->fused_smem); break;
      }
    }
  }
  if (h->cell2) {
    h->cell2_wc = cell2_pick_wc(C);
    h->cell2_sb = cell2_pick_sb(S);
    const size_t budget = h->cosched ? (h->ypass5 ? (h->y5_spec ? 48 * 1024 : 56 * 1024) : 92 * 1024) : 200 * 1024;
    h->cell2_panels = cell2_smem_panels(h->cell2_wc, h->cell2_sb, C, budget, h->fused_warps);
    if (h->cell2_panels < 1) fail("variant cell2: the coefficient table of one panel (%zu bytes) does not fit into shared memory", cell2_panel_bytes(h->cell2_wc, h->cell2_sb));
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS")) h->cell2_panels = std::max(1, std::min(h->cell2_panels, atoi(e)));   // test hook: several rounds
    h->cell2_smem = cell2_smem_bytes(h->cell2_wc, h->cell2_sb, C, h->cell2_panels, h->fused_warps);
    h->gene2_panels = std::max(1, gene2_smem_panels(h->cell2_wc, h->cell2_sb, (h->cosched && h->ypass5) ? 48 * 1024 : 96 * 1024));   // two 512-thread blocks per SM (one next to the Y pass)
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS")) h->gene2_panels = std::max(1, std::min(h->gene2_panels, atoi(e)));
    h->gene2_smem = gene2_smem_bytes(h->cell2_wc, h->cell2_sb, h->gene2_panels);
    cell2_dispatch(h->cell2_wc, h->cell2_sb, [&](auto wc, auto sb) {
      constexpr int WC = decltype(wc)::value, SB = decltype(sb)::value;
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused2<WC, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene2_smem));
      CUDA_OK(cudaFuncSetAttribute(k_cell_fused2<EPI_TRAIN, WC, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cell2_smem));
      CUDA_OK(cudaFuncSetAttribute(k_cell_fused2<EPI_EVAL, WC, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cell2_smem));
      CUDA_OK(cudaFuncSetAttribute(k_cell_fused2<EPI_INIT, WC, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cell2_smem));
    });
  }
  size_t smem = epi_smem_bytes(h->SCp, C, J, h->tc);
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) fail("S*C too large for the per-cell epilogue (%zu bytes of shared memory)", smem);
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (h->p2p) {
    h->p2p_cnt = (int64_t)G * (2 + KP) + C;
    h->p2p_cnt_pad = round_up64(h->p2p_cnt, 4);
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)c.world * h->p2p_cnt_pad;
    // one allocation (one IPC handle): slots, then the flags on their own 256-byte line
    h->p2p_buf = (float*)h->alloc<unsigned char>(slot_bytes + 256 + sizeof(unsigned) * 2 * kP2PMaxWorld);
    h->p2p_ticket = h->alloc<unsigned>(1);
    h->p2p_err = h->alloc<int>(1);
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
}

struct ArrayRef {
  const float* p;
  int64_t rows;
  int cols, ld, off;
  bool writable;
};

bool lookup(ca_handle* h, const std::string& n, ArrayRef& r) {
  const int64_t N = h->N;
  const int G = h->G, C = h->C, K = h->K, P = h->P, KP = h->KP, SC = h->SC, J = h->J;
  auto set = [&](const float* p, int64_t rows, int cols, int ld, int off, bool w) { r = {p, rows, cols, ld, off, w}; return true; };
  if (n == "W") return set(h->Vm, G, K, KP, 0, true);
  if (n == "beta") return set(h->Vm, G, P, KP, K, true);
  if (n == "psi") return set(h->U, N, K, KP, 0, true);
  if (n == "chi_raw") return set(h->chi_raw, K, 1, 1, 0, true);
  if (n == "alpha_unconstr") return set(h->u, C, 1, 1, 0, true);
  if (n == "loc") return set(h->loc, G, 1, 1, 0, true);
  if (n == "lsd") return set(h->lsd, G, 1, 1, 0, true);
  if (n == "gamma_logits") return set(h->t, N, C, C, 0, true);
  if (n == "grad_W") return set(h->g_V, G, K, KP, 0, false);
  if (n == "grad_beta") return set(h->g_V, G, P, KP, K, false);
  if (n == "grad_psi") return set(h->g_U, N, K, KP, 0, false);
  if (n == "grad_chi_raw") return set(h->g_chi, K, 1, 1, 0, false);
  if (n == "grad_alpha_unconstr") return set(h->g_u, C, 1, 1, 0, false);
  if (n == "grad_loc") return set(h->g_loc, G, 1, 1, 0, false);
  if (n == "grad_lsd") return set(h->g_lsd, G, 1, 1, 0, false);
  if (n == "grad_gamma_logits") return set(h->g_t, N, C, C, 0, false);
  if (n == "Z") return set(h->Zx, N, SC, J, 0, false);
  if (n == "Zx") return set(h->Zx, N, J, J, 0, false);
  if (n == "R" && h->Rx) return set(h->Rx, N, SC, h->cell2 ? h->Jn : J, 0, false);
  if (n == "dM") return set(h->dM_sum, G, SC, J, 0, false);
  if (n == "dMx") return set(h->dM_sum, G, J, J, 0, false);
  if (n == "F") return set(h->Fout, N, C, C, 0, false);
  if (n == "YV") return set(h->YV, N, KP, KP, 0, false);
  if (n == "YtU") return set(h->YtU, G, KP, KP, 0, false);
  if (n == "B") return set(h->Bm, N, C, C, 0, false);
  if (n == "v") return set(h->vA, N, C, C, 0, false);
  if (n == "s") return set(h->s, N, 1, 1, 0, false);
  if (n == "colsum") return set(h->colsum, G, 1, 1, 0, false);
  if (n == "shift") return set(h->shift, N, 1, 1, 0, false);
  if (n == "mu_samples") return set(h->mu, h->S, G, G, 0, false);   // NOTE: returned as S x G column-major
  return false;
}

}  // namespace
