// clonealign_b200 core: device state + the C-ABI declared in include/clonealign_b200.h.
//
// One ca_handle == one TensorFlow session of the reference (R/inference-tflow.R:351-457) for one
// cell shard on one GPU.  No CPU fallback exists: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

#include "../../include/clonealign_b200.h"
#include "common.cuh"
#include "kernels_expgemm.cuh"
#include "kernels_cell.cuh"
#include "kernels_fused.cuh"
#include "kernels_interp.cuh"
#include "kernels_p2p.cuh"
#include "kernels_pca.cuh"
#include "kernels_small.cuh"
#include CA_TC_HEADER   // platform.cuh: kernels_tc.cuh
#include "kernels_ypass.cuh"
#include CA_Y7_HEADER   // platform.cuh: kernels_ypass_tma.cuh

using namespace ca;

#include "core_support.inl"
#include "core_state.inl"
#include "core_step.inl"
#include "core_build.inl"

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int ca_core_abi_version(void) { return CA_ABI_VERSION; }

int ca_core_shutdown(void) {
  std::lock_guard<std::mutex> lk(comm_mu());
  for (auto& kv : comm_pool())
    for (void* c : kv.second) {
      cudaSetDevice(kv.first.dev);
      nccl().CommDestroy(c);
    }
  comm_pool().clear();
  return 0;
}

int ca_core_device_count(int* count, char* err, size_t errlen) {
  try {
    int n = 0;
    CUDA_OK(cudaGetDeviceCount(&n));
    if (count) *count = n;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_nccl_unique_id(void* out128, char* err, size_t errlen) {
  try {
    if (!out128) fail("null output");
    NCCL_OK(nccl().GetUniqueId(out128));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_create(ca_handle** out, const ca_config* cfg, const void* Y, const double* L, const double* psi_init,
                   const double* loc_init, const double* X, const double* colsum_total, const double* clone_allele,
                   const double* alt, const double* cov, char* err, size_t errlen) {
  ca_handle* h = nullptr;
  try {
    if (!out || !cfg) fail("null argument");
    h = new ca_handle();
    h->cfg = *cfg;
    build(h, Y, L, psi_init, loc_init, X, colsum_total, clone_allele, alt, cov);
    h->cfg.nccl_id = nullptr;   // never retain caller pointers
    h->cfg.y_indptr = nullptr;
    h->cfg.y_indices = nullptr;
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    destroy(h);
    return report(e, err, errlen);
  }
}

int ca_core_data_create(ca_data** out, const ca_config* cfg, const void* Y, const double* L, const double* colsum_total,
                        const double* clone_allele, const double* alt, const double* cov, char* err, size_t errlen) {
  ca_handle* t = nullptr;
  try {
    if (!out || !cfg) fail("null argument");
    t = new ca_handle();
    t->cfg = *cfg;
    t->cfg.S = std::max(1, cfg->S);
    t->cfg.K = 0; t->cfg.P = 0; t->cfg.path = CA_PATH_CUDACORE; t->cfg.variants = 0; t->cfg.world = 1; t->cfg.rank = 0;
    t->data_only = true;
    build(t, Y, L, nullptr, nullptr, nullptr, colsum_total, clone_allele, alt, cov);
    ca_data* d = new ca_data();
    d->dev = t->dev; d->N = t->N; d->ldY = t->ldY; d->G = t->G; d->C = t->C; d->V = t->V; d->ystore = t->ystore;
    d->poison = t->poison; d->const_sum = t->const_sum;
    d->Y = t->Y; d->L = t->L; d->Bm = t->Bm; d->vA = t->vA; d->s = t->s; d->colsum = t->colsum; d->snv = t->snv;
    for (void* p : t->allocs)
      if (p) d->allocs.push_back(p);
    t->allocs.clear();     // ownership moved
    destroy(t);
    *out = d;
    return 0;
  } catch (const std::exception& e) {
    destroy(t);
    return report(e, err, errlen);
  }
}

int ca_core_data_destroy(ca_data* d, char* err, size_t errlen) {
  try {
    if (!d) return 0;
    if (d->refs.load() > 0) fail("ca_core_data_destroy: %d session(s) still use these inputs", d->refs.load());
    cudaSetDevice(d->dev);
    for (void* p : d->allocs) cudaFree(p);
    delete d;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_data_stats(ca_data* d, double* rowsum, double* colsum, double* mu_guess, char* err, size_t errlen) {
  try {
    if (!d) fail("null argument");
    CUDA_OK(cudaSetDevice(d->dev));
    const int64_t N = d->N;
    const int G = d->G;
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    double *d_row = nullptr, *d_part = nullptr, *d_sums = nullptr;
    struct Guard {
      double*& a; double*& b; double*& c;
      ~Guard() { cudaFree(a); cudaFree(b); cudaFree(c); }
    } guard{d_row, d_part, d_sums};
    CUDA_OK(cudaMalloc(&d_row, sizeof(double) * N));
    CUDA_OK(cudaMalloc(&d_part, sizeof(double) * (size_t)RS * G * 2));
    CUDA_OK(cudaMalloc(&d_sums, sizeof(double) * (size_t)G * 2));
    cudaStream_t st = nullptr;   // the legacy default stream: the inputs are immutable and nothing else is in flight on them
    auto run = [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      CA_LAUNCH(k_stats_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, st)(Yp, d->ldY, N, G, d_row);
      KCHECK();
      CA_LAUNCH(k_stats_cols<T>, dim3((G + 127) / 128, RS), 128, 0, st)(Yp, d->ldY, N, G, RS, d_row, d_part);
      KCHECK();
    };
    switch (d->ystore) {
      case CA_STORE_F32: run((const float*)d->Y); break;
      case CA_STORE_U16: run((const uint16_t*)d->Y); break;
      case CA_STORE_U8: run((const uint8_t*)d->Y); break;
      default: fail("bad y_store");
    }
    CA_LAUNCH(k_pca_colstats_reduce, (G + 127) / 128, 128, 0, st)(d_part, RS, G, d_sums);
    KCHECK();
    std::vector<double> hs((size_t)2 * G);
    CUDA_OK(cudaMemcpyAsync(hs.data(), d_sums, sizeof(double) * 2 * G, cudaMemcpyDeviceToHost, st));
    if (rowsum) CUDA_OK(cudaMemcpyAsync(rowsum, d_row, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int g = 0; g < G; ++g) {
      if (colsum) colsum[g] = hs[2 * (size_t)g];
      if (mu_guess) mu_guess[g] = hs[2 * (size_t)g + 1] * (double)G / (double)N;   // colMeans(Y / rowMeans(Y)), :222
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// rowSums(Y[, keep]) over the resident matrix: the cell filter of preprocess_for_clonealign (R/preprocess.R:138-139), which
// counts only the genes that survived the gene filters.  gene_keep: G bytes, non-zero = gene retained.
int ca_core_data_masked_rowsums(ca_data* d, const uint8_t* gene_keep, double* rowsum, char* err, size_t errlen) {
  try {
    if (!d || !gene_keep || !rowsum) fail("null argument");
    CUDA_OK(cudaSetDevice(d->dev));
    const int64_t N = d->N;
    const int G = d->G;
    double* d_row = nullptr;
    unsigned char* d_keep = nullptr;
    struct Guard {
      double*& a; unsigned char*& b;
      ~Guard() { cudaFree(a); cudaFree(b); }
    } guard{d_row, d_keep};
    CUDA_OK(cudaMalloc(&d_row, sizeof(double) * N));
    CUDA_OK(cudaMalloc(&d_keep, (size_t)G));
    cudaStream_t st = nullptr;
    CUDA_OK(cudaMemcpyAsync(d_keep, gene_keep, (size_t)G, cudaMemcpyHostToDevice, st));
    auto run = [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      CA_LAUNCH(k_stats_rows_masked<T>, (unsigned)ceil_div64(N, 8), 256, 0, st)(Yp, d->ldY, N, G, d_keep, d_row);
      KCHECK();
    };
    switch (d->ystore) {
      case CA_STORE_F32: run((const float*)d->Y); break;
      case CA_STORE_U16: run((const uint16_t*)d->Y); break;
      case CA_STORE_U8: run((const uint8_t*)d->Y); break;
      default: fail("bad y_store");
    }
    CUDA_OK(cudaMemcpyAsync(rowsum, d_row, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_create_shared(ca_handle** out, const ca_config* cfg, ca_data* data, const double* psi_init, const double* loc_init,
                          const double* X, char* err, size_t errlen) {
  ca_handle* h = nullptr;
  try {
    if (!out || !cfg || !data) fail("null argument");
    h = new ca_handle();
    h->cfg = *cfg;
    h->shared = data;
    data->refs++;
    build(h, nullptr, nullptr, psi_init, loc_init, X, nullptr, nullptr, nullptr, nullptr);
    h->cfg.nccl_id = nullptr;
    h->cfg.y_indptr = nullptr;
    h->cfg.y_indices = nullptr;
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    destroy(h);
    return report(e, err, errlen);
  }
}

int ca_core_destroy(ca_handle* h) {
  destroy(h);
  return 0;
}

int ca_core_init_gamma(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    run_forward(h, EPI_INIT);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_step(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    train_step(h);
    return 0;   // asynchronous: the next call on this handle is stream-ordered behind it
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_grads(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    h->inspect = true;
    run_train(h, false);
    h->inspect = false;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_elbo(ca_handle* h, double* elbo, char* err, size_t errlen) {
  try {
    if (!h || !elbo) fail("null argument");
    CUDA_OK(cudaSetDevice(h->dev));
    eval_step(h);
    CUDA_OK(cudaMemcpyAsync(elbo, h->elbo_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    int p2p_failed = 0;
    if (h->p2p_err) CUDA_OK(cudaMemcpyAsync(&p2p_failed, h->p2p_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (p2p_failed) fail("variant p2p: the all-reduce kernel timed out waiting for a peer's contribution");
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// n evaluations of the ELBO with fresh draws, queued back to back on the stream and fetched with ONE device-to-host
// copy (the 20 evaluations behind final_elbo / sd_final_elbo, R/inference-tflow.R:447-449, each a sess$run(elbo) with
// its own host round trip in the reference).  Same draw sequence, same values as n calls of ca_core_elbo; the
// parameters do not change in between, so only the first evaluation can need a Y pass.
int ca_core_elbo_many(ca_handle* h, int32_t n, double* elbo, char* err, size_t errlen) {
  double* d_out = nullptr;
  try {
    if (!h || !elbo || n < 0) fail("bad argument");
    if (n == 0) return 0;
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * (size_t)n));
    for (int i = 0; i < n; ++i) {
      eval_step(h);
      CUDA_OK(cudaMemcpyAsync(d_out + i, h->elbo_dev, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_OK(cudaMemcpyAsync(elbo, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    int p2p_failed = 0;
    if (h->p2p_err) CUDA_OK(cudaMemcpyAsync(&p2p_failed, h->p2p_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d_out);
    d_out = nullptr;
    if (p2p_failed) fail("variant p2p: the all-reduce kernel timed out waiting for a peer's contribution");
    return 0;
  } catch (const std::exception& e) {
    if (d_out) { cudaStreamSynchronize(h->stream); cudaFree(d_out); }
    return report(e, err, errlen);
  }
}

int ca_core_params(ca_handle* h, double* mu, double* clone_probs, double* s, double* alpha, double* psi, double* W,
                   double* chi, double* beta, double* clone_probs_from_snv, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    const int64_t N = h->N;
    const int G = h->G, C = h->C, K = h->K, KP = h->KP;
    if (mu) {   // tf$nn$softplus(qmu$distribution$loc), R/inference-tflow.R:424
      std::vector<double> tmp(G);
      download_colmajor(h, h->loc, G, 1, 1, 0, tmp.data());
      for (int g = 0; g < G; ++g) { double x = tmp[g]; mu[g] = x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
    }
    if (clone_probs) {   // softmax(gamma_logits), :273,424 -- on the device (2.4 M host exp() calls took longer than the whole fit loop of the e2e bench)
      if (!h->cp_scratch) h->cp_scratch = h->alloc<double>((size_t)N * C, false);   // kept for the session (no allocation / free per call)
      double* d_cp = h->cp_scratch;
      CA_LAUNCH(k_softmax_rows_f64, (unsigned)ceil_div64(N, 128), 128, 0, h->stream)(h->t, N, C, d_cp);
      KCHECK();
      std::vector<double> tmp((size_t)N * C);
      CUDA_OK(cudaMemcpyAsync(tmp.data(), d_cp, sizeof(double) * (size_t)N * C, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      for (int c = 0; c < C; ++c)
        for (int64_t n = 0; n < N; ++n) clone_probs[(size_t)c * N + n] = tmp[(size_t)n * C + c];
    }
    if (s) download_colmajor(h, h->s, N, 1, 1, 0, s);
    if (alpha) {   // exp(log_softmax(alpha_unconstr))
      std::vector<double> tmp(C);
      download_colmajor(h, h->u, C, 1, 1, 0, tmp.data());
      double mx = -1e300, z = 0;
      for (int c = 0; c < C; ++c) mx = std::max(mx, tmp[c]);
      for (int c = 0; c < C; ++c) z += exp(tmp[c] - mx);
      for (int c = 0; c < C; ++c) alpha[c] = exp(tmp[c] - mx) / z;
    }
    if (psi && K > 0) download_colmajor(h, h->U, N, K, KP, 0, psi);
    if (W && K > 0) download_colmajor(h, h->Vm, G, K, KP, 0, W);
    if (beta && h->P > 0) download_colmajor(h, h->Vm, G, h->P, KP, K, beta);
    if (chi && K > 0) {
      std::vector<double> tmp(K);
      download_colmajor(h, h->chi_raw, K, 1, 1, 0, tmp.data());
      for (int k = 0; k < K; ++k) chi[k] = exp(tmp[k]);
    }
    if (clone_probs_from_snv) {
      if (!h->snv) fail("clone_probs_from_snv requested but the allele-specific likelihood is not in use");
      download_colmajor(h, h->snv, N, C, C, 0, clone_probs_from_snv);
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_set_eps(ca_handle* h, const float* eps, int64_t n_draws, char* err, size_t errlen) {
  try {
    if (!h || !eps || n_draws < 0) fail("bad argument");
    size_t per = (size_t)h->S * h->G;
    h->eps_queue.insert(h->eps_queue.end(), eps, eps + per * (size_t)n_draws);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_get_eps(ca_handle* h, float* eps, char* err, size_t errlen) {
  try {
    if (!h || !eps) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaMemcpyAsync(eps, h->eps, sizeof(float) * h->S * h->G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_get_array(ca_handle* h, const char* name, double* out, int64_t n, char* err, size_t errlen) {
  try {
    if (!h || !name || !out) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    ArrayRef r;
    if (!lookup(h, name, r)) fail("unknown array '%s'", name);
    if (n < r.rows * r.cols) fail("buffer too small for '%s': need %lld", name, (long long)(r.rows * r.cols));
    download_colmajor(h, r.p, r.rows, r.cols, r.ld, r.off, out);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_set_array(ca_handle* h, const char* name, const double* in, int64_t n, char* err, size_t errlen) {
  try {
    if (!h || !name || !in) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    ArrayRef r;
    if (!lookup(h, name, r) || !r.writable) fail("array '%s' is not writable", name);
    if (n != r.rows * r.cols) fail("size mismatch for '%s': expected %lld", name, (long long)(r.rows * r.cols));
    upload_colmajor(h, in, r.rows, r.cols, const_cast<float*>(r.p), r.ld, r.off);
    h->ydirty = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_time_steps(ca_handle* h, int32_t n_steps, int32_t with_eval, double* ms, char* err, size_t errlen) {
  try {
    if (!h || !ms || n_steps < 0) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    cudaEvent_t a, b;
    CUDA_OK(cudaEventCreate(&a));
    CUDA_OK(cudaEventCreate(&b));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaEventRecord(a, h->stream));
    for (int i = 0; i < n_steps; ++i) {
      train_step(h);
      if (with_eval) eval_step(h);
    }
    CUDA_OK(cudaEventRecord(b, h->stream));
    CUDA_OK(cudaEventSynchronize(b));
    float f = 0.f;
    CUDA_OK(cudaEventElapsedTime(&f, a, b));
    *ms = (double)f;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_profile_step(ca_handle* h, char* names, size_t names_len, double* ms, int32_t cap, int32_t* n_k, char* err,
                         size_t errlen) {
  try {
    if (!h || !names || !ms || !n_k) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    for (auto& p : h->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    h->prof.clear();
    h->prof_overlap = getenv("CLONEALIGN_B200_PROF_OVERLAP") != nullptr;
    h->prof_on = true;
    run_train(h, true);
    h->prof_on = false;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream2));
    std::string all;
    int k = 0;
    for (auto& p : h->prof) {
      if (k >= cap) break;
      float f = 0.f;
      CUDA_OK(cudaEventElapsedTime(&f, p.a, p.b));
      ms[k++] = (double)f;
      if (!all.empty()) all += ";";
      all += p.name;
    }
    if (h->prof_overlap && !h->prof.empty()) {   // timeline: start of every launch relative to the first event of the step
      cudaEvent_t t0 = h->prof[0].a;
      for (auto& p : h->prof) {
        float f = 0.f;
        if (cudaEventElapsedTime(&f, t0, p.a) != cudaSuccess) { cudaGetLastError(); f = 0.f; }   // (the forked Y pass may start first)
        if (k >= cap) break;
        ms[k++] = (double)f;
        all += ";t0:" + p.name;
      }
    }
    *n_k = k;
    strncpy(names, all.c_str(), names_len - 1);
    names[names_len - 1] = 0;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_correlations(ca_handle* h, const int32_t* clone_idx, const double* L, double* out, char* err, size_t errlen) {
  try {
    if (!h || !clone_idx || !out) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    const int64_t N = h->N;
    const int G = h->G, C = h->C;
    if ((size_t)C * 128 * sizeof(float) > 48 * 1024) fail("too many clones for the correlation kernel");
    int* d_z = nullptr;
    float* d_L = nullptr;
    double *d_part = nullptr, *d_out = nullptr, *d_sums = nullptr;
    struct Guard {   // scratch is released on every exit path
      int*& z; float*& l; double*& p; double*& o; double*& s;
      ~Guard() { cudaFree(z); cudaFree(l); cudaFree(p); cudaFree(o); cudaFree(s); }
    } guard{d_z, d_L, d_part, d_out, d_sums};
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    CUDA_OK(cudaMalloc(&d_z, sizeof(int) * N));
    CUDA_OK(cudaMalloc(&d_L, sizeof(float) * (size_t)G * C));
    CUDA_OK(cudaMalloc(&d_part, sizeof(double) * (size_t)RS * G * 5));
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * G));
    CUDA_OK(cudaMalloc(&d_sums, sizeof(double) * ((size_t)5 * G + 1 + C)));
    std::vector<double> tail(1 + (size_t)C, 0.0);      // number of assigned cells, then cells per clone (this shard)
    for (int64_t n = 0; n < N; ++n)
      if (clone_idx[n] >= 0 && clone_idx[n] < C) { tail[0] += 1.0; tail[1 + clone_idx[n]] += 1.0; }
    CUDA_OK(cudaMemcpyAsync(d_z, clone_idx, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
    if (L) upload_colmajor(h, L, G, C, d_L, C, 0);
    else CUDA_OK(cudaMemcpyAsync(d_L, h->L, sizeof(float) * (size_t)G * C, cudaMemcpyDeviceToDevice, h->stream));
    dispatch_y(h, [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      dim3 grid((G + 127) / 128, RS);
      CA_LAUNCH(k_corr_part<T>, grid, 128, sizeof(float) * C * 128, h->stream)(Yp, h->ldY, N, G, C, d_z, d_L, RS, d_part);
      KCHECK();
    });
    CA_LAUNCH(k_corr_reduce, (G + 127) / 128, 128, 0, h->stream)(d_part, RS, G, d_sums);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(d_sums + (size_t)5 * G, tail.data(), sizeof(double) * tail.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // `tail` lives on this stack frame
    if (h->cfg.world > 1)   // collective: the sums over cells run over every shard
      NCCL_OK(nccl().AllReduce(d_sums, d_sums, (size_t)5 * G + 1 + C, kNcclFloat64, kNcclSum, h->comm, h->stream));
    CA_LAUNCH(k_corr_final, (G + 127) / 128, 128, 0, h->stream)(d_sums, d_L, G, C, d_out);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_pca_scores(ca_handle* h, int32_t max_iter, double tol, double* scores, int32_t* iters_out, char* err, size_t errlen) {
  try {
    if (!h || !scores || max_iter < 1) fail("bad argument");
    const bool sharded = h->cfg.world > 1;   // collective call: column statistics and X^T t are summed over the cell shards
    CUDA_OK(cudaSetDevice(h->dev));
    const int64_t N = h->N;
    const int G = h->G;
    if (h->Ntot < 2) fail("need at least two cells");
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    std::vector<void*> tmp;
    auto dalloc = [&](size_t n) {
      void* p = nullptr;
      CUDA_OK(cudaMalloc(&p, sizeof(double) * (n ? n : 1)));
      tmp.push_back(p);
      return (double*)p;
    };
    double *part = dalloc((size_t)RS * G * 2), *mean = dalloc(G), *inv_sd = dalloc(G), *v = dalloc(G), *w = dalloc(G), *a = dalloc(G),
           *b = dalloc(1), *t = dalloc(N), *tsum = dalloc(RS), *out2 = dalloc(2), *sums = dalloc((size_t)2 * G + 2);
    int* bad = (int*)dalloc(1);
    CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    int status = 0;
    std::string msg;
    try {
      dispatch_y(h, [&](auto* Yp) {
        using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
        dim3 gridc((G + 127) / 128, RS);
        CA_LAUNCH(k_pca_colstats<T>, gridc, 128, 0, h->stream)(Yp, h->ldY, N, G, RS, part);
        KCHECK();
        CA_LAUNCH(k_pca_colstats_reduce, (G + 127) / 128, 128, 0, h->stream)(part, RS, G, sums);
        KCHECK();
        if (sharded) NCCL_OK(nccl().AllReduce(sums, sums, (size_t)2 * G, kNcclFloat64, kNcclSum, h->comm, h->stream));
        CA_LAUNCH(k_pca_colstats_final, (G + 127) / 128, 128, 0, h->stream)(sums, G, (double)h->Ntot, mean, inv_sd, bad);
        KCHECK();
        int hbad = 0;
        CUDA_OK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        if (hbad) fail("cannot rescale a constant/zero column to unit variance");   // prcomp(..., scale = TRUE)
        // deterministic start: v_g proportional to 1 + (g mod 7) / 7 (not orthogonal to a dominant direction in practice)
        std::vector<double> v0(G);
        double nn = 0.0;
        for (int g = 0; g < G; ++g) { v0[g] = 1.0 + (double)(g % 7) / 7.0; nn += v0[g] * v0[g]; }
        for (int g = 0; g < G; ++g) v0[g] /= sqrt(nn);
        CUDA_OK(cudaMemcpyAsync(v, v0.data(), sizeof(double) * G, cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        int it = 0;
        for (; it < max_iter; ++it) {
          CA_LAUNCH(k_pca_prepare, 1, 1024, 0, h->stream)(v, mean, inv_sd, G, a, b);
          KCHECK();
          CA_LAUNCH(k_pca_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, h->stream)(Yp, h->ldY, N, G, a, b, t);
          KCHECK();
          CA_LAUNCH(k_pca_cols<T>, gridc, 128, 0, h->stream)(Yp, h->ldY, N, G, RS, t, part, tsum);
          KCHECK();
          CA_LAUNCH(k_pca_cols_reduce, (G + 1 + 127) / 128, 128, 0, h->stream)(part, tsum, RS, G, sums);
          KCHECK();
          if (sharded) NCCL_OK(nccl().AllReduce(sums, sums, (size_t)G + 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
          CA_LAUNCH(k_pca_update, 1, 1024, 0, h->stream)(sums, G, mean, inv_sd, v, w, out2);
          KCHECK();
          double o2[2];
          CUDA_OK(cudaMemcpyAsync(o2, out2, sizeof o2, cudaMemcpyDeviceToHost, h->stream));
          CUDA_OK(cudaStreamSynchronize(h->stream));
          if (o2[1] < tol) { ++it; break; }
        }
        if (iters_out) *iters_out = it;
        // scores of the converged direction: t = X v
        CA_LAUNCH(k_pca_prepare, 1, 1024, 0, h->stream)(v, mean, inv_sd, G, a, b);
        KCHECK();
        CA_LAUNCH(k_pca_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, h->stream)(Yp, h->ldY, N, G, a, b, t);
        KCHECK();
      });
      // sign convention (prcomp's is arbitrary): the loading of largest magnitude is positive
      std::vector<double> hv(G);
      CUDA_OK(cudaMemcpyAsync(hv.data(), v, sizeof(double) * G, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaMemcpyAsync(scores, t, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      int gmax = 0;
      for (int g = 1; g < G; ++g)
        if (fabs(hv[g]) > fabs(hv[gmax])) gmax = g;
      if (hv[gmax] < 0.0)
        for (int64_t n = 0; n < N; ++n) scores[n] = -scores[n];
    } catch (const std::exception& e) {
      status = 1;
      msg = e.what();
    }
    for (void* p : tmp) cudaFree(p);
    if (status) fail("%s", msg.c_str());
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_export(ca_handle* h, void* handle64, char* err, size_t errlen) {
  try {
    if (!h || !handle64) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_export: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // the zero-fill of the flags must have landed before a peer can signal
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t hd;
    CUDA_OK(cudaIpcGetMemHandle(&hd, h->p2p_buf));
    memcpy(handle64, &hd, 64);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_connect(ca_handle* h, const void* handles, char* err, size_t errlen) {
  try {
    if (!h || !handles) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_connect: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    const int world = h->cfg.world;
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)world * h->p2p_cnt_pad;
    for (int r = 0; r < world; ++r) {
      void* base = nullptr;
      if (r == h->cfg.rank) {
        base = h->p2p_buf;
      } else {
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles + 64 * (size_t)r, 64);
        CUDA_OK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_mapped[r] = base;
      }
      h->p2p_slots[r] = (float*)base;
      h->p2p_flags[r] = (unsigned*)((char*)base + slot_bytes + 256);
    }
    h->p2p_ready = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// variant p2p inside ONE process (ca_core_multi_*): the peers' exchange buffers are ordinary device pointers (CUDA IPC
// handles cannot be opened by the process that exported them); peer access is enabled on demand.
int ca_core_p2p_base(ca_handle* h, void** base, char* err, size_t errlen) {
  try {
    if (!h || !base) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_base: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // the zero-fill of the flags must have landed before a peer can signal
    *base = h->p2p_buf;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_connect_ptrs(ca_handle* h, void* const* bases, const int32_t* devices, char* err, size_t errlen) {
  try {
    if (!h || !bases || !devices) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_connect_ptrs: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    const int world = h->cfg.world;
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)world * h->p2p_cnt_pad;
    for (int r = 0; r < world; ++r) {
      if (r != h->cfg.rank && devices[r] != h->dev) {
        int can = 0;
        CUDA_OK(cudaDeviceCanAccessPeer(&can, h->dev, devices[r]));
        if (!can) fail("variant p2p: device %d cannot access device %d", h->dev, devices[r]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(e);
        cudaGetLastError();
      }
      h->p2p_slots[r] = (float*)bases[r];
      h->p2p_flags[r] = (unsigned*)((char*)bases[r] + slot_bytes + 256);
    }
    h->p2p_ready = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_ypass_many(ca_handle* const* hs, int32_t n, char* err, size_t errlen) {
  try {
    if (!hs || n < 1) fail("bad argument");
    ca_handle* h0 = hs[0];
    if (!h0) fail("null handle");
    for (int i = 0; i < n; ++i) {
      ca_handle* h = hs[i];
      if (!h) fail("null handle");
      if (h->KP != 1) fail("ca_core_ypass_many needs K + P == 1");
      if ((h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4)) && n > 1) fail("ca_core_ypass_many: the batched kernel uses the column tiling of ypass2 (sessions with variant ypass3 / ypass4 run their own pass)");
      if (h->Y != h0->Y || h->dev != h0->dev || h->N != h0->N || h->G != h0->G || h->ystore != h0->ystore)
        fail("ca_core_ypass_many: the sessions do not share one count matrix (create them with ca_core_create_shared)");
    }
    CUDA_OK(cudaSetDevice(h0->dev));
    for (int g0 = 0; g0 < n; g0 += kYMultiMax) {
      const int R = std::min(kYMultiMax, n - g0);
      ca_handle* lead = hs[g0];
      if (R == 1) {            // a lone fit: its own pass
        lead->ydirty = true;
        run_ypass(lead, lead->stream);
        continue;
      }
      // the pass runs on the first session's stream: it must see the parameter updates of the others, and their next
      // kernels must see its partial sums
      for (int r = 1; r < R; ++r) {
        CUDA_OK(cudaEventRecord(hs[g0 + r]->ev_fork, hs[g0 + r]->stream));
        CUDA_OK(cudaStreamWaitEvent(lead->stream, hs[g0 + r]->ev_fork, 0));
      }
      YMultiArgs a;
      for (int r = 0; r < kYMultiMax; ++r) {
        ca_handle* h = hs[g0 + (r < R ? r : 0)];
        a.U[r] = h->U; a.Vm[r] = h->Vm; a.rowpart[r] = h->rowpart; a.colpart[r] = h->colpart;
      }
      dispatch_y(lead, [&](auto* Yp) {
        using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
        dim3 grid(lead->nCB, lead->nRB);
        if (R == 2) { auto k = k_ypass_k1_multi<T, 2>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        else if (R == 3) { auto k = k_ypass_k1_multi<T, 3>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        else { auto k = k_ypass_k1_multi<T, 4>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        KCHECK();
      });
      CUDA_OK(cudaEventRecord(lead->ev_join, lead->stream));
      for (int r = 0; r < R; ++r) {
        if (r > 0) CUDA_OK(cudaStreamWaitEvent(hs[g0 + r]->stream, lead->ev_join, 0));
        hs[g0 + r]->ydirty = false;
      }
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_describe(ca_handle* h, char* json, size_t json_len) {
  if (!h || !json || !json_len) return 1;
  const char* st = h->ystore == CA_STORE_F32 ? "f32" : (h->ystore == CA_STORE_U16 ? "u16" : "u8");
  int bpe = h->ystore == CA_STORE_F32 ? 4 : (h->ystore == CA_STORE_U16 ? 2 : 1);
  // interp path: the panel structure of the last step (device-side data: the node work is proportional to it)
  InterpPlan pl;
  memset(&pl, 0, sizeof pl);
  if (h->iplan) {
    cudaSetDevice(h->dev);
    cudaMemcpyAsync(&pl, h->iplan, sizeof pl, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  snprintf(json, json_len,
           "{\"N\": %lld, \"G\": %d, \"C\": %d, \"S\": %d, \"K\": %d, \"P\": %d, \"path\": \"%s\", \"y_store\": \"%s\", "
           "\"y_bytes_per_entry\": %d, \"ldY\": %lld, \"launches_last_step\": %d, \"nsplit\": %d, \"fsplit\": %d, "
           "\"SCp\": %d, \"J\": %d, \"world\": %d, \"rank\": %d, \"variants\": %u, \"ypass_grid\": [%d, %d], \"ypass_rows_per_block\": %d, "
           "\"num_sms\": %d, \"panels\": {\"nf_neg\": %d, \"nf_pos\": %d, \"nb\": %d, \"w_range\": [%.6g, %.6g], \"psi_range\": [%.6g, %.6g]}}",
           (long long)h->N, h->G, h->C, h->S, h->K, h->P, h->interp ? "interp" : (h->tc ? "tcgen05" : "cudacore"), st, bpe, (long long)h->ldY,
           h->launches_last_step, h->nsplit, h->tc ? h->tcplan.fsplit : 1, h->SCp, h->J, h->cfg.world, h->cfg.rank, h->variants, h->nCB, h->nRB, h->RB,
           h->num_sms, pl.nf_neg, pl.nf_pos, pl.nb, pl.wmin, pl.wmax, pl.pmin, pl.pmax);
  return 0;
}

}  // extern "C"
