/* .Call shim between R and libclonealign_b200.so (include/clonealign_b200.h).
 *
 * NOT LINKED OR RUN IN THIS REPOSITORY'S IMAGE: there is no R toolchain here (no R, no Rinternals.h); it is only
 * syntax- and type-checked against stub declarations of the R API (tests/r_stub/, tests/test_host.py).
 * It is the binding a clonealign maintainer adds under src/ (plus `useDynLib(clonealign, .registration = TRUE)` in
 * NAMESPACE and `LinkingTo`-free `PKG_LIBS = -lclonealign_b200` in src/Makevars).  Every numeric operation lives in
 * the extern "C" core, which is what the Python/ctypes tests exercise with R-layout (column-major double) inputs.
 *
 * Session lifecycle replaced (R/inference-tflow.R): sess$run(init) :353 -> ca_create; gamma_init/init_gamma :368-369 ->
 * ca_init_gamma; sess$run(train) :401 -> ca_step; sess$run(elbo) :372,403,448 -> ca_elbo; fetch :424-440 -> ca_params;
 * sess$close() :457 -> ca_destroy (also the external pointer's finalizer).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "clonealign_b200.h"

#define ERRLEN 1024

static void ca_finalizer(SEXP ptr) {
  ca_handle* h = (ca_handle*)R_ExternalPtrAddr(ptr);
  if (h) {
    ca_core_destroy(h);
    R_ClearExternalPtr(ptr);
  }
}

static ca_handle* get_handle(SEXP ptr) {
  ca_handle* h = (ca_handle*)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("clonealign CUDA session is closed");
  return h;
}

static const double* real_or_null(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

/* ca_create(Y, L, psi_init, loc_init, X, clone_allele, alt, cov, S, K, lr, seed, device)
 * Y: numeric or integer N x G matrix (column-major, as R stores it). */
SEXP ca_create(SEXP Y, SEXP L, SEXP psi_init, SEXP loc_init, SEXP X, SEXP clone_allele, SEXP alt, SEXP cov,
               SEXP S, SEXP K, SEXP lr, SEXP seed, SEXP device) {
  char err[ERRLEN] = {0};
  SEXP dim = Rf_getAttrib(Y, R_DimSymbol);
  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = INTEGER(dim)[0];
  cfg.N_total = cfg.N;
  cfg.G = INTEGER(dim)[1];
  cfg.C = Rf_ncols(L);
  cfg.S = Rf_asInteger(S);
  cfg.K = Rf_asInteger(K);
  cfg.P = Rf_isNull(X) ? 0 : Rf_ncols(X);
  cfg.V = Rf_isNull(clone_allele) ? 0 : Rf_nrows(clone_allele);
  cfg.learning_rate = Rf_asReal(lr);
  cfg.seed = (uint64_t)Rf_asInteger(seed);
  cfg.device = Rf_asInteger(device);
  cfg.rank = 0;
  cfg.world = 1;
  cfg.y_dtype = (TYPEOF(Y) == INTSXP) ? CA_Y_I32 : CA_Y_F64;
  cfg.y_layout = CA_Y_COLMAJOR;
  cfg.y_mem = CA_Y_HOST;
  cfg.y_store = CA_STORE_AUTO;
  cfg.path = CA_PATH_AUTO;
  const void* yptr = (TYPEOF(Y) == INTSXP) ? (const void*)INTEGER(Y) : (const void*)REAL(Y);
  ca_handle* h = NULL;
  int st = ca_core_create(&h, &cfg, yptr, REAL(L), real_or_null(psi_init), REAL(loc_init), real_or_null(X), NULL,
                          real_or_null(clone_allele), real_or_null(alt), real_or_null(cov), err, ERRLEN);
  if (st != 0) Rf_error("%s", err);   /* nothing allocated on the R side yet */
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ca_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* ca_create_sparse(dim, p, i, x, L, ...): the counts assay of a SingleCellExperiment as it is -- a genes x cells
 * dgCMatrix (slots @Dim, @p, @i, @x) is the compressed-row form of the cells x genes matrix, so nothing is transposed
 * or densified (replaces t(as.matrix(assay(...))), R/clonealign.R:217).  The gene filter must already be applied. */
SEXP ca_create_sparse(SEXP dim, SEXP p, SEXP i, SEXP x, SEXP L, SEXP psi_init, SEXP loc_init, SEXP X, SEXP clone_allele,
                      SEXP alt, SEXP cov, SEXP S, SEXP K, SEXP lr, SEXP seed, SEXP device) {
  char err[ERRLEN] = {0};
  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.G = INTEGER(dim)[0];        /* dgCMatrix rows = genes */
  cfg.N = INTEGER(dim)[1];        /* dgCMatrix columns = cells */
  cfg.N_total = cfg.N;
  cfg.C = Rf_ncols(L);
  cfg.S = Rf_asInteger(S);
  cfg.K = Rf_asInteger(K);
  cfg.P = Rf_isNull(X) ? 0 : Rf_ncols(X);
  cfg.V = Rf_isNull(clone_allele) ? 0 : Rf_nrows(clone_allele);
  cfg.learning_rate = Rf_asReal(lr);
  cfg.seed = (uint64_t)Rf_asInteger(seed);
  cfg.device = Rf_asInteger(device);
  cfg.world = 1;
  cfg.y_dtype = CA_Y_F64;         /* @x is numeric */
  cfg.y_layout = CA_Y_CSR;
  cfg.y_mem = CA_Y_HOST;
  cfg.y_store = CA_STORE_AUTO;
  cfg.path = CA_PATH_AUTO;
  cfg.y_indptr = INTEGER(p);
  cfg.y_indices = INTEGER(i);
  ca_handle* h = NULL;
  int st = ca_core_create(&h, &cfg, REAL(x), REAL(L), real_or_null(psi_init), REAL(loc_init), real_or_null(X), NULL,
                          real_or_null(clone_allele), real_or_null(alt), real_or_null(cov), err, ERRLEN);
  if (st != 0) Rf_error("%s", err);
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ca_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* prcomp(log2(Y_dat + 1), center = TRUE, scale = TRUE)$x[, 1] (R/inference-tflow.R:203-204) on the resident Y */
SEXP ca_pca_scores(SEXP ptr, SEXP n_cells, SEXP max_iter, SEXP tol) {
  char err[ERRLEN] = {0};
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_cells)));
  int iters = 0;
  int st = ca_core_pca_scores(get_handle(ptr), Rf_asInteger(max_iter), Rf_asReal(tol), REAL(out), &iters, err, ERRLEN);
  UNPROTECT(1);
  if (st != 0) Rf_error("%s", err);
  return out;
}

/* psi <- scale(pcs) + rnorm noise (:205-207), written after ca_pca_scores */
SEXP ca_set_psi(SEXP ptr, SEXP psi) {
  char err[ERRLEN] = {0};
  if (ca_core_set_array(get_handle(ptr), "psi", REAL(psi), (int64_t)XLENGTH(psi), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}

/* compute_correlations(Y, L, clones) (R/clonealign.R:318-334) on the resident Y; clone_idx: 0-based, -1 = unassigned */
SEXP ca_correlations(SEXP ptr, SEXP clone_idx, SEXP L, SEXP n_genes) {
  char err[ERRLEN] = {0};
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
  int st = ca_core_correlations(get_handle(ptr), INTEGER(clone_idx), real_or_null(L), REAL(out), err, ERRLEN);
  UNPROTECT(1);
  if (st != 0) Rf_error("%s", err);
  return out;   /* NaN where cor() gives NA */
}

/* ---- restarts of run_clonealign (R/clonealign.R:50-56) on shared device inputs ------------------------------------
 * ca_data_create(Y, L, clone_allele, alt, cov, device): upload + preprocess once; ca_create_shared(data, psi_init,
 * loc_init, X, S, K, lr, seed, dims = c(N, G, C, V)): one session per restart; ca_ypass_many(list of sessions): one pass
 * over the shared count matrix for all of them (call between their ca_step and ca_elbo calls). */
static void ca_data_finalizer(SEXP ptr) {
  ca_data* d = (ca_data*)R_ExternalPtrAddr(ptr);
  if (d) {
    char err[ERRLEN] = {0};
    if (ca_core_data_destroy(d, err, ERRLEN) == 0) R_ClearExternalPtr(ptr);   /* else: sessions still alive, retried later */
  }
}

SEXP ca_data_create(SEXP Y, SEXP L, SEXP clone_allele, SEXP alt, SEXP cov, SEXP device) {
  char err[ERRLEN] = {0};
  SEXP dim = Rf_getAttrib(Y, R_DimSymbol);
  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = INTEGER(dim)[0];
  cfg.N_total = cfg.N;
  cfg.G = INTEGER(dim)[1];
  cfg.C = Rf_ncols(L);
  cfg.S = 1;
  cfg.V = Rf_isNull(clone_allele) ? 0 : Rf_nrows(clone_allele);
  cfg.device = Rf_asInteger(device);
  cfg.world = 1;
  cfg.y_dtype = (TYPEOF(Y) == INTSXP) ? CA_Y_I32 : CA_Y_F64;
  cfg.y_layout = CA_Y_COLMAJOR;
  cfg.y_mem = CA_Y_HOST;
  cfg.y_store = CA_STORE_AUTO;
  const void* yptr = (TYPEOF(Y) == INTSXP) ? (const void*)INTEGER(Y) : (const void*)REAL(Y);
  ca_data* d = NULL;
  if (ca_core_data_create(&d, &cfg, yptr, REAL(L), NULL, real_or_null(clone_allele), real_or_null(alt), real_or_null(cov), err,
                          ERRLEN))
    Rf_error("%s", err);
  SEXP ptr = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ca_data_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

SEXP ca_create_shared(SEXP data, SEXP psi_init, SEXP loc_init, SEXP X, SEXP S, SEXP K, SEXP lr, SEXP seed, SEXP dims) {
  char err[ERRLEN] = {0};
  ca_data* d = (ca_data*)R_ExternalPtrAddr(data);
  if (!d) Rf_error("clonealign CUDA inputs are closed");
  int* dm = INTEGER(dims);   /* c(N, G, C, V, device) */
  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = dm[0]; cfg.N_total = dm[0]; cfg.G = dm[1]; cfg.C = dm[2]; cfg.V = dm[3]; cfg.device = dm[4];
  cfg.S = Rf_asInteger(S);
  cfg.K = Rf_asInteger(K);
  cfg.P = Rf_isNull(X) ? 0 : Rf_ncols(X);
  cfg.learning_rate = Rf_asReal(lr);
  cfg.seed = (uint64_t)Rf_asInteger(seed);
  cfg.world = 1;
  cfg.path = CA_PATH_AUTO;
  ca_handle* h = NULL;
  if (ca_core_create_shared(&h, &cfg, d, real_or_null(psi_init), REAL(loc_init), real_or_null(X), err, ERRLEN)) Rf_error("%s", err);
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, data, R_NilValue));   /* tag = the inputs: kept alive as long as the session */
  R_RegisterCFinalizerEx(ptr, ca_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

SEXP ca_ypass_many(SEXP sessions) {   /* list of session external pointers */
  char err[ERRLEN] = {0};
  R_xlen_t n = XLENGTH(sessions);
  ca_handle** hs = (ca_handle**)R_alloc((size_t)n, sizeof(ca_handle*));
  for (R_xlen_t i = 0; i < n; ++i) hs[i] = get_handle(VECTOR_ELT(sessions, i));
  if (ca_core_ypass_many(hs, (int32_t)n, err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}

SEXP ca_init_gamma(SEXP ptr) {
  char err[ERRLEN] = {0};
  if (ca_core_init_gamma(get_handle(ptr), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}

SEXP ca_step(SEXP ptr) {
  char err[ERRLEN] = {0};
  if (ca_core_step(get_handle(ptr), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}

SEXP ca_elbo(SEXP ptr) {
  char err[ERRLEN] = {0};
  double e = NA_REAL;
  if (ca_core_elbo(get_handle(ptr), &e, err, ERRLEN)) Rf_error("%s", err);
  return Rf_ScalarReal(e);   /* NaN propagates; R keeps stop("Initial elbo is NA") */
}

/* replicate(n, sess$run(elbo)) (R/inference-tflow.R:447-449) with one device-to-host copy */
SEXP ca_elbo_many(SEXP ptr, SEXP n) {
  char err[ERRLEN] = {0};
  int k = Rf_asInteger(n);
  if (k < 0) Rf_error("ca_elbo_many: n must be >= 0");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, k));
  if (ca_core_elbo_many(get_handle(ptr), k, REAL(out), err, ERRLEN)) {
    UNPROTECT(1);
    Rf_error("%s", err);
  }
  UNPROTECT(1);
  return out;
}

/* named list: mu, clone_probs, s, alpha [, beta] [, psi, W, chi] [, clone_probs_from_snv] */
SEXP ca_params(SEXP ptr, SEXP dims) {   /* dims = c(N, G, C, K, P, V) */
  char err[ERRLEN] = {0};
  int* d = INTEGER(dims);
  int N = d[0], G = d[1], C = d[2], K = d[3], P = d[4], V = d[5];
  int np = 0;
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, G)); np++;
  SEXP cp = PROTECT(Rf_allocMatrix(REALSXP, N, C)); np++;
  SEXP s = PROTECT(Rf_allocVector(REALSXP, N)); np++;
  SEXP alpha = PROTECT(Rf_allocVector(REALSXP, C)); np++;
  SEXP psi = PROTECT(K > 0 ? Rf_allocMatrix(REALSXP, N, K) : R_NilValue); np++;
  SEXP W = PROTECT(K > 0 ? Rf_allocMatrix(REALSXP, G, K) : R_NilValue); np++;
  SEXP chi = PROTECT(K > 0 ? Rf_allocVector(REALSXP, K) : R_NilValue); np++;
  SEXP beta = PROTECT(P > 0 ? Rf_allocMatrix(REALSXP, G, P) : R_NilValue); np++;
  SEXP snv = PROTECT(V > 0 ? Rf_allocMatrix(REALSXP, N, C) : R_NilValue); np++;
  int st = ca_core_params(get_handle(ptr), REAL(mu), REAL(cp), REAL(s), REAL(alpha), K > 0 ? REAL(psi) : NULL,
                          K > 0 ? REAL(W) : NULL, K > 0 ? REAL(chi) : NULL, P > 0 ? REAL(beta) : NULL,
                          V > 0 ? REAL(snv) : NULL, err, ERRLEN);
  if (st != 0) { UNPROTECT(np); Rf_error("%s", err); }
  const char* names[] = {"mu", "clone_probs", "s", "alpha", "psi", "W", "chi", "beta", "clone_probs_from_snv", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names)); np++;
  SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, cp); SET_VECTOR_ELT(out, 2, s); SET_VECTOR_ELT(out, 3, alpha);
  SET_VECTOR_ELT(out, 4, psi); SET_VECTOR_ELT(out, 5, W); SET_VECTOR_ELT(out, 6, chi); SET_VECTOR_ELT(out, 7, beta);
  SET_VECTOR_ELT(out, 8, snv);
  UNPROTECT(np);
  return out;
}

SEXP ca_set_eps(SEXP ptr, SEXP eps, SEXP n_draws) {   /* eps: numeric, sample-major S*G per draw */
  char err[ERRLEN] = {0};
  R_xlen_t n = XLENGTH(eps);
  float* f = (float*)R_alloc(n, sizeof(float));
  for (R_xlen_t i = 0; i < n; ++i) f[i] = (float)REAL(eps)[i];
  if (ca_core_set_eps(get_handle(ptr), f, Rf_asInteger(n_draws), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}

SEXP ca_destroy(SEXP ptr) {
  ca_finalizer(ptr);
  return R_NilValue;
}

/* ---- one fit over several GPUs from the (single-threaded) R interpreter: options(clonealign.gpus = c(0, 1, ...)) ----
 * ca_core_multi_* shards the cells inside the library (one worker thread per device, none of them touches R) and mirrors
 * the same session lifecycle; ca_multi_create takes the arguments of ca_create with `devices` (integer vector) last. */
static void ca_multi_finalizer(SEXP ptr) {
  ca_multi* m = (ca_multi*)R_ExternalPtrAddr(ptr);
  if (m) {
    ca_core_multi_destroy(m);
    R_ClearExternalPtr(ptr);
  }
}
static ca_multi* get_multi(SEXP ptr) {
  ca_multi* m = (ca_multi*)R_ExternalPtrAddr(ptr);
  if (!m) Rf_error("clonealign CUDA session is closed");
  return m;
}
SEXP ca_multi_create(SEXP Y, SEXP L, SEXP psi_init, SEXP loc_init, SEXP X, SEXP clone_allele, SEXP alt, SEXP cov,
                     SEXP S, SEXP K, SEXP lr, SEXP seed, SEXP devices) {
  char err[ERRLEN] = {0};
  SEXP dim = Rf_getAttrib(Y, R_DimSymbol);
  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = INTEGER(dim)[0];
  cfg.N_total = cfg.N;
  cfg.G = INTEGER(dim)[1];
  cfg.C = Rf_ncols(L);
  cfg.S = Rf_asInteger(S);
  cfg.K = Rf_asInteger(K);
  cfg.P = Rf_isNull(X) ? 0 : Rf_ncols(X);
  cfg.V = Rf_isNull(clone_allele) ? 0 : Rf_nrows(clone_allele);
  cfg.learning_rate = Rf_asReal(lr);
  cfg.seed = (uint64_t)Rf_asInteger(seed);
  cfg.world = 1;
  cfg.y_dtype = (TYPEOF(Y) == INTSXP) ? CA_Y_I32 : CA_Y_F64;
  cfg.y_layout = CA_Y_COLMAJOR;
  cfg.y_mem = CA_Y_HOST;
  cfg.y_store = CA_STORE_AUTO;
  cfg.path = CA_PATH_AUTO;
  const void* yptr = (TYPEOF(Y) == INTSXP) ? (const void*)INTEGER(Y) : (const void*)REAL(Y);
  ca_multi* m = NULL;
  int st = ca_core_multi_create(&m, &cfg, INTEGER(devices), (int32_t)XLENGTH(devices), yptr, REAL(L), real_or_null(psi_init),
                                REAL(loc_init), real_or_null(X), real_or_null(clone_allele), real_or_null(alt), real_or_null(cov),
                                err, ERRLEN);
  if (st != 0) Rf_error("%s", err);
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ca_multi_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
SEXP ca_multi_init_gamma(SEXP ptr) {
  char err[ERRLEN] = {0};
  if (ca_core_multi_init_gamma(get_multi(ptr), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}
SEXP ca_multi_step(SEXP ptr) {
  char err[ERRLEN] = {0};
  if (ca_core_multi_step(get_multi(ptr), err, ERRLEN)) Rf_error("%s", err);
  return R_NilValue;
}
SEXP ca_multi_elbo(SEXP ptr) {
  char err[ERRLEN] = {0};
  double e = NA_REAL;
  if (ca_core_multi_elbo(get_multi(ptr), &e, err, ERRLEN)) Rf_error("%s", err);
  return Rf_ScalarReal(e);
}
SEXP ca_multi_elbo_many(SEXP ptr, SEXP n) {
  char err[ERRLEN] = {0};
  int k = Rf_asInteger(n);
  if (k < 0) Rf_error("ca_multi_elbo_many: n must be >= 0");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, k));
  if (ca_core_multi_elbo_many(get_multi(ptr), k, REAL(out), err, ERRLEN)) {
    UNPROTECT(1);
    Rf_error("%s", err);
  }
  UNPROTECT(1);
  return out;
}
SEXP ca_multi_params(SEXP ptr, SEXP dims) {   /* dims = c(N, G, C, K, P, V); N = all cells */
  char err[ERRLEN] = {0};
  int* d = INTEGER(dims);
  int N = d[0], G = d[1], C = d[2], K = d[3], P = d[4], V = d[5];
  int np = 0;
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, G)); np++;
  SEXP cp = PROTECT(Rf_allocMatrix(REALSXP, N, C)); np++;
  SEXP s = PROTECT(Rf_allocVector(REALSXP, N)); np++;
  SEXP alpha = PROTECT(Rf_allocVector(REALSXP, C)); np++;
  SEXP psi = PROTECT(K > 0 ? Rf_allocMatrix(REALSXP, N, K) : R_NilValue); np++;
  SEXP W = PROTECT(K > 0 ? Rf_allocMatrix(REALSXP, G, K) : R_NilValue); np++;
  SEXP chi = PROTECT(K > 0 ? Rf_allocVector(REALSXP, K) : R_NilValue); np++;
  SEXP beta = PROTECT(P > 0 ? Rf_allocMatrix(REALSXP, G, P) : R_NilValue); np++;
  SEXP snv = PROTECT(V > 0 ? Rf_allocMatrix(REALSXP, N, C) : R_NilValue); np++;
  int st = ca_core_multi_params(get_multi(ptr), REAL(mu), REAL(cp), REAL(s), REAL(alpha), K > 0 ? REAL(psi) : NULL,
                                K > 0 ? REAL(W) : NULL, K > 0 ? REAL(chi) : NULL, P > 0 ? REAL(beta) : NULL,
                                V > 0 ? REAL(snv) : NULL, err, ERRLEN);
  if (st != 0) { UNPROTECT(np); Rf_error("%s", err); }
  const char* names[] = {"mu", "clone_probs", "s", "alpha", "psi", "W", "chi", "beta", "clone_probs_from_snv", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names)); np++;
  SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, cp); SET_VECTOR_ELT(out, 2, s); SET_VECTOR_ELT(out, 3, alpha);
  SET_VECTOR_ELT(out, 4, psi); SET_VECTOR_ELT(out, 5, W); SET_VECTOR_ELT(out, 6, chi); SET_VECTOR_ELT(out, 7, beta);
  SET_VECTOR_ELT(out, 8, snv);
  UNPROTECT(np);
  return out;
}
SEXP ca_multi_destroy(SEXP ptr) {
  ca_multi_finalizer(ptr);
  return R_NilValue;
}

static const R_CallMethodDef call_methods[] = {
    {"ca_multi_create", (DL_FUNC)&ca_multi_create, 13}, {"ca_multi_init_gamma", (DL_FUNC)&ca_multi_init_gamma, 1},
    {"ca_multi_step", (DL_FUNC)&ca_multi_step, 1}, {"ca_multi_elbo", (DL_FUNC)&ca_multi_elbo, 1},
    {"ca_multi_elbo_many", (DL_FUNC)&ca_multi_elbo_many, 2}, {"ca_multi_params", (DL_FUNC)&ca_multi_params, 2},
    {"ca_multi_destroy", (DL_FUNC)&ca_multi_destroy, 1},
    {"ca_create", (DL_FUNC)&ca_create, 13}, {"ca_init_gamma", (DL_FUNC)&ca_init_gamma, 1},
    {"ca_step", (DL_FUNC)&ca_step, 1},      {"ca_elbo", (DL_FUNC)&ca_elbo, 1},
    {"ca_params", (DL_FUNC)&ca_params, 2},  {"ca_set_eps", (DL_FUNC)&ca_set_eps, 3},
    {"ca_destroy", (DL_FUNC)&ca_destroy, 1}, {"ca_create_sparse", (DL_FUNC)&ca_create_sparse, 16},
    {"ca_pca_scores", (DL_FUNC)&ca_pca_scores, 4}, {"ca_set_psi", (DL_FUNC)&ca_set_psi, 2},
    {"ca_correlations", (DL_FUNC)&ca_correlations, 4}, {"ca_data_create", (DL_FUNC)&ca_data_create, 6},
    {"ca_create_shared", (DL_FUNC)&ca_create_shared, 9}, {"ca_ypass_many", (DL_FUNC)&ca_ypass_many, 1},
    {"ca_elbo_many", (DL_FUNC)&ca_elbo_many, 2}, {NULL, NULL, 0}};

void R_init_clonealign(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
