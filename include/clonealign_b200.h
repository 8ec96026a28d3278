/* clonealign_b200 — C-ABI of the B200 backend for clonealign's variational hot path.
 *
 * This is the drop-in boundary.  The reference (kieranrcampbell/clonealign, R) runs the path as a
 * TensorFlow-1 session inside `inference_tflow()` (R/inference-tflow.R:71-481); each entry point
 * below replaces one use of that session and is what an R `.Call` shim (r/src/ca_shim.c,
 * INTEGRATION.md) or the Python host mirror (clonealign_b200/session.py, ctypes) binds:
 *
 *   ca_core_create      graph build + sess$run(init)            R/inference-tflow.R:240-353
 *   ca_core_init_gamma  sess$run(gamma_init) + assign           R/inference-tflow.R:338-342,368-369
 *   ca_core_step        sess$run(train)                         R/inference-tflow.R:345-346,401
 *   ca_core_elbo        sess$run(elbo)                          R/inference-tflow.R:336,372,403,448
 *   ca_core_elbo_many   replicate(20, sess$run(elbo))           R/inference-tflow.R:447-449 (one host round trip)
 *   ca_core_params      sess$run(list(softplus(loc),gamma,...)) R/inference-tflow.R:424-440
 *   ca_core_destroy     sess$close()                            R/inference-tflow.R:457
 *
 * and, for the host work on either side of the session:
 *
 *   ca_core_create, y_layout = CA_Y_CSR   t(as.matrix(assay(sce, "counts")))           R/clonealign.R:217
 *   ca_core_pca_scores                    prcomp(log2(Y_dat + 1), center, scale)$x     R/inference-tflow.R:203-205
 *   ca_core_correlations                  compute_correlations(Y, L, clones)           R/clonealign.R:292-294,318-334
 *   ca_core_data_stats                    rowSums / colSums / colMeans(Y / rowMeans(Y)) R/inference-tflow.R:117,210,222
 *   ca_core_data_create / _create_shared  the per-restart repetition of the set-up in  R/clonealign.R:50-56
 *   ca_core_ypass_many                    run_clonealign()'s loop (one Y pass for all restarts of a device)
 *   ca_core_p2p_export / _connect         (no counterpart: the reference is single-process; SURVEY.md 8e exchange)
 *
 * Conventions
 *   - Plain pointers and sizes only; no C++/torch types.  Every call returns 0 on success and a
 *     non-zero status with a message in `err` (NUL-terminated, truncated to errlen) otherwise.
 *     Nothing throws or longjmps across this boundary.
 *   - Small dense matrices cross the ABI as COLUMN-MAJOR doubles, exactly as R stores them
 *     (L: G x C, psi: N x K, X: N x P, clone_allele: V x C, alt/cov: N x V, outputs likewise).
 *   - The count matrix Y (cells x genes) is the one large operand; its element type, layout and
 *     memory space are described by ca_config so that R (column-major double / integer on the
 *     host) and a CUDA-aware caller (row-major float already in HBM) both avoid extra copies.
 *   - The caller owns every input buffer; the library copies what it needs during ca_core_create
 *     and never retains caller pointers.
 *   - There is NO CPU fallback: without a usable CUDA device every call fails with a message.
 */
#ifndef CLONEALIGN_B200_H
#define CLONEALIGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_ABI_VERSION 6
#if defined(__GNUC__)
#define CA_API __attribute__((visibility("default")))
#else
#define CA_API
#endif

typedef struct ca_handle ca_handle;
typedef struct ca_data ca_data;

enum ca_y_dtype  { CA_Y_F64 = 0, CA_Y_F32 = 1, CA_Y_I32 = 2,
                   CA_Y_U8 = 3, CA_Y_U16 = 4 /* compact host counts: 8x / 4x less to move over PCIe than R doubles */ };
enum ca_y_layout { CA_Y_COLMAJOR = 0 /* R matrix: cell index fastest */, CA_Y_ROWMAJOR = 1 /* gene index fastest */,
                   /* compressed sparse rows of the cells x genes matrix == the genes x cells dgCMatrix of a
                    * SingleCellExperiment as it is (slots @p, @i, @x): Y points at the nnz values (y_dtype),
                    * y_indptr at the N + 1 row offsets, y_indices at the nnz gene indices (0-based, unique per cell).
                    * Host memory only.  Avoids t(as.matrix(assay(...))) (R/clonealign.R:217), SURVEY.md 8f-2. */
                   CA_Y_CSR = 2 };
enum ca_y_mem    { CA_Y_HOST = 0, CA_Y_DEVICE = 1 };
/* how Y is kept in HBM: fp32 (the reference's tensor dtype) or, when every count is an integer that
 * fits, a narrower unsigned type.  AUTO picks the narrowest exact representation. */
enum ca_y_store  { CA_STORE_AUTO = 0, CA_STORE_F32 = 1, CA_STORE_U16 = 2, CA_STORE_U8 = 3 };
/* which contraction kernels run: AUTO = tcgen05 path when K == 1 and P == 0 (the reference's default
 * model), CUDA-core fp32 path otherwise.  INTERP (K == 1, P == 0 only) replaces both contractions by piecewise
 * Chebyshev interpolation of the univariate functions they reduce to (kernels_interp.cuh); opt-in until it has
 * been validated on hardware. */
enum ca_path     { CA_PATH_AUTO = 0, CA_PATH_CUDACORE = 1, CA_PATH_TENSOR = 2, CA_PATH_INTERP = 3 };
/* kernel variants (bit mask in ca_config.variants; 0 = the kernels measured in round 1).  Each bit swaps ONE kernel for
 * a re-engineered version with the same inputs/outputs; they are opt-in until measured on hardware (bench.py validates
 * them on the device against the default kernels before using them).
 *   YPASS2: Y pass on packed fp32 pairs (add/fma.rn.f32x2), see kernels_ypass.cuh
 *   YPASS3: Y pass without a conversion of the integer counts (the stored byte / half-word is used as a denormal fp32
 *           operand of the packed FMA, the other operand carries the scale) and 16 columns per thread for u8 storage;
 *           alternative to YPASS2, see kernels_ypass.cuh
 *   DEFER : (with EPI2 + LEAN) the per-cell kernel does not read the Y-pass partials: psi_n (YW)_n joins the ELBO and (YW)_n
 *           joins d psi_n afterwards (k_yv_dot, k_adam_all), so the Y pass is joined only before the gene-gradient kernel and,
 *           with OVERLAP, runs next to the per-cell kernel and the backward node sums instead of before them
 *   EPI2  : (interp path) evaluation of the interpolants fused into a leaner per-cell epilogue, see kernels_fused.cuh
 *   LEAN  : (with EPI2) gene-level / scalar / optimiser work in 3 launches instead of 12, see kernels_fused.cuh
 *   P2P   : (world > 1) the per-step all-reduce as one kernel over NVLink peer memory instead of ncclAllReduce; needs
 *           ca_core_p2p_export / ca_core_p2p_connect after ca_core_create, see kernels_p2p.cuh
 *   OVERLAP: the Y pass (HBM-bound; needs only Y, psi, W) is forked onto a second stream at the start of the step and
 *           joined before the per-cell kernel, so it runs next to the small gene-level launches instead of after them
 *   YPASS4: the arithmetic and tiling of YPASS3 with the rows staged through a per-thread shared-memory ring by cp.async
 *           (bytes in flight while the thread computes, half the registers), launched as a persistent grid of 2 CTAs per SM
 *   COSCHED: (with DEFER + YPASS4) the Y pass is started FIRST in the step on the second stream and joined before the
 *           gene-gradient kernel; the per-cell kernel runs with 16 warps so that both fit on every SM: the HBM-bound stream
 *           runs next to the issue-bound kernels of the whole step
 *   CELL2 : (with EPI2 + LEAN + DEFER; takes effect for S <= 8) second-generation per-cell / per-gene kernels, see kernels_cell.cuh:
 *           lane = (cell, clone) resp. (gene, clone), Horner evaluation of the interpolants in the monomial basis, the w-weighted
 *           columns taken from the derivative of the interpolants (node sums over the S*C normaliser columns only), the gene kernel
 *           in front of the Y-pass join, the gamma-logit Adam update inside the per-cell kernel; part of the default set
 *   YPASS5: (takes effect with YPASS4 and counts stored as u8) the two products of the Y pass as exact integer contractions on
 *           the tensor pipe (mma.sync u8 x s8 on base-128 digits of W and psi), one persistent CTA per SM, see kernels_ypass.cuh.
 *           path = auto adds it, on its tensor-copy kernel (k_ypass_k1_v7, kernels_ypass_tma.cuh), for u8 matrices of >= 2^25 counts
 *           per rank; passed explicitly it selects the row-copy kernels k_ypass_k1_v5 / v6 (A/B partners) */
enum ca_variant  { CA_VAR_YPASS2 = 1, CA_VAR_EPI2 = 2, CA_VAR_LEAN = 4, CA_VAR_P2P = 8, CA_VAR_OVERLAP = 16, CA_VAR_YPASS3 = 32, CA_VAR_DEFER = 64,
                   CA_VAR_YPASS4 = 128, CA_VAR_COSCHED = 256, CA_VAR_CELL2 = 512, CA_VAR_YPASS5 = 1024 };

typedef struct ca_config {
  int64_t N;            /* cells held by this handle (this rank's shard)                       */
  int64_t N_total;      /* cells over all ranks (== N when world == 1)                         */
  int32_t G;            /* genes (after the host-side gene filter)                             */
  int32_t C;            /* clones                                                              */
  int32_t S;            /* Monte-Carlo samples (mc_samples)                                    */
  int32_t K;            /* latent dimensions of psi / W                                        */
  int32_t P;            /* covariates (columns of x)                                           */
  int32_t V;            /* variants for the allele-specific likelihood; 0 = not used           */
  double  learning_rate;
  uint64_t seed;        /* seeds the counter-based N(0,1) generator (get_next_seed() in R)     */
  int32_t device;       /* CUDA device ordinal                                                 */
  int32_t rank;         /* 0..world-1                                                          */
  int32_t world;        /* number of cell shards / GPUs                                        */
  int32_t y_dtype;      /* enum ca_y_dtype                                                     */
  int32_t y_layout;     /* enum ca_y_layout                                                    */
  int32_t y_mem;        /* enum ca_y_mem                                                       */
  int32_t y_store;      /* enum ca_y_store                                                     */
  int32_t path;         /* enum ca_path                                                        */
  int64_t y_ld;         /* leading dimension of Y in elements (0 = tight)                      */
  const void* nccl_id;  /* 128-byte ncclUniqueId shared by all ranks when world > 1, else NULL */
  uint32_t variants;    /* bit mask of enum ca_variant; 0 = default kernels                    */
  const int32_t* y_indptr;   /* CA_Y_CSR: N + 1 offsets into y_indices / Y                     */
  const int32_t* y_indices;  /* CA_Y_CSR: gene index of every stored value                     */
} ca_config;

/* library / device discovery */
CA_API int ca_core_abi_version(void);
CA_API int ca_core_device_count(int* count, char* err, size_t errlen);
/* writes a fresh 128-byte ncclUniqueId (rank 0 calls this and ships it to the other ranks) */
CA_API int ca_core_nccl_unique_id(void* out128, char* err, size_t errlen);

/* Build the device state.  Y: N x G counts as described by cfg.  L: G x C copy number (already
 * saturated, R/clonealign.R:394-397).  psi_init: N x K.  loc_init: G values of
 * safe_inverse_softplus(mu_guess) (R/inference-tflow.R:262).  X: N x P or NULL.
 * colsum_total: G global column sums of Y over ALL ranks, or NULL: the library then sums the shards' column sums itself
 * (one all-reduce during set-up; with world > 1 every rank must pass NULL or every rank a value).
 * With world > 1 the call is collective; the NCCL communicator of a (world, rank, device) is built once per process and
 * re-used by later sessions of the same shape (released by ca_core_shutdown).
 * Allele inputs (V > 0): clone_allele V x C, alt and cov N x V (R/allele-specific.R:17-48). */
CA_API int ca_core_create(ca_handle** out, const ca_config* cfg, const void* Y, const double* L,
                   const double* psi_init, const double* loc_init, const double* X,
                   const double* colsum_total, const double* clone_allele, const double* alt,
                   const double* cov, char* err, size_t errlen);
CA_API int ca_core_destroy(ca_handle* h);

/* Restarts (run_clonealign, R/clonealign.R:50-56): the inputs that do not depend on the restart -- the
 * count matrix as stored in HBM and everything derived from it once (library sizes, B = Y log L, multinomial constants,
 * column sums, allele term) -- can be built once per device and shared, read-only, by any number of sessions.
 * ca_core_data_create takes the Y / L / allele arguments of ca_core_create (cfg: N, G, C, V, device, y_*; world == 1);
 * ca_core_create_shared takes the per-fit arguments (cfg: S, K, P, seed, learning_rate, path, variants, ...).
 * ca_core_data_destroy fails while sessions created from it are alive. */
CA_API int ca_core_data_create(ca_data** out, const ca_config* cfg, const void* Y, const double* L,
                        const double* colsum_total, const double* clone_allele, const double* alt, const double* cov,
                        char* err, size_t errlen);
CA_API int ca_core_data_destroy(ca_data* d, char* err, size_t errlen);
/* Initial values the host code derives from Y (R/inference-tflow.R:210 s_init = rowSums(Y), :117 colSums(Y) for the gene
 * filter, :222 mu_guess = colMeans(Y / rowMeans(Y))), computed from the resident matrix in fp64 instead of by passes over
 * the N x G host matrix.  Any output may be NULL.  rowsum: N, colsum: G, mu_guess: G. */
CA_API int ca_core_data_stats(ca_data* d, double* rowsum, double* colsum, double* mu_guess, char* err, size_t errlen);
/* rowSums(Y[, keep]) over the resident matrix: the cell filter of preprocess_for_clonealign (R/preprocess.R:138-139) counts
 * only the genes that survived its gene filters (:114-135, which need colSums(Y) from ca_core_data_stats and the copy-number
 * matrix only).  gene_keep: G bytes (non-zero = retained); rowsum: N doubles.  SURVEY.md 8f-2. */
CA_API int ca_core_data_masked_rowsums(ca_data* d, const uint8_t* gene_keep, double* rowsum, char* err, size_t errlen);
CA_API int ca_core_create_shared(ca_handle** out, const ca_config* cfg, ca_data* data, const double* psi_init,
                          const double* loc_init, const double* X, char* err, size_t errlen);

/* One pass over the shared count matrix for `n` sessions created from the same ca_data (K + P == 1): computes every
 * session's Y W and Y^T psi partial sums together (kernels_ypass.cuh, k_ypass_k1_multi), so that their next
 * ca_core_step / ca_core_elbo calls skip their own pass.  Call it when all of them have new parameters (after their
 * ca_core_step calls, before their ca_core_elbo calls): the restarts of run_clonealign then read Y once per iteration
 * instead of once per restart and iteration.  Stream-ordered with respect to every session in `hs`. */
CA_API int ca_core_ypass_many(ca_handle* const* hs, int32_t n, char* err, size_t errlen);

/* the session operations */
CA_API int ca_core_init_gamma(ca_handle* h, char* err, size_t errlen);
CA_API int ca_core_step(ca_handle* h, char* err, size_t errlen);
CA_API int ca_core_elbo(ca_handle* h, double* elbo, char* err, size_t errlen);
/* n ELBO evaluations with fresh draws (the 20 behind final_elbo / sd_final_elbo, R/inference-tflow.R:447-449) queued on
 * the stream and fetched with one device-to-host copy; elbo: n doubles, the same values n calls of ca_core_elbo give.
 * Collective under cell sharding like ca_core_elbo (every rank passes the same n). */
CA_API int ca_core_elbo_many(ca_handle* h, int32_t n, double* elbo, char* err, size_t errlen);
/* any output pointer may be NULL.  mu: G, clone_probs: N x C, s: N, alpha: C, psi: N x K,
 * W: G x K, chi: K, beta: G x P, clone_probs_from_snv: N x C (only when V > 0). */
CA_API int ca_core_params(ca_handle* h, double* mu, double* clone_probs, double* s, double* alpha,
                   double* psi, double* W, double* chi, double* beta, double* clone_probs_from_snv,
                   char* err, size_t errlen);

/* Test / parity hooks ("fixed MC draws"): feed the next eps draws from the host (S x G floats,
 * sample-major: eps[s*G + g]); n_draws consecutive draws are queued and consumed one per
 * init_gamma / step / elbo call; when the queue is empty the device generator is used. */
CA_API int ca_core_set_eps(ca_handle* h, const float* eps, int64_t n_draws, char* err, size_t errlen);
/* eps used by the most recent init_gamma / step / elbo call (S x G floats) */
CA_API int ca_core_get_eps(ca_handle* h, float* eps, char* err, size_t errlen);
/* gradients of the ELBO without the Adam update (same draw consumption as ca_core_step) */
CA_API int ca_core_grads(ca_handle* h, char* err, size_t errlen);
/* read / write a named device array as column-major doubles.  Names: W chi_raw psi beta
 * alpha_unconstr loc lsd gamma_logits (trainable); grad_<name>; mu_samples(S x G, sample-major)
 * Z(N x S*C) R(N x S*C) dM(G x S*C) F(N x C) YV(N x (K+P)) YtU(G x (K+P)) B(N x C) v(N x C)
 * const(N) colsum(G) shift(N).  `n` is the capacity / length in elements. */
CA_API int ca_core_get_array(ca_handle* h, const char* name, double* out, int64_t n, char* err, size_t errlen);
CA_API int ca_core_set_array(ca_handle* h, const char* name, const double* in, int64_t n, char* err, size_t errlen);

/* Post-hoc per-gene Pearson correlation between expression and the copy number of each cell's assigned clone
 * (compute_correlations, R/clonealign.R:318-334, called at :292-294) on the Y already resident in HBM.
 * clone_idx: N clone indices (0-based), < 0 = "unassigned" (excluded).  L: G x C copy number (column-major; the
 * reference passes the UNsaturated matrix) or NULL for the session's saturated one.  out: G values, NaN where R gives NA.
 * With world > 1 it is a collective call (clone_idx: this rank's cells; every rank gets the same G values). */
CA_API int ca_core_correlations(ca_handle* h, const int32_t* clone_idx, const double* L, double* out, char* err, size_t errlen);

/* psi initialisation on the device: scores of the leading principal component of the centred, scaled log2(Y + 1)
 * (prcomp(log2(Y_dat + 1), center = TRUE, scale = TRUE)$x[, 1], R/inference-tflow.R:203-204; K == 1) by power iteration
 * on the resident Y.  scores: N values (sign: the loading of largest magnitude is positive); the caller applies
 * scale() and adds its own N(0, 0.05^2) noise (:205-207) and writes the result with ca_core_set_array("psi").
 * Stops when 1 - |<v_new, v_old>| < tol or after max_iter iterations; *iters = iterations used.  With world > 1 it is a
 * collective call (column statistics and X^T t are all-reduced; every rank gets the scores of its own cells). */
CA_API int ca_core_pca_scores(ca_handle* h, int32_t max_iter, double tol, double* scores, int32_t* iters, char* err, size_t errlen);

/* Variant P2P (world > 1): every rank exports the 64-byte CUDA IPC handle of its exchange buffer, the caller gathers
 * the `world` handles in rank order (any transport) and hands them to every rank; from then on the per-step all-reduce of
 * the gene-level gradient partials runs as one kernel over peer memory (kernels_p2p.cuh).  Collective calls. */
CA_API int ca_core_p2p_export(ca_handle* h, void* handle64, char* err, size_t errlen);
CA_API int ca_core_p2p_connect(ca_handle* h, const void* handles, char* err, size_t errlen);
/* the same inside one process (ca_core_multi_* does this itself): the exchange buffers as plain device pointers */
CA_API int ca_core_p2p_base(ca_handle* h, void** base, char* err, size_t errlen);
CA_API int ca_core_p2p_connect_ptrs(ca_handle* h, void* const* bases, const int32_t* devices, char* err, size_t errlen);

/* Measurement hooks (bench.py): run n_steps train steps (and, if with_eval != 0, one ELBO
 * evaluation after each, as the reference loop does) back to back on the handle's stream,
 * bracketed by CUDA events on that stream; *ms = elapsed milliseconds. */
CA_API int ca_core_time_steps(ca_handle* h, int32_t n_steps, int32_t with_eval, double* ms, char* err, size_t errlen);
/* per-kernel device time of ONE train step, measured with events around every launch.
 * names: buffer for a ';'-separated list of kernel labels; ms[i] their durations; *n_k count. */
CA_API int ca_core_profile_step(ca_handle* h, char* names, size_t names_len, double* ms, int32_t cap, int32_t* n_k,
                         char* err, size_t errlen);
/* static facts for the roofline: bytes of Y as stored, kernels per step, path in use ... as a JSON string */
CA_API int ca_core_describe(ca_handle* h, char* json, size_t json_len);

/* destroys the NCCL communicators parked by destroyed sessions (optional; process exit does the same) */
CA_API int ca_core_shutdown(void);

/* ------------------------------------------------------------------------------------------------------------------
 * ONE fit, cells sharded over several GPUs, driven from ONE host thread (the R interpreter is single-threaded; SURVEY.md
 * section 8b "Threading", 8e): the library owns one worker thread per device (they never touch R), builds the
 * communicator itself and mirrors the session lifecycle of R/inference-tflow.R:351-457 over all shards.  Rows of Y,
 * psi_init, X, alt and cov are split into contiguous, balanced blocks (the first N %% n devices get one more cell);
 * per-cell state never leaves its GPU, gene-level gradients are summed once per step (one all-reduce per train step).
 * cfg->N is the TOTAL number of cells; cfg->rank / world / nccl_id / device / N_total are ignored; Y must be in HOST
 * memory (any ca_y_layout / ca_y_dtype).  Results equal the torchrun path (one process per GPU) bit for bit.
 * ca_core_multi_params fills the arrays of ca_core_params for ALL cells.  ca_core_multi_shard lends shard i (for the
 * test / measurement hooks of a single shard; collective calls on it must be issued for every shard). */
typedef struct ca_multi ca_multi;
CA_API int ca_core_multi_create(ca_multi** out, const ca_config* cfg, const int32_t* devices, int32_t n_devices, const void* Y,
                         const double* L, const double* psi_init, const double* loc_init, const double* X,
                         const double* clone_allele, const double* alt, const double* cov, char* err, size_t errlen);
CA_API int ca_core_multi_destroy(ca_multi* m);
CA_API int ca_core_multi_init_gamma(ca_multi* m, char* err, size_t errlen);
CA_API int ca_core_multi_step(ca_multi* m, char* err, size_t errlen);
CA_API int ca_core_multi_elbo(ca_multi* m, double* elbo, char* err, size_t errlen);
CA_API int ca_core_multi_elbo_many(ca_multi* m, int32_t n, double* elbo, char* err, size_t errlen);
CA_API int ca_core_multi_params(ca_multi* m, double* mu, double* clone_probs, double* s, double* alpha, double* psi,
                         double* W, double* chi, double* beta, double* clone_probs_from_snv, char* err, size_t errlen);
CA_API int ca_core_multi_time_steps(ca_multi* m, int32_t n_steps, int32_t with_eval, double* ms, char* err, size_t errlen);
CA_API int ca_core_multi_shard(ca_multi* m, int32_t i, ca_handle** out, int64_t* row_begin, int64_t* row_end);
CA_API int ca_core_multi_size(ca_multi* m);

#ifdef __cplusplus
}
#endif
#endif /* CLONEALIGN_B200_H */
