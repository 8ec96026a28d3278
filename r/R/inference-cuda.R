# Replacement for the session-dependent part of R/inference-tflow.R (reference lines 238-457).
# NOT RUN IN THIS REPOSITORY (no R here).  Everything before line 238 of the reference (gene filter, saturate,
# PCA psi init, s_init, mu_guess) and everything after line 459 (convergence_info, naming, return list) stays
# byte-for-byte; only the graph construction and the sess$run calls are replaced by .Call()s into src/ca_shim.c.
# The Python mirror clonealign_b200/inference.py is the executable twin of this file and is what the tests cover.

# Optional (SURVEY.md 8f; each falls back to the reference's host code when left at its default):
#   counts_dgC  : the genes x cells dgCMatrix of the (gene-filtered) counts assay; passed to the device compressed
#                 (ca_create_sparse) instead of Y_dat = t(as.matrix(assay(...))) (R/clonealign.R:217)
#   device_pca  : pcs (R/inference-tflow.R:203-205) from ca_pca_scores on the resident Y; scale() and the rnorm noise
#                 (:205-207) stay in R, so set.seed() governs the same draws (pass pcs = NULL)
#   batch_final_elbo : the 20 evaluations behind final_elbo (:447-449) with one device-to-host copy (ca_elbo_many); same values
#   cor_with    : list(L = unsaturated copy number of the retained genes, p = clone_call_probability): run
#                 compute_correlations (R/clonealign.R:292-294,318-334) on the device before the session closes
inference_cuda_session <- function(Y_dat, L_dat, pcs, mu_guess, x, clone_allele, alt, cov,
                                   learning_rate, K, mc_samples, max_iter, rel_tol, verbose,
                                   counts_dgC = NULL, device_pca = FALSE, cor_with = NULL, batch_final_elbo = FALSE) {
  N <- if (is.null(counts_dgC)) nrow(Y_dat) else ncol(counts_dgC)
  G <- if (is.null(counts_dgC)) ncol(Y_dat) else nrow(counts_dgC)
  C <- ncol(L_dat)
  pca_noise <- NULL
  if (device_pca && K == 1) {                              # rnorm drawn where the reference draws it (:207)
    pca_noise <- matrix(rnorm(N, mean = 0, sd = .05), nrow = N)
    pcs <- matrix(0, N, 1)
  }
  P <- if (is.null(x)) 0L else ncol(x)
  V <- if (is.null(clone_allele)) 0L else nrow(clone_allele)
  if (K == 0) x <- NULL                                    # reference quirk: covariates ignored without latent dims (:279-285)
  storage.mode(L_dat) <- "double"
  alt_nv <- if (V > 0) t(alt) else NULL; cov_nv <- if (V > 0) t(cov) else NULL   # :177-180 transposes undone: ABI takes N x V
  # options(clonealign.gpus = c(0, 1, ...)): shard the cells of this fit over several GPUs (SURVEY.md 8e).  R stays
  # single-threaded: the library owns one worker thread per device (ca_core_multi_*); same calls, same results.
  gpus <- as.integer(getOption("clonealign.gpus", 0L))
  multi <- length(gpus) > 1
  if (multi && (!is.null(counts_dgC) || device_pca || !is.null(cor_with)))
    stop("options(clonealign.gpus) with several devices supports dense input without the device-side extras")
  cc <- function(name, ...) .Call(if (multi) sub("^ca_", "ca_multi_", name) else name, ...)
  sess <- if (multi) {
    .Call("ca_multi_create", Y_dat, L_dat, pcs, safe_inverse_softplus(mu_guess), x, clone_allele, alt_nv, cov_nv,
          as.integer(mc_samples), as.integer(K), learning_rate, get_next_seed(), gpus)
  } else if (is.null(counts_dgC)) {
    .Call("ca_create", Y_dat, L_dat, pcs, safe_inverse_softplus(mu_guess), x, clone_allele, alt_nv, cov_nv,
          as.integer(mc_samples), as.integer(K), learning_rate, get_next_seed(), gpus[1])
  } else {
    .Call("ca_create_sparse", counts_dgC@Dim, counts_dgC@p, counts_dgC@i, counts_dgC@x, L_dat, pcs,
          safe_inverse_softplus(mu_guess), x, clone_allele, alt_nv, cov_nv,
          as.integer(mc_samples), as.integer(K), learning_rate, get_next_seed(), gpus[1])
  }
  on.exit(cc("ca_destroy", sess), add = TRUE)              # sess$close(), :457
  if (!is.null(pca_noise)) {
    pcs <- scale(matrix(.Call("ca_pca_scores", sess, N, 500L, 1e-12), nrow = N))   # :203-205
    .Call("ca_set_psi", sess, pcs + pca_noise)                                     # :207
  }

  cc("ca_init_gamma", sess)                                # :368-369
  elbo_val <- cc("ca_elbo", sess)                          # :372
  if (is.na(elbo_val)) stop("Initial elbo is NA")          # :374-376

  elbo_diffs <- rep(1e3, 10); elbos <- elbo_val            # :379-380
  for (i in seq_len(max_iter)) {                           # :394
    cc("ca_step", sess)                                    # :401
    elbo_new <- cc("ca_elbo", sess)                        # :403
    elbo_diff <- (elbo_new - elbo_val) / abs(elbo_val)
    elbo_diffs <- c(elbo_diffs[-1], elbo_diff)
    elbos <- c(elbos, elbo_new); elbo_val <- elbo_new
    if (mean(abs(elbo_diffs)) < rel_tol) break             # :414
  }
  rlist <- cc("ca_params", sess, c(N, G, C, as.integer(K), P, V))      # :424-440
  final_elbo <- if (batch_final_elbo) cc("ca_elbo_many", sess, 20L)      # 20 fresh-draw evaluations, one round trip
                else replicate(20, cc("ca_elbo", sess))                  # :447-449
  correlations <- NULL
  if (!is.null(cor_with)) {                                # clone_assignment (:22-29) as 0-based indices, -1 = unassigned
    cp <- rlist$clone_probs
    idx <- ifelse(apply(cp, 1, max) < cor_with$p, -1L, max.col(cp, ties.method = "first") - 1L)
    correlations <- .Call("ca_correlations", sess, as.integer(idx), cor_with$L, G)
  }
  list(rlist = rlist, elbos = elbos, final_elbo = final_elbo, correlations = correlations)
}
