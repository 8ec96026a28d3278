# Replacement for the session-dependent part of R/inference-tflow.R (reference lines 238-457).
# NOT RUN IN THIS REPOSITORY (no R here).  Everything before line 238 of the reference (gene filter, saturate,
# PCA psi init, s_init, mu_guess) and everything after line 459 (convergence_info, naming, return list) stays
# byte-for-byte; only the graph construction and the sess$run calls are replaced by .Call()s into src/ca_shim.c.
# The Python mirror clonealign_b200/inference.py is the executable twin of this file and is what the tests cover.

inference_cuda_session <- function(Y_dat, L_dat, pcs, mu_guess, x, clone_allele, alt, cov,
                                   learning_rate, K, mc_samples, max_iter, rel_tol, verbose) {
  N <- nrow(Y_dat); G <- ncol(Y_dat); C <- ncol(L_dat)
  P <- if (is.null(x)) 0L else ncol(x)
  V <- if (is.null(clone_allele)) 0L else nrow(clone_allele)
  if (K == 0) x <- NULL                                    # reference quirk: covariates ignored without latent dims (:279-285)
  storage.mode(L_dat) <- "double"
  sess <- .Call("ca_create", Y_dat, L_dat, pcs, safe_inverse_softplus(mu_guess), x,
                clone_allele, if (V > 0) t(alt) else NULL, if (V > 0) t(cov) else NULL,   # :177-180 transposes undone: ABI takes N x V
                as.integer(mc_samples), as.integer(K), learning_rate, get_next_seed(), 0L)
  on.exit(.Call("ca_destroy", sess), add = TRUE)           # sess$close(), :457

  .Call("ca_init_gamma", sess)                             # :368-369
  elbo_val <- .Call("ca_elbo", sess)                       # :372
  if (is.na(elbo_val)) stop("Initial elbo is NA")          # :374-376

  elbo_diffs <- rep(1e3, 10); elbos <- elbo_val            # :379-380
  for (i in seq_len(max_iter)) {                           # :394
    .Call("ca_step", sess)                                 # :401
    elbo_new <- .Call("ca_elbo", sess)                     # :403
    elbo_diff <- (elbo_new - elbo_val) / abs(elbo_val)
    elbo_diffs <- c(elbo_diffs[-1], elbo_diff)
    elbos <- c(elbos, elbo_new); elbo_val <- elbo_new
    if (mean(abs(elbo_diffs)) < rel_tol) break             # :414
  }
  rlist <- .Call("ca_params", sess, c(N, G, C, as.integer(K), P, V))   # :424-440
  final_elbo <- replicate(20, .Call("ca_elbo", sess))      # :447-449
  list(rlist = rlist, elbos = elbos, final_elbo = final_elbo)
}
