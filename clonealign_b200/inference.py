"""Host-side mirror of the reference's `inference_tflow()` (R/inference-tflow.R:71-481).

Same arguments, same order of operations, same return structure; the TensorFlow session is
replaced by `clonealign_b200.session.Session` (CUDA).  The optimisation loop, convergence test and
messages stay on the host exactly as in the reference (:378-417) so interrupt / progress semantics
do not change.  All randomness derives from one host RNG (`seed`), as the reference derives it from
R's RNG (tests/testthat/test_clonealign.R:42-66).
"""
from __future__ import annotations

import sys

import numpy as np

from .session import Session


def inverse_softplus(x):
    """R/inference-tflow.R:2-4."""
    return np.log(np.exp(x) - 1.0)


def safe_inverse_softplus(x):
    """R/inference-tflow.R:6-11."""
    x = np.asarray(x, dtype=np.float64)
    if np.any(x < 0):
        raise ValueError("Inverse softplus only takes positive values")
    return np.log(1.0 - np.exp(-np.abs(x))) + np.maximum(x, 0.0)


def softplus(x):
    """R/inference-tflow.R:13-15."""
    return np.logaddexp(0.0, x)


def saturate(x, threshold=4):
    """R/clonealign.R:394-397."""
    x = np.array(x, dtype=np.float64, copy=True)
    x[x > threshold] = threshold
    return x


def clone_assignment(gamma, clone_names, clone_assignment_probability=0.95):
    """R/inference-tflow.R:22-29."""
    gamma = np.asarray(gamma)
    mx = gamma.max(axis=1)
    am = gamma.argmax(axis=1)
    return [("unassigned" if mx[i] < clone_assignment_probability else clone_names[am[i]]) for i in range(len(mx))]


def get_next_seed(rng):
    """R/inference-tflow.R:49-51 — sample(.Machine$integer.max - 1, 1)."""
    return int(rng.integers(1, 2 ** 31 - 1))


def _message(verbose, msg):
    if verbose:
        print(msg, file=sys.stderr)


def pca_init(Y, K, rng, truncated=None):
    """psi initialisation, R/inference-tflow.R:204-208: prcomp(log2(Y+1), center, scale)$x[,1:K], scale(), + N(0,.05^2)."""
    pcs = pca_scaled(Y, K, truncated)
    return pcs + rng.normal(0.0, 0.05, size=pcs.shape)


def pca_scaled(Y, K, truncated=None):
    """scale(prcomp(log2(Y+1), center, scale)$x[,1:K]) (R/inference-tflow.R:204-206), without the noise of :207.

    The reference runs a full `prcomp` (O(N G^2)); only the K leading components are used, so large inputs go through
    a truncated Lanczos SVD (same components up to sign, which the model does not see: W starts at 0)."""
    Y = np.asarray(Y)
    big = (Y.shape[0] * Y.shape[1] > (1 << 24)) if truncated is None else bool(truncated)
    X = np.log2(np.asarray(Y, dtype=np.float32 if big else np.float64) + 1.0)
    X -= X.mean(axis=0)
    sd = X.std(axis=0, ddof=1)
    if np.any(sd == 0):
        raise ValueError("cannot rescale a constant/zero column to unit variance")
    X /= sd
    if big:
        # only the K leading principal components are used (:205): Lanczos on the standardised matrix
        from scipy.sparse.linalg import svds
        k = max(1, K)
        U, S, _ = svds(X, k=k, which="LM", random_state=np.random.RandomState(0))
        order = np.argsort(-S)
        pcs = (U[:, order] * S[order]).astype(np.float64)[:, :K]
    else:
        U, S, _ = np.linalg.svd(X, full_matrices=False)
        pcs = (U * S)[:, :K]
    return (pcs - pcs.mean(axis=0)) / pcs.std(axis=0, ddof=1)


def inference_tflow(*args, **kwargs):
    """CUDA-backed equivalent of the reference's `inference_tflow` (R/inference-tflow.R:71-481); arguments and return
    value: see `inference_steps`, which this function simply runs to completion."""
    gen = inference_steps(*args, **kwargs)
    try:
        while True:
            next(gen)
    except StopIteration as done:
        return done.value


def inference_steps(Y_dat, L_dat, max_iter=100, rel_tol=1e-5, learning_rate=0.1, gene_filter_threshold=0,
                    x=None, clone_allele=None, cov=None, ref=None, fix_alpha=False, dtype="float32",
                    saturate_=True, saturation_threshold=6, K=1, mc_samples=1, verbose=True, initial_shrink=5,
                    data_init_mu=True, seed=None, device=0, psi_init=None, y_store="auto", path="auto",
                    gene_names=None, variants=None, correlations_with=None, device_pca=False, cache=None,
                    device_stats=False, batch_final_elbo=False, devices=None):
    """The fit of `inference_tflow` as a generator: it yields its `Session` every time the parameters have just changed
    and the next operation is an ELBO evaluation (after gamma-init and after every train step), i.e. exactly where one
    batched pass over a shared count matrix can serve several restarts (`session.ypass_many`; `run_clonealign(
    batch_y_pass=True)` advances the restarts of a device in lock-step through these points).  Ignoring the yields gives
    the plain fit.  The generator's return value is the reference's result list:

    Y_dat: cell x gene counts; L_dat: gene x clone copy number.  Returns the reference's list as a dict:
    ml_params {mu, clone_probs, s, alpha [, beta] [, psi, W, chi]}, convergence_info {final_elbo,
    sd_final_elbo, elbo}, retained_genes, clone_probs_from_snv (R/inference-tflow.R:475-480).
    `fix_alpha` and `initial_shrink` are accepted and ignored, as in the reference (:81,:88).
    `psi_init` (N x K) skips the PCA; it exists for callers that compute the initialisation elsewhere.
    `device_pca=True` (K == 1): the leading principal component that initialises psi (:203-205) is computed by power
    iteration on the count matrix resident in HBM (ca_core_pca_scores) instead of a host-side SVD; scale() and the
    N(0, 0.05^2) noise (:205-207) stay on the host and use the same RNG stream positions as the host path.
    `cache` (a dict owned by the caller, e.g. `run_clonealign`): restart-independent work is done once and reused by later
    calls WITH IDENTICAL INPUTS -- the un-noised principal components, and per device the count matrix resident in HBM with
    everything derived from it (`DeviceData`, ca_core_data_create); the RNG stream positions do not change, so a cached
    restart is bit-identical to an uncached one.  The caller closes `cache[("data", device)]` when done.
    `device_stats=True` (needs `cache`): s_init = rowSums(Y) (:210) and mu_guess = colMeans(Y / rowMeans(Y)) (:222) come from
    the resident matrix (ca_core_data_stats, fp64) instead of host passes over the N x G matrix; mu_guess then agrees with
    the host value to ~1e-15 relative (different summation order), so fits are no longer bit-identical to host-initialised ones.
    `batch_final_elbo=True`: the 20 fresh-draw ELBO evaluations behind final_elbo / sd_final_elbo (:447-449) are queued on the
    stream and fetched with one device-to-host copy (ca_core_elbo_many): same draws, bit-identical values; opt-in like the other
    entry points that have not run on hardware yet.
    `devices=[0, 1, ...]` (the R twin: `options(clonealign.gpus = ...)`): the cells of THIS fit are sharded over several GPUs of
    this process (`MultiSession`, ca_core_multi_*: one worker thread per device inside the library, one all-reduce of the
    gene-level gradients per step); dense or sparse host input; not combined with the device-side extras above.
    `correlations_with = (L_unsaturated, clone_call_probability)`: also run the caller's post-hoc
    `compute_correlations` (R/clonealign.R:292-294,318-334) on the device while Y is still resident; the result is
    returned under "correlations" (retained genes only).
    """
    if dtype not in ("float32", "float64"):
        raise ValueError("'arg' should be one of 'float32', 'float64'")
    if dtype == "float64":
        raise ValueError("dtype='float64' is broken in the reference graph (tf$to_float, "
                         "R/inference-tflow.R:323); only float32 is supported")
    rng = np.random.default_rng(seed)
    _message(verbose, "Constructing CUDA session")

    # a scipy.sparse cells x genes matrix (the transposed dgCMatrix of a SingleCellExperiment) stays compressed all the
    # way to the device (CA_Y_CSR) instead of being densified as t(as.matrix(assay(...))) does (R/clonealign.R:217)
    sparse = hasattr(Y_dat, "tocsr") and hasattr(Y_dat, "nnz")
    Y_dat = Y_dat.tocsr() if sparse else np.asarray(Y_dat)
    L_dat = np.asarray(L_dat, dtype=np.float64)
    if Y_dat.ndim != 2 or L_dat.ndim != 2 or L_dat.shape[0] != Y_dat.shape[1]:
        raise ValueError("nrow(L_dat) == G is not TRUE")                              # :139 (R fails in the subset at :124)
    if cache is not None:                                                             # restarts share identical inputs:
        import contextlib                                                             # one pass over the matrix, not one per fit
        with cache.get("lock") or contextlib.nullcontext():
            zero_gene_means = cache.get("gene_filter")
            if zero_gene_means is None:
                zero_gene_means = cache["gene_filter"] = np.asarray(Y_dat.sum(axis=0)).ravel() <= gene_filter_threshold
    else:
        zero_gene_means = np.asarray(Y_dat.sum(axis=0)).ravel() <= gene_filter_threshold  # :117
    _message(verbose, f"Removing {int(zero_gene_means.sum())} genes with low counts")  # :120
    if zero_gene_means.any():
        Y = Y_dat[:, ~zero_gene_means]
        L = L_dat[~zero_gene_means, :]
    else:                       # nothing to drop: no copy of the N x G matrix (16 GB of doubles at 100k x 20k)
        Y, L = Y_dat, L_dat
    if sparse and psi_init is None and not (device_pca and K == 1):
        Y, sparse = np.asarray(Y.todense()), False     # the host PCA needs the dense matrix (use device_pca=True to avoid it)
    if gene_names is not None:
        retained_genes = [g for g, z in zip(gene_names, zero_gene_means) if not z]    # :127-128
    else:
        retained_genes = (np.nonzero(~zero_gene_means)[0] + 1).tolist()               # which(), 1-based :130
    N, G = Y.shape
    if L.shape[0] != G:
        raise ValueError("nrow(L_dat) == G is not TRUE")                              # :139
    K = int(K)
    if saturate_:
        L = saturate(L, saturation_threshold)                                         # :142-144
    if x is not None:                                                                 # :147-153
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        if x.shape[0] != N:
            raise ValueError("nrow(x) == N is not TRUE")

    if K == 0:
        x = None   # reference quirk (:279-285): without latent dimensions the covariates never enter the graph
    use_allele = clone_allele is not None and ref is not None and cov is not None     # :167
    alt = None
    if use_allele:
        _message(verbose, "Using allelic imbalance info")
        clone_allele = np.asarray(clone_allele, dtype=np.float64)
        cov = np.asarray(cov, dtype=np.float64)
        ref = np.asarray(ref, dtype=np.float64)
        V = clone_allele.shape[0]
        # sanitize_allele_info, R/allele-specific.R:61-71
        if clone_allele.shape[1] != L.shape[1] or cov.shape != (N, V) or ref.shape != (N, V):
            raise ValueError("allele inputs have inconsistent dimensions")
        alt = cov - ref                                                               # :180

    pca_noise = None
    if psi_init is None:
        if device_pca and K == 1:
            pca_noise = rng.normal(0.0, 0.05, size=(N, 1))                            # :207 (same stream position)
            psi_init = np.zeros((N, 1))                                               # replaced below, before any use
        elif K > 0:
            if cache is None:
                pcs = pca_scaled(Y, K)                                                # :204-206
            else:
                import contextlib
                with cache.get("lock") or contextlib.nullcontext():                   # restarts on several devices: compute once
                    pcs = cache.get("pcs")
                    if pcs is None:
                        pcs = cache["pcs"] = pca_scaled(Y, K)
            psi_init = pcs + rng.normal(0.0, 0.05, size=pcs.shape)                    # :207
        else:
            psi_init = np.zeros((N, 0))
    def get_data():
        """Restart-independent device inputs: upload + preprocess once per device (built on first use)."""
        import contextlib
        from .session import DeviceData
        # concurrent restarts on one device build the inputs once; different devices upload concurrently (a lock per device,
        # the shared lock only guards the dictionary)
        import threading
        with cache.get("lock") or contextlib.nullcontext():
            dev_lock = cache.setdefault(("lock", device), threading.Lock())
        with dev_lock:
            dd = cache.get(("data", device))
            if dd is None:
                dd = cache[("data", device)] = DeviceData(Y, L, device=device, clone_allele=clone_allele if use_allele else None,
                                                          alt=alt, cov=cov if use_allele else None, y_store=y_store)
        return dd

    dev_stats = None
    if device_stats:
        if cache is None:
            raise ValueError("device_stats=True needs a cache (the statistics come from the shared device inputs)")
        dev_stats = cache.get("stats")
        if dev_stats is None:
            dev_stats = cache["stats"] = get_data().stats()
    s_init = dev_stats["rowsum"] if dev_stats is not None else \
        np.asarray(Y.sum(axis=1), dtype=np.float64).ravel()                           # :210
    if np.any(s_init == 0):
        raise ValueError("Some cells have no counts mapping")                         # :212-214
    if isinstance(data_init_mu, (bool, np.bool_)):                                    # :220-235
        if data_init_mu and dev_stats is not None:
            mu_guess = dev_stats["mu_guess"]
        elif data_init_mu and sparse:
            import scipy.sparse as sp
            inv_rowmean = sp.diags(Y.shape[1] / s_init)                               # 1 / rowMeans(Y)
            mu_guess = np.asarray((inv_rowmean @ Y.astype(np.float64)).mean(axis=0)).ravel()
        elif data_init_mu:
            Yd = np.asarray(Y, dtype=np.float64)
            mu_guess = (Yd / Yd.mean(axis=1, keepdims=True)).mean(axis=0)
        else:
            mu_guess = np.ones(G)
    else:
        _message(verbose, "Using user-provided mu values to start")
        d = np.asarray(data_init_mu, dtype=np.float64)
        mu_guess = d / d.mean()

    op_seed = get_next_seed(rng)                                                      # :269
    multi = devices is not None and len(list(devices)) > 1
    if multi and (device_pca or device_stats or cache is not None or correlations_with is not None):
        raise ValueError("devices=[...] (one fit over several GPUs) cannot be combined with device_pca, device_stats, "
                         "shared restart inputs or device correlations")
    data = get_data() if cache is not None else None
    if multi:
        from .session import MultiSession
        sess = MultiSession(Y, L, psi_init, safe_inverse_softplus(mu_guess), devices=list(devices), mc_samples=int(mc_samples),
                            K=K, x=x, learning_rate=learning_rate, seed=op_seed,
                            clone_allele=clone_allele if use_allele else None, alt=alt, cov=cov if use_allele else None,
                            y_store=y_store, path=path, variants=variants)
    else:
        if devices is not None and len(list(devices)) == 1:
            device = list(devices)[0]
        sess = Session(Y, L, psi_init, safe_inverse_softplus(mu_guess), mc_samples=int(mc_samples), K=K, x=x,
                       learning_rate=learning_rate, seed=op_seed, device=device,
                       clone_allele=clone_allele if use_allele else None, alt=alt, cov=cov if use_allele else None,
                       y_store=y_store, path=path, variants=variants, data=data)
    correlations = None
    try:
        if pca_noise is not None:
            pcs = cache.get("pcs_device") if cache is not None else None
            if pcs is None:
                pcs, n_it = sess.pca_scores(max_iter=500)                             # :203-204 on the device
                if n_it >= 500:
                    import warnings
                    warnings.warn("device PCA: power iteration stopped at 500 iterations before reaching its tolerance "
                                  "(leading components nearly degenerate); the psi initialisation is approximate")
                pcs = (pcs - pcs.mean()) / pcs.std(ddof=1)                            # scale(pcs), :205
                if cache is not None:
                    cache["pcs_device"] = pcs
            sess.set_array("psi", pcs[:, None] + pca_noise)
        sess.init_gamma()                                                             # :368-369
        yield sess
        elbo_val = sess.elbo()                                                        # :372
        if np.isnan(elbo_val):
            raise ValueError("Initial elbo is NA")                                    # :374-376
        elbo_diffs = [1e3] * 10                                                       # :379
        elbos = [elbo_val]
        _message(verbose, "Optimizing ELBO")
        for _ in range(int(max_iter)):                                                # :394
            sess.step()                                                               # :401
            yield sess
            elbo_new = sess.elbo()                                                    # :403
            if np.isnan(elbo_new):
                # the reference's `if(mean(abs(elbo_diffs)) < rel_tol)` (:414) stops with "missing value where TRUE/FALSE
                # needed" once the ELBO is NA; a diverged fit must not run on to max_iter and be returned as a result
                raise ValueError(f"ELBO is NA after iteration {len(elbos)}: missing value where TRUE/FALSE needed")
            elbo_diff = (elbo_new - elbo_val) / abs(elbo_val)
            elbo_diffs = elbo_diffs[1:] + [elbo_diff]
            elbos.append(elbo_new)
            elbo_val = elbo_new
            if np.mean(np.abs(elbo_diffs)) < rel_tol:                                 # :414
                break
        _message(verbose, "\nELBO converged or reached max iterations")
        rlist = sess.params()                                                         # :424-434
        clone_probs_from_snv = rlist.pop("clone_probs_from_snv", None)                # :436-440
        _message(verbose, "Computing final ELBO")
        if batch_final_elbo:
            final_elbo = [float(e) for e in sess.elbo_many(20)]                       # :447-449, one host round trip
        else:
            final_elbo = [sess.elbo() for _ in range(20)]                             # :447-449
        if correlations_with is not None:
            L_full, call_p = correlations_with
            cp = rlist["clone_probs"]
            zidx = np.where(cp.max(axis=1) < call_p, -1, cp.argmax(axis=1)).astype(np.int32)   # clone_assignment, :22-29
            correlations = sess.correlations(zidx, np.asarray(L_full, dtype=np.float64)[~zero_gene_means])
    finally:
        sess.close()                                                                  # :457
    convergence_info = {"final_elbo": float(np.mean(final_elbo)),
                        "sd_final_elbo": float(np.std(final_elbo, ddof=1)),
                        "elbo": np.array(elbos)}
    out = {"ml_params": rlist, "convergence_info": convergence_info, "retained_genes": retained_genes,
           "clone_probs_from_snv": clone_probs_from_snv, "retained_mask": ~zero_gene_means}
    if correlations is not None:
        out["correlations"] = correlations
    return out
