"""`clonealign()` / `run_clonealign()` host mirrors (R/clonealign.R:35-75, 184-305).

Only what sits directly on either side of the hot path is mirrored here: input parsing into the
cell x gene matrix, the call into `inference_tflow`, clone calling, the post-hoc correlations and the
best-of-restarts selection.  Plotting, preprocessing and printing are out of scope.
"""
from __future__ import annotations

import warnings

import numpy as np

from .inference import clone_assignment, inference_steps, inference_tflow


class CloneAlignFit(dict):
    """The reference's `clonealign_fit` list (R/clonealign.R:303): dict with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __repr__(self):   # print.clonealign_fit, R/clonealign.R:348-357
        N = len(self["clone"])
        G = len(self["ml_params"]["mu"])
        C = self["ml_params"]["clone_probs"].shape[1]
        return (f"A clonealign_fit for {N} cells, {G} genes, and {C} clones\n"
                "To access clone assignments, call x['clone']\n"
                "To access ML parameter estimates, call x['ml_params']")


def compute_correlations(Y, L, clones, clone_names):
    """R/clonealign.R:318-334 — per-gene Pearson correlation of scaled expression with assigned copy number."""
    clones = np.asarray(clones, dtype=object)
    keep = clones != "unassigned"
    Y = np.asarray(Y, dtype=np.float64)[keep]
    idx = np.array([clone_names.index(c) for c in clones[keep]], dtype=int)
    out = np.full(Y.shape[1], np.nan)
    if Y.shape[0] < 2:
        return out
    # cor(x, scale(y)) == cor(x, y); constant x or y gives NA (R's cor warns and returns NA), all genes at once
    X = np.asarray(L, dtype=np.float64)[:, idx].T                  # cells x genes: copy number of the assigned clone
    Yc = Y - Y.mean(axis=0)
    Xc = X - X.mean(axis=0)
    sy = np.sqrt((Yc * Yc).sum(axis=0))
    sx = np.sqrt((Xc * Xc).sum(axis=0))
    ok = (sy > 0) & (np.ptp(X, axis=0) > 0)     # a constant x is detected exactly (X - mean can leave a rounding residue)
    out[ok] = (Xc[:, ok] * Yc[:, ok]).sum(axis=0) / (sx[ok] * sy[ok])
    return out


def clonealign(*args, **kwargs):
    """Assign cells to clones (R/clonealign.R:184-305): arguments and return value as `clonealign_steps`, run to completion."""
    gen = clonealign_steps(*args, **kwargs)
    try:
        while True:
            next(gen)
    except StopIteration as done:
        return done.value


def clonealign_steps(gene_expression_data, copy_number_data, max_iter=200, rel_tol=1e-6, gene_filter_threshold=0,
               learning_rate=0.1, x=None, clone_allele=None, cov=None, ref=None, fix_alpha=False, dtype="float32",
               saturate=True, saturation_threshold=6, K=None, mc_samples=1, verbose=True, initial_shrink=5,
               clone_call_probability=0.95, data_init_mu=True, clone_names=None, gene_names=None, seed=None,
               device=0, fix_ref_bug=False, device_correlations=False, **backend):
    """Assign cells to clones (R/clonealign.R:184-305), as a generator over the synchronisation points of the fit (see
    `inference_steps`; `run_clonealign(batch_y_pass=True)` uses them, `clonealign()` ignores them).

    gene_expression_data: cell x gene count matrix (an R user passes a SingleCellExperiment whose
    counts assay is transposed to this, :212-222).  copy_number_data: gene x clone matrix.
    `fix_ref_bug=False` keeps the reference's `ref = cov` forwarding (:271) which makes the
    alternate-allele count identically zero; pass True to forward `ref` as documented.
    `device_correlations=True` computes the post-hoc gene/copy-number correlations (:292-294) with one pass over the Y
    that is already resident in HBM (ca_core_correlations) instead of the host loop; the default stays the host mirror
    until that kernel has been run on hardware (it is verified on the CPU emulation, tests/test_emul_parity.py).
    """
    sparse = hasattr(gene_expression_data, "tocsr") and hasattr(gene_expression_data, "nnz")   # scipy.sparse, cells x genes
    Y = gene_expression_data.tocsr() if sparse else np.asarray(gene_expression_data)
    if Y.ndim != 2:
        raise ValueError("Input gene_expression_data must be SingleCellExperiment, SummarizedExperiment, or matrix")
    N, G = Y.shape
    if K is None:
        K = 1                                                                           # :226-232
    L = np.asarray(copy_number_data, dtype=np.float64)
    if L.ndim != 2:
        raise ValueError("copy_number_data must be a matrix, data.frame or DataFrame.")
    if L.shape[0] != G:
        raise ValueError("copy_number_data must have same number of genes (rows) as gene_expression_data")
    C = L.shape[1]
    if clone_names is None:
        clone_names = [f"clone_{chr(ord('a') + i)}" for i in range(C)]                  # :251-253
    clone_names = list(clone_names)
    if gene_names is None:
        gene_names = [f"gene_{i + 1}" for i in range(G)]

    res = yield from inference_steps(Y, L, max_iter=max_iter, rel_tol=rel_tol, learning_rate=learning_rate,
                          gene_filter_threshold=gene_filter_threshold, x=x, clone_allele=clone_allele, cov=cov,
                          ref=(ref if fix_ref_bug else cov), fix_alpha=fix_alpha, dtype=dtype, saturate_=saturate,
                          saturation_threshold=saturation_threshold, K=K, mc_samples=mc_samples, verbose=verbose,
                          initial_shrink=initial_shrink, data_init_mu=data_init_mu, seed=seed, device=device,
                          gene_names=gene_names, correlations_with=(L, clone_call_probability) if device_correlations else None,
                          **backend)                                                    # :262-280
    dev_cor = res.pop("correlations", None)
    ridx = np.nonzero(res.pop("retained_mask"))[0]          # positions of the retained genes (names may repeat)
    fit = CloneAlignFit(res)
    fit["clone"] = clone_assignment(res["ml_params"]["clone_probs"], clone_names, clone_call_probability)   # :283
    fit["clone_names"] = clone_names
    if dev_cor is None:
        Yr = Y[:, ridx]
        dev_cor = compute_correlations(np.asarray(Yr.todense()) if sparse else Yr, L[ridx, :], fit["clone"], clone_names)
    fit["correlations"] = dev_cor                                                                           # :292-294
    cor = fit["correlations"]
    if np.any(~np.isnan(cor)) and np.nanquantile(cor, 0.25) < 0:                        # :296-300
        warnings.warn("Less than 75% of genes positively correlated with expression - assignment may have failed")
    return fit


def run_clonealign(gene_expression_data, copy_number_data, initial_shrinks=(0, 5, 10), n_repeats=3,
                   print_elbos=True, seed=None, devices=None, share_inputs=False, restarts_in_flight=1,
                   batch_y_pass=False, **kwargs):
    """Best-of-restarts wrapper (R/clonealign.R:35-75).  Restarts are independent fits; `devices`
    (list of CUDA ordinals) spreads them round-robin over GPUs (replicas only, no communication).
    `share_inputs=True`: the restarts differ only through the RNG (psi noise, op seed), so the principal components and,
    per device, the count matrix in HBM with everything derived from it are built once and shared read-only by the
    restarts on that device (`DeviceData`) instead of being recomputed / re-uploaded 9 times; the fits
    are bit-identical to unshared ones.  Off by default until it has run on hardware (verified on the CPU emulation).
    `restarts_in_flight=k`: up to k restarts of one device run concurrently (one host thread and one CUDA stream each; the
    C-ABI calls release the GIL), so the HBM-bound Y pass of one fit overlaps the issue-bound per-cell / gene kernels of
    another; every fit stays deterministic and the selection below sees them in the serial order.
    `batch_y_pass=True` (implies share_inputs): the restarts of a device advance in lock-step and ONE pass over the shared
    count matrix per iteration serves all of them (ca_core_ypass_many) instead of one pass per restart and iteration."""
    rng = np.random.default_rng(seed)
    if batch_y_pass:
        # the batched pass serves fits with ONE latent dimension and no covariates, on the 8-column tiling of ypass2
        # (core: ca_core_ypass_many); anything else runs its own pass per restart -- decided here, before any session exists
        if kwargs.get("K") not in (None, 1) or kwargs.get("x") is not None:
            warnings.warn("batch_y_pass needs K = 1 and no covariates: running one Y pass per restart instead")
            batch_y_pass, share_inputs = False, True
        else:
            from .session import variant_mask, _VARIANT
            names = kwargs.get("variants") or ""
            if variant_mask(names) & (_VARIANT["ypass3"] | _VARIANT["ypass4"]):
                raise ValueError("batch_y_pass uses the column tiling of variant ypass2; it cannot be combined with ypass3 / ypass4")
            if not (variant_mask(names) & _VARIANT["ypass2"]):
                kwargs["variants"] = ",".join([v for v in (names.split(",") if isinstance(names, str) else list(names)) if v] + ["ypass2"])
    jobs = []
    for is_ in initial_shrinks:
        for _ in range(n_repeats):
            dev = devices[len(jobs) % len(devices)] if devices else kwargs.get("device", 0)
            kw = dict(kwargs)
            kw.update(initial_shrink=is_, seed=int(rng.integers(0, 2 ** 31 - 1)), device=dev)   # seeds fixed up front
            jobs.append(kw)
    cache = None
    if share_inputs or batch_y_pass:
        import threading
        cache = {"lock": threading.Lock()}
        for kw in jobs:
            kw["cache"] = cache
    try:
        return _run_restarts(gene_expression_data, copy_number_data, jobs, devices, print_elbos, int(restarts_in_flight),
                             bool(batch_y_pass))
    finally:
        if cache is not None:
            for k, v in list(cache.items()):
                if isinstance(k, tuple) and k[0] == "data":
                    v.close()


def _lockstep(gens):
    """Advance several fits through their synchronisation points together; one batched Y pass per round."""
    from .session import ypass_many
    results, live = [None] * len(gens), dict(enumerate(gens))
    try:
        while live:
            waiting = {}
            for i, g in list(live.items()):
                try:
                    waiting[i] = next(g)
                except StopIteration as done:
                    results[i] = done.value
                    del live[i]
            if len(waiting) > 1:
                ypass_many(list(waiting.values()))
    finally:
        # a fit that raised (e.g. "Initial elbo is NA") must not leave the others suspended with their sessions open:
        # closing a generator runs its `finally: sess.close()`, so the shared inputs can be released by the caller
        for g in live.values():
            g.close()
    return results


def _run_restarts(gene_expression_data, copy_number_data, jobs, devices, print_elbos, in_flight=1, batch_y_pass=False):
    if batch_y_pass:
        from concurrent.futures import ThreadPoolExecutor
        devs = list(dict.fromkeys(kw["device"] for kw in jobs))

        def run_on(dev):
            idx = [i for i, kw in enumerate(jobs) if kw["device"] == dev]
            return idx, _lockstep([clonealign_steps(gene_expression_data, copy_number_data, **jobs[i]) for i in idx])

        fits = [None] * len(jobs)
        with ThreadPoolExecutor(max_workers=len(devs)) as pool:
            for idx, res in pool.map(run_on, devs):
                for i, r in zip(idx, res):
                    fits[i] = r
    elif in_flight > 1:
        from concurrent.futures import ThreadPoolExecutor
        devs = list(dict.fromkeys(kw["device"] for kw in jobs))
        pools = {d: ThreadPoolExecutor(max_workers=in_flight) for d in devs}       # k host threads per device
        try:
            futs = [pools[kw["device"]].submit(clonealign, gene_expression_data, copy_number_data, **kw) for kw in jobs]
            fits = [f.result() for f in futs]
        finally:
            for p in pools.values():
                p.shutdown(wait=True)
    elif devices and len(devices) > 1:
        # replicas only: one host thread per GPU, each running its share of the restarts one after another
        # (the C-ABI calls release the GIL); results keep the serial order, so the selection below is unchanged
        from concurrent.futures import ThreadPoolExecutor

        def run_on(dev):
            return [(i, clonealign(gene_expression_data, copy_number_data, **kw))
                    for i, kw in enumerate(jobs) if kw["device"] == dev]

        with ThreadPoolExecutor(max_workers=len(devices)) as pool:
            done = [r for part in pool.map(run_on, list(dict.fromkeys(devices))) for r in part]
        fits = [f for _, f in sorted(done, key=lambda t: t[0])]
    else:
        fits = [clonealign(gene_expression_data, copy_number_data, **kw) for kw in jobs]
    final_elbos = np.array([f["convergence_info"]["final_elbo"] for f in fits])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        median_correlations = np.array([np.nanmedian(f["correlations"]) for f in fits])
    if print_elbos:
        print("ELBOs:  " + " ".join(str(e) for e in final_elbos))
    if np.all(np.isnan(final_elbos)):
        raise ValueError("every restart ended with an NA final ELBO")
    best = fits[int(np.nanargmax(final_elbos))]                 # which.max skips NA (R/clonealign.R:65)
    best["multirun_info"] = {
        "clone_prevalences_at_different_shrinks": [dict(zip(*np.unique(f["clone"], return_counts=True))) for f in fits],
        "elbos": final_elbos,
        "median_correlations": median_correlations,
    }
    return best
