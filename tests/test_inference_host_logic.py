"""CPU tests of the host mirror of `inference_tflow()` / `clonealign()` with the CUDA session replaced by a
recording fake: everything the reference does in R around the session (R/inference-tflow.R:117-235, 368-480;
R/clonealign.R:184-305) is checked without a GPU."""
import numpy as np
import pytest

from clonealign_b200 import api, inference


class FakeSession:
    """Records what the host code hands to the device session and replays a scripted ELBO sequence."""
    last = None
    script = None          # iterable of ELBO values; default: slowly improving

    def __init__(self, Y, L, psi_init, loc_init, **kw):
        FakeSession.last = self
        self.Y, self.L, self.psi_init, self.loc_init, self.kw = np.asarray(Y), np.asarray(L), psi_init, loc_init, kw
        self.calls = []
        self._it = iter(FakeSession.script) if FakeSession.script is not None else None
        self._e = -1000.0

    def init_gamma(self):
        self.calls.append("init_gamma")

    def step(self):
        self.calls.append("step")

    def elbo(self):
        self.calls.append("elbo")
        if self._it is not None:
            return next(self._it)
        self._e += 1.0
        return self._e

    def elbo_many(self, n):
        return np.array([self.elbo() for _ in range(n)])

    def params(self):
        self.calls.append("params")
        N, G = self.Y.shape
        C = self.L.shape[1]
        cp = np.full((N, C), 0.01)
        cp[:, 0] = 1.0 - 0.01 * (C - 1)
        out = {"mu": np.ones(G), "clone_probs": cp, "s": self.Y.sum(1).astype(float), "alpha": np.full(C, 1.0 / C)}
        if self.kw.get("K", 1) > 0:
            out.update(psi=np.zeros((N, 1)), W=np.zeros((G, 1)), chi=np.ones(1))
        if self.kw.get("clone_allele") is not None:
            out["clone_probs_from_snv"] = np.full((N, C), 1.0 / C)
        return out

    def close(self):
        self.calls.append("close")


@pytest.fixture(autouse=True)
def fake_session(monkeypatch):
    FakeSession.script = None
    monkeypatch.setattr(inference, "Session", FakeSession)
    yield


def _data(N=30, G=12, C=3, seed=0):
    rng = np.random.default_rng(seed)
    Y = rng.poisson(4.0, size=(N, G)).astype(float)
    L = rng.integers(1, 9, size=(G, C)).astype(float)
    return Y, L


def test_gene_filter_saturation_and_inits():
    Y, L = _data()
    Y[:, 3] = 0                                   # zero-count gene: removed (:117-124)
    Y[:, 7] = 0
    res = inference.inference_tflow(Y, L, max_iter=3, verbose=False, seed=1)
    s = FakeSession.last
    assert s.Y.shape == (30, 10) and s.L.shape == (10, 3)
    assert res["retained_genes"] == [1, 2, 3, 5, 6, 7, 9, 10, 11, 12]          # which(), 1-based (:130)
    assert s.L.max() <= 6 and np.array_equal(s.L, np.minimum(L[[0, 1, 2, 4, 5, 6, 8, 9, 10, 11]], 6))   # saturate (:142-144)
    keep = [0, 1, 2, 4, 5, 6, 8, 9, 10, 11]
    Yk = Y[:, keep]
    mu_guess = (Yk / Yk.mean(axis=1, keepdims=True)).mean(axis=0)                # :222
    np.testing.assert_allclose(s.loc_init, inference.safe_inverse_softplus(mu_guess))
    assert s.psi_init.shape == (30, 1) and abs(s.psi_init.std(ddof=1) - 1.0) < 0.05
    assert len(res["ml_params"]["mu"]) == len(res["retained_genes"])
    named = inference.inference_tflow(Y, L, max_iter=1, verbose=False, seed=1, gene_names=[f"g{i}" for i in range(12)])
    assert named["retained_genes"] == [f"g{i}" for i in keep]                    # colnames kept when present (:127-128)


def test_loop_semantics_match_reference():
    """:372-417, :447-454 — ELBO_0 after gamma init, (train, eval) per iteration, never stops before 10 iterations,
    20 final evaluations, trace has iters + 1 entries."""
    Y, L = _data()
    FakeSession.script = [-100.0] * 500          # flat ELBO: converged from the start
    res = inference.inference_tflow(Y, L, max_iter=50, rel_tol=1e-6, verbose=False, seed=1)
    calls = FakeSession.last.calls
    assert calls[0] == "init_gamma" and calls[1] == "elbo"
    assert calls.count("step") == 10                                             # window of rep(1e3, 10) (:379)
    assert len(res["convergence_info"]["elbo"]) == 11
    assert calls[2:22] == ["step", "elbo"] * 10
    assert calls[22] == "params" and calls[23:43] == ["elbo"] * 20 and calls[43] == "close"
    assert res["convergence_info"]["final_elbo"] == -100.0 and res["convergence_info"]["sd_final_elbo"] == 0.0
    FakeSession.script = None
    res = inference.inference_tflow(Y, L, max_iter=7, verbose=False, seed=1)
    assert FakeSession.last.calls.count("step") == 7 and len(res["convergence_info"]["elbo"]) == 8


def test_initial_elbo_na_and_session_closed():
    Y, L = _data()
    FakeSession.script = [float("nan")]
    with pytest.raises(ValueError, match="Initial elbo is NA"):
        inference.inference_tflow(Y, L, max_iter=3, verbose=False, seed=1)
    assert FakeSession.last.calls[-1] == "close"                                 # sess$close() even on error


def test_argument_validation_messages():
    Y, L = _data()
    Y0 = Y.copy()
    Y0[4] = 0
    with pytest.raises(ValueError, match="Some cells have no counts mapping"):    # :212-214
        inference.inference_tflow(Y0, L, verbose=False)
    with pytest.raises(ValueError):                                              # stopifnot(nrow(L_dat) == G) :139
        inference.inference_tflow(Y, L[:-1], verbose=False)
    with pytest.raises(ValueError):
        inference.inference_tflow(Y, L, dtype="float16", verbose=False)
    with pytest.raises(ValueError, match="float64"):
        inference.inference_tflow(Y, L, dtype="float64", verbose=False)
    with pytest.raises(ValueError):                                              # stopifnot(nrow(x) == N) :152
        inference.inference_tflow(Y, L, x=np.ones((5, 1)), verbose=False)


def test_covariates_and_k0_quirk():
    Y, L = _data()
    x = np.arange(30.0)
    inference.inference_tflow(Y, L, x=x, max_iter=1, verbose=False, seed=1)
    assert FakeSession.last.kw["x"].shape == (30, 1)                             # vector -> one-column matrix (:149)
    inference.inference_tflow(Y, L, x=x, K=0, max_iter=1, verbose=False, seed=1)
    assert FakeSession.last.kw["x"] is None                                      # :279-285: covariates unused when K == 0
    assert FakeSession.last.psi_init.shape == (30, 0)


def test_allele_inputs_and_ref_equals_cov_quirk():
    Y, L = _data()
    rng = np.random.default_rng(3)
    V = 5
    ca = rng.integers(1, 4, size=(V, 3)).astype(float)
    cov = rng.poisson(2.0, size=(30, V)).astype(float)
    ref = np.minimum(cov, rng.poisson(1.0, size=(30, V))).astype(float)
    inference.inference_tflow(Y, L, clone_allele=ca, cov=cov, ref=ref, max_iter=1, verbose=False, seed=1)
    np.testing.assert_array_equal(FakeSession.last.kw["alt"], cov - ref)         # alt = cov - ref (:180)
    # through clonealign(): the reference forwards ref = cov (R/clonealign.R:271), so alt is identically zero
    fit = api.clonealign(Y, L, clone_allele=ca, cov=cov, ref=ref, max_iter=1, verbose=False, seed=1)
    assert np.all(FakeSession.last.kw["alt"] == 0)
    assert fit["clone_probs_from_snv"].shape == (30, 3)
    api.clonealign(Y, L, clone_allele=ca, cov=cov, ref=ref, max_iter=1, verbose=False, seed=1, fix_ref_bug=True)
    np.testing.assert_array_equal(FakeSession.last.kw["alt"], cov - ref)
    with pytest.raises(ValueError):                                              # sanitize_allele_info
        inference.inference_tflow(Y, L, clone_allele=ca, cov=cov[:, :3], ref=ref, verbose=False)
    inference.inference_tflow(Y, L, clone_allele=ca, cov=cov, ref=None, max_iter=1, verbose=False, seed=1)
    assert FakeSession.last.kw["clone_allele"] is None                           # needs all three (:167)


def test_clonealign_object_and_clone_calls():
    """R/clonealign.R:283-303 and tests/testthat/test_clonealign.R:4-39 on the host side."""
    import warnings
    Y, L = _data(N=40, G=15)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fit = api.clonealign(Y, L, max_iter=2, verbose=False, seed=2)
    assert fit["clone_names"] == ["clone_a", "clone_b", "clone_c"]               # default dimnames (:251-253)
    assert set(fit["clone"]) == {"clone_a"}                                      # fake gamma: 0.98 on the first clone
    assert len(fit["correlations"]) == 15
    assert {"clone", "convergence_info", "retained_genes", "correlations", "ml_params"} <= set(fit)
    assert "A clonealign_fit for 40 cells, 15 genes, and 3 clones" in repr(fit)
    fit2 = api.clonealign(Y, L, max_iter=2, verbose=False, seed=2, clone_call_probability=0.99)
    assert set(fit2["clone"]) == {"unassigned"}                                  # max prob 0.98 < 0.99
    with pytest.raises(ValueError, match="same number of genes"):
        api.clonealign(Y, L[:-2], verbose=False)


def test_seed_controls_all_host_randomness():
    """tests/testthat/test_clonealign.R:42-66: psi noise and the op seed both derive from one RNG."""
    Y, L = _data()
    inference.inference_tflow(Y, L, max_iter=1, verbose=False, seed=12345)
    a = (FakeSession.last.psi_init.copy(), FakeSession.last.kw["seed"])
    inference.inference_tflow(Y, L, max_iter=1, verbose=False, seed=12345)
    b = (FakeSession.last.psi_init.copy(), FakeSession.last.kw["seed"])
    inference.inference_tflow(Y, L, max_iter=1, verbose=False, seed=54321)
    c = (FakeSession.last.psi_init.copy(), FakeSession.last.kw["seed"])
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert not np.array_equal(a[0], c[0]) and a[1] != c[1]


def test_diverged_fit_stops_and_na_restarts_are_not_selected(monkeypatch):
    """The reference's `if(mean(abs(elbo_diffs)) < rel_tol)` (R/inference-tflow.R:414) errors once the ELBO is NA, and
    `which.max` (R/clonealign.R:65) skips NA: a diverged fit must neither run on to max_iter nor be returned as the best."""
    rng = np.random.default_rng(0)
    Y = rng.poisson(2.0, size=(30, 12)).astype(float) + 1.0
    L = rng.integers(1, 4, size=(12, 3)).astype(float)
    FakeSession.script = [-100.0, -90.0, float("nan"), -80.0]
    with pytest.raises(ValueError, match="ELBO is NA after iteration 2"):
        inference.inference_tflow(Y, L, max_iter=50, verbose=False, seed=1)
    assert FakeSession.last.calls[-1] == "close" and FakeSession.last.calls.count("step") == 2
    # best-of-restarts: NA final ELBOs are skipped, all-NA is an error
    fits = iter([{"convergence_info": {"final_elbo": float("nan")}, "correlations": np.zeros(3), "clone": ["A"]},
                 {"convergence_info": {"final_elbo": -5.0}, "correlations": np.zeros(3), "clone": ["B"]},
                 {"convergence_info": {"final_elbo": -7.0}, "correlations": np.zeros(3), "clone": ["C"]}])
    monkeypatch.setattr(api, "clonealign", lambda *a, **k: next(fits))
    best = api.run_clonealign(Y, L, initial_shrinks=(0,), n_repeats=3, print_elbos=False, seed=2)
    assert best["clone"] == ["B"] and np.isnan(best["multirun_info"]["elbos"][0])
    allnan = iter([{"convergence_info": {"final_elbo": float("nan")}, "correlations": np.zeros(3), "clone": ["A"]}] * 2)
    monkeypatch.setattr(api, "clonealign", lambda *a, **k: next(allnan))
    with pytest.raises(ValueError, match="every restart ended with an NA final ELBO"):
        api.run_clonealign(Y, L, initial_shrinks=(0,), n_repeats=2, print_elbos=False, seed=2)


def test_duplicate_gene_names_keep_their_own_columns():
    """The post-hoc correlations use the POSITIONS of the retained genes (the mask of the gene filter), so repeated gene
    names cannot alias each other's columns (they did through a name -> first-index lookup)."""
    rng = np.random.default_rng(3)
    Y = rng.poisson(3.0, size=(40, 6)).astype(float) + 1.0
    Y[:, 2] = 0.0                                   # filtered out
    L = rng.integers(1, 5, size=(6, 3)).astype(float)
    names = ["g", "g", "x", "g", "h", "h"]
    fit = api.clonealign(Y, L, max_iter=3, verbose=False, seed=1, gene_names=names, clone_names=["A", "B", "C"])
    assert fit["retained_genes"] == ["g", "g", "g", "h", "h"] and len(fit["correlations"]) == 5
    keep = [0, 1, 3, 4, 5]
    want = api.compute_correlations(Y[:, keep], L[keep], fit["clone"], ["A", "B", "C"])
    np.testing.assert_array_equal(np.isnan(fit["correlations"]), np.isnan(want))
    np.testing.assert_allclose(np.nan_to_num(fit["correlations"]), np.nan_to_num(want))
