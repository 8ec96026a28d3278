"""CPU parity tests THROUGH THE C-ABI on the emulated build (tests/cuda_emul/): the same checks as
tests/test_gpu_parity.py, for the kernels that need no tensor core (paths "cudacore" and "interp"), run in this
container without a GPU.  What is exercised is the real core.cu (ingest, set-up, launch sequences, Adam, ABI) and the
real kernel sources compiled for the host; what is NOT covered is the tcgen05 path and anything hardware-specific.
The oracle is the checker, exactly as on the GPU.
"""
import warnings

import numpy as np
import pytest

from oracle import clonealign_oracle as O
from test_gpu_parity import ELBO_RTOL, PARAM_RTOL, _case, _check_grads, _load_params, _relmax, _run_trace, _session

pytestmark = pytest.mark.usefixtures("emulated_library")
# (contraction path, kernel variants): the default kernels of both non-tensor paths and the re-engineered variants
# (packed-fp32 Y pass, fused Clenshaw + per-cell epilogue) that bench.py validates on the device before using them
# (ypass2 on the CUDA-core path, the epi2-only and the overlap sets are covered by the storage-format test, test_random_shapes_and_variants, test_variants_agree_with_default_kernels and
# bench.py's candidate tests: on the synchronous emulation `overlap` only changes which stream handle a launch names)
PATHS = [("cudacore", ""), ("interp", ""), ("interp", "ypass2,epi2,lean"),
         ("interp", "ypass3,epi2,lean,defer,overlap"), ("auto", ""),
         ("interp", "ypass4,epi2,lean,defer,cosched"), ("cudacore", "ypass4"), ("interp", "ypass3,epi2,lean,defer,cell2")]


def test_emulated_library_is_not_the_product(emulated_library):
    from clonealign_b200 import _lib
    assert "cuda_emul" in emulated_library and "cuda_emul" in _lib.LIB_PATH
    with _session(np.ones((4, 3)), np.ones((3, 2)), np.zeros((4, 1)), np.ones(3), path="auto") as sess:
        d = sess.describe()                                     # the default model runs the interpolation kernel set
        assert d["path"] == "interp" and d["variants"] == 2 | 4 | 64 | 128 | 256 | 512
    with _session(np.ones((4, 3)), np.ones((3, 2)), np.zeros((4, 2)), np.ones(3), K=2, path="auto") as sess:
        assert sess.describe()["path"] == "cudacore"            # tcgen05 is unavailable under emulation
    from clonealign_b200._lib import CloneAlignLibraryError
    with pytest.raises(CloneAlignLibraryError, match="not available under the CPU emulation|tensor path"):
        _session(np.ones((4, 3)), np.ones((3, 2)), np.zeros((4, 1)), np.ones(3), path="tensor")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("S", [1, 3])
def test_gradients_and_elbo_match_oracle_c1(example_sce, path, S):
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y, L, K=1, seed=S)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=S, K=1, path=path[0], variants=path[1], seed=1) as sess:
        assert sess.describe()["path"] == ("interp" if path[0] == "auto" else path[0])
        _load_params(sess, p)
        errs = _check_grads(sess, d, p, S)
        assert errs["Z"] < 1e-5


@pytest.mark.parametrize("K,P,use_v", [(2, 1, True), (0, 0, False), (1, 2, False), (3, 0, True)])
def test_general_path_covariates_allele(example_sce, K, P, use_v):
    Y, L = example_sce
    d, p, mu_guess, al = _case(Y[:60], L, K=K, P=P, use_v=use_v, seed=K * 7 + P)
    kw = dict(al) if use_v else {}
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=2, K=K, x=d.X, path="cudacore", seed=1, **kw) as sess:
        if use_v:
            assert _relmax(sess.get_array("v"), d.v) < 1e-5
        _load_params(sess, p)
        _check_grads(sess, d, p, 2)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("N,G,C,S", [(130, 70, 5, 3), (257, 193, 2, 1), (64, 640, 7, 8), (33, 2100, 3, 2), (40, 50, 16, 8),
                                     (37, 45, 32, 4), (35, 40, 11, 9)])
def test_ragged_shapes(path, N, G, C, S):
    from clonealign_b200.synthetic import make_synthetic
    syn = make_synthetic(N, G, C, seed=N + G)
    d, p, mu_guess, _ = _case(syn["Y"].astype(np.float64), np.minimum(syn["L"], 6.0), K=1, seed=N)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=S, K=1, path=path[0], variants=path[1], seed=1) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, S)


@pytest.mark.parametrize("path", PATHS)
def test_wide_exponent_range(example_sce, path):
    """psi scaled up: several interpolation panels on each side / large row shifts."""
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y, L, K=1, seed=5, scale=1.0)
    p.psi *= 2.0
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=2, K=1, path=path[0], variants=path[1], seed=1) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, 2)


@pytest.mark.parametrize("path", PATHS)
def test_one_sided_psi(example_sce, path):
    """All psi >= 0 / all psi < 0: one side of the forward panel structure is empty."""
    Y, L = example_sce
    for sign in (+1.0, -1.0):
        d, p, mu_guess, _ = _case(Y[:80], L, K=1, seed=9)
        p.psi = sign * np.abs(p.psi) - sign * 1e-3
        with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=1, K=1, path=path[0], variants=path[1], seed=1) as sess:
            _load_params(sess, p)
            _check_grads(sess, d, p, 1)


@pytest.mark.parametrize("path", PATHS)
def test_allele_fused(example_sce, path):
    Y, L = example_sce
    d, p, mu_guess, al = _case(Y, L, K=1, use_v=True, seed=21)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=1, K=1, path=path[0], variants=path[1], seed=1, **al) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, 1)
        snv = sess.params()["clone_probs_from_snv"]
        ref = np.exp(d.v - np.logaddexp.reduce(d.v, axis=1, keepdims=True))
        assert np.abs(snv - ref).max() < 1e-5


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("S", [1, 3])
def test_loop_matches_golden(example_sce, golden_c1, path, S):
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=None)
    eps = golden_c1[f"eps_S{S}"]
    with _session(hi["Y"], hi["L"], golden_c1["psi_init"], golden_c1["mu_guess"], mc_samples=S, K=1, path=path[0],
                  variants=path[1], learning_rate=0.1, seed=3) as sess:
        sess.set_eps(eps)
        sess.init_gamma()
        elbos = [sess.elbo()]
        for _ in range(5):
            sess.step()
            elbos.append(sess.elbo())
        final = [sess.elbo() for _ in range(3)]
        prm = sess.params()
    ref = golden_c1[f"elbos_S{S}"]
    assert (np.abs(np.array(elbos) - ref) / np.abs(ref)).max() <= ELBO_RTOL, (elbos, ref)
    assert abs(np.mean(final) - golden_c1[f"final_elbo_S{S}"]) <= ELBO_RTOL * abs(golden_c1[f"final_elbo_S{S}"])
    names = ["A", "B", "C"]
    assert O.clone_assignment(prm["clone_probs"], names) == O.clone_assignment(golden_c1[f"clone_probs_S{S}"], names)
    assert np.abs(prm["clone_probs"] - golden_c1[f"clone_probs_S{S}"]).max() <= 2e-3
    assert _relmax(prm["mu"], golden_c1[f"mu_S{S}"]) <= PARAM_RTOL
    assert _relmax(prm["W"], golden_c1[f"W_S{S}"]) <= 5e-3
    assert _relmax(prm["psi"], golden_c1[f"psi_S{S}"]) <= PARAM_RTOL
    assert _relmax(prm["alpha"], golden_c1[f"alpha_S{S}"]) <= PARAM_RTOL
    np.testing.assert_allclose(prm["s"], hi["s"], rtol=0, atol=0)


@pytest.mark.parametrize("path", PATHS)
def test_loop_with_device_rng_matches_oracle(example_sce, path):
    """Philox draws made by the kernels fed back to the oracle: reference loop semantics (fresh draw per run)."""
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(5))
    d = O.Data(hi["Y"], hi["L"])
    S, draws = 2, []
    with _session(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], mc_samples=S, K=1, seed=99, path=path[0], variants=path[1]) as sess:
        def rec():
            draws.append(sess.get_eps().astype(np.float64))
        sess.init_gamma(); rec()
        elbos = [sess.elbo()]; rec()
        for _ in range(4):
            sess.step(); rec()
            elbos.append(sess.elbo()); rec()
        prm = sess.params()
    allz = np.concatenate([x.ravel() for x in draws])
    assert abs(allz.mean()) < 0.08 and abs(allz.std() - 1.0) < 0.08
    assert len({x.tobytes() for x in draws}) == len(draws)
    it = iter(draws)
    p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
    r = O.fit(d, p0, lambda: next(it), max_iter=4, rel_tol=0.0, n_final=0)
    assert (np.abs(np.array(elbos) - r["elbos"]) / np.abs(r["elbos"])).max() <= ELBO_RTOL
    names = ["A", "B", "C"]
    assert O.clone_assignment(prm["clone_probs"], names) == O.clone_assignment(r["clone_probs"], names)


@pytest.mark.parametrize("path", PATHS)
def test_elbo_many_equals_repeated_elbo(example_sce, path):
    """ca_core_elbo_many (the 20 fresh-draw evaluations behind final_elbo / sd_final_elbo, R/inference-tflow.R:447-449,
    fetched with one device-to-host copy) returns bit-for-bit what the same number of ca_core_elbo calls returns, consumes
    the same draws, and leaves the session in the same state."""
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(2))

    def run(many):
        with _session(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], mc_samples=2, K=1, seed=31, path=path[0], variants=path[1]) as sess:
            sess.init_gamma()
            sess.step()
            e = sess.elbo_many(5) if many else np.array([sess.elbo() for _ in range(5)])
            sess.step()
            return e, sess.elbo(), sess.elbo_many(0)
    a, a_next, empty = run(True)
    b, b_next, _ = run(False)
    assert a.tobytes() == b.tobytes() and a_next == b_next and empty.size == 0
    assert len(set(a.tolist())) == 5                                   # fresh draws: five different values


@pytest.mark.parametrize("path", PATHS)
def test_same_seed_bitwise_identical(example_sce, path):
    """tests/testthat/test_clonealign.R:42-66 (fixed-order reductions: also independent of the host thread count)."""
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    a, ga = _run_trace(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], n=3, mc_samples=2, seed=12345, path=path[0], variants=path[1])
    b, gb = _run_trace(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], n=3, mc_samples=2, seed=12345, path=path[0], variants=path[1])
    c, _ = _run_trace(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], n=3, mc_samples=2, seed=54321, path=path[0], variants=path[1])
    assert a.tobytes() == b.tobytes() and ga.tobytes() == gb.tobytes()
    assert a.tobytes() != c.tobytes()


def test_storage_formats_and_input_layouts_agree(example_sce):
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    run = lambda y, **kw: _run_trace(y, hi["L"], hi["psi_init"], hi["mu_guess"], n=2, seed=7, **kw)[0]
    # the contraction-kernel paths widen the counts and use the 8-column tiling for every storage type: bit-identical
    traces = [run(hi["Y"], y_store=s, path="cudacore") for s in ("f32", "u16", "u8")]
    assert traces[0].tobytes() == traces[1].tobytes() == traces[2].tobytes()
    packed = [run(hi["Y"], y_store=s, path="cudacore", variants="ypass2") for s in ("f32", "u16", "u8")]   # f32x2 Y pass: same exactness
    assert packed[0].tobytes() == packed[1].tobytes() == packed[2].tobytes()
    assert np.abs(packed[0] - traces[0]).max() <= 1e-6 * np.abs(traces[0]).max()             # re-associated sums only
    # ypass3: the stored integer is used as a denormal fp32 operand, the other operand carries the scale -> bit-identical to
    # the arithmetic on widened counts (f32 storage runs the same tiling unscaled); u8 owns 16 columns per thread instead
    # of 8, which only re-associates the row sums
    y3 = [run(hi["Y"], y_store=s, path="cudacore", variants="ypass3") for s in ("f32", "u16", "u8")]
    assert y3[0].tobytes() == y3[1].tobytes()
    assert np.abs(y3[2] - y3[0]).max() <= 1e-6 * np.abs(y3[0]).max() and np.abs(y3[0] - traces[0]).max() <= 1e-6 * np.abs(traces[0]).max()
    # the default kernel set (path = auto: interp + ypass3): same statement
    d3 = [run(hi["Y"], y_store=s) for s in ("f32", "u16", "u8")]
    assert np.all(np.isfinite(d3[0])) and d3[0].tobytes() == d3[1].tobytes() and np.abs(d3[2] - d3[0]).max() <= 1e-6 * np.abs(d3[0]).max()
    t_f = run(np.asfortranarray(hi["Y"]), y_store="f32")                    # an R double matrix
    t_i = run(np.asfortranarray(hi["Y"].astype(np.int32)))                  # an R integer matrix (stored u8)
    t_32 = run(hi["Y"].astype(np.float32))
    assert d3[0].tobytes() == t_f.tobytes() and d3[2].tobytes() == t_i.tobytes() == t_32.tobytes()
    import scipy.sparse as sp
    for compact in (hi["Y"].astype(np.uint8), np.asfortranarray(hi["Y"].astype(np.uint16)), sp.csr_matrix(hi["Y"].astype(np.uint8))):
        assert run(compact).tobytes() == d3[2].tobytes()                      # compact host counts (CA_Y_U8 / CA_Y_U16)
    from clonealign_b200._lib import CloneAlignLibraryError
    with pytest.raises(CloneAlignLibraryError, match="u8"):
        run(hi["Y"] + 0.5, y_store="u8")
    big = hi["Y"].copy()
    big[0, 0] = 300.0
    with _session(big, hi["L"], hi["psi_init"], hi["mu_guess"]) as sess:
        assert sess.describe()["y_store"] == "u16"
    big[0, 0] = 70000.0
    with _session(big, hi["L"], hi["psi_init"], hi["mu_guess"]) as sess:
        assert sess.describe()["y_store"] == "f32"


def test_fused_epilogue_coefficients_through_l2(example_sce, monkeypatch):
    """More active panels than the shared-memory table holds: the fused kernel reads coefficients through L2."""
    monkeypatch.setenv("CLONEALIGN_B200_FUSED_PANELS", "1")
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y, L, K=1, seed=5, scale=1.0)
    p.psi *= 2.0
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=3, K=1, path="interp", variants="epi2", seed=1) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, 3)


@pytest.mark.parametrize("panels", ["1", "2"])
def test_cell2_panels_staged_in_rounds(example_sce, monkeypatch, panels):
    """CELL2 set with more active panels than the shared-memory tables of the per-cell / per-gene kernels hold: the blocks
    walk their cells (genes) once per subset of panels, every cell is worked on in the round that holds its panel."""
    monkeypatch.setenv("CLONEALIGN_B200_FUSED_PANELS", panels)
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y, L, K=1, seed=5, scale=1.0)
    p.psi *= 2.0
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=3, K=1, path="auto", seed=1) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, 3)
        pan = sess.describe()["panels"]
        assert sess.describe()["variants"] & 512 and pan["nf_neg"] + pan["nf_pos"] > 2 and pan["nb"] > 2


def test_variant_validation(example_sce):
    from clonealign_b200._lib import CloneAlignLibraryError
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y[:40], L, K=1, seed=1)
    with pytest.raises(CloneAlignLibraryError, match="epi2 belongs to the interp path"):
        _session(d.Y, d.L, p.psi, mu_guess, path="cudacore", variants="epi2")
    from clonealign_b200.synthetic import make_synthetic
    syn = make_synthetic(20, 30, 33, seed=3)
    with pytest.raises(CloneAlignLibraryError, match="epi2 needs C <= 32"):
        _session(syn["Y"].astype(np.float64) + 1.0, np.minimum(syn["L"], 6.0), np.zeros((20, 1)), np.ones(30), path="interp",
                 variants="epi2")
    with pytest.raises(ValueError, match="unknown kernel variant"):
        _session(d.Y, d.L, p.psi, mu_guess, variants="nope")
    with _session(d.Y, d.L, p.psi, mu_guess, path="interp", variants=["ypass2", "epi2"]) as sess:
        assert sess.describe()["variants"] == 3
    with pytest.raises(CloneAlignLibraryError, match="lean needs variant epi2"):
        _session(d.Y, d.L, p.psi, mu_guess, path="interp", variants="lean")
    with pytest.raises(CloneAlignLibraryError, match="defer needs variants epi2 and lean"):
        _session(d.Y, d.L, p.psi, mu_guess, path="interp", variants="epi2,defer")
    with pytest.raises(CloneAlignLibraryError, match="ypass2 and ypass3 are alternatives"):
        _session(d.Y, d.L, p.psi, mu_guess, variants="ypass2,ypass3")


def test_variants_agree_with_default_kernels(example_sce):
    """The re-engineered kernels against the default ones on the same inputs (what bench.py's on-device self-check
    does at full size): ELBO trace and fitted parameters after 3 steps."""
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    out = {}
    for name, (path, var) in {"ref": ("cudacore", ""), "new": ("interp", "ypass2,epi2,lean")}.items():
        with _session(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], mc_samples=2, seed=5, path=path, variants=var) as sess:
            sess.init_gamma()
            tr = [sess.elbo()]
            for _ in range(3):
                sess.step()
                tr.append(sess.elbo())
            out[name] = (np.array(tr), sess.params())
    assert (np.abs(out["ref"][0] - out["new"][0]) / np.abs(out["ref"][0])).max() < 1e-6
    for k in ("psi", "W", "mu", "clone_probs", "alpha"):
        assert _relmax(out["new"][1][k], out["ref"][1][k]) < 1e-4, k


# ---------------------------------------------------------------------------------------------------
# the reference-facing API over the emulated ABI (mirrors tests/testthat/test_clonealign.R)
# ---------------------------------------------------------------------------------------------------
def test_clonealign_returns_valid_object(example_sce):
    from clonealign_b200 import clonealign
    Y, L = example_sce
    N, G = Y.shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = clonealign(Y, L, max_iter=5, clone_names=["A", "B", "C"], verbose=False, seed=1)
    assert len(cal["clone"]) == N and set(cal["clone"]) <= {"A", "B", "C", "unassigned"}
    assert cal["ml_params"]["clone_probs"].shape == (N, 3)
    assert len(cal["retained_genes"]) == len(cal["ml_params"]["mu"]) <= G
    assert {"clone", "convergence_info", "retained_genes", "correlations", "ml_params"} <= set(cal)
    assert len(cal["convergence_info"]["elbo"]) == 6
    np.testing.assert_allclose(cal["ml_params"]["clone_probs"].sum(1), 1.0, atol=1e-6)
    # the same call with the cells of the fit sharded over "two GPUs" of this process (options(clonealign.gpus) in R)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        two = clonealign(Y, L, max_iter=5, clone_names=["A", "B", "C"], verbose=False, seed=1, devices=[0, 0])
    assert two["clone"] == cal["clone"]
    assert np.abs(two["convergence_info"]["elbo"] / cal["convergence_info"]["elbo"] - 1.0).max() < 1e-6
    assert np.abs(two["ml_params"]["clone_probs"] - cal["ml_params"]["clone_probs"]).max() < 1e-4


def test_seed_setting_works_and_na_paths(example_sce):
    from clonealign_b200 import clonealign, inference_tflow
    Y, L = example_sce
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = clonealign(Y, L, max_iter=3, verbose=False, seed=12345)
        b = clonealign(Y, L, max_iter=3, verbose=False, seed=12345)
    assert a["convergence_info"]["final_elbo"] == b["convergence_info"]["final_elbo"]
    L0 = L.copy()
    L0[3, 1] = 0.0
    with pytest.raises(ValueError, match="Initial elbo is NA"):
        inference_tflow(Y, L0, max_iter=2, verbose=False, seed=1)


def test_device_correlations_match_host_mirror(example_sce):
    """ca_core_correlations (post-hoc compute_correlations, R/clonealign.R:318-334) against the host mirror, with
    unassigned cells, a constant gene (NA), an unsaturated L, and through clonealign() itself."""
    from clonealign_b200 import clonealign, compute_correlations
    Y, L = example_sce
    keep = Y.sum(0) > 0
    Y, L = Y[:, keep].copy(), L[keep].copy() * 1.7
    Y[:, 3] = 2.0                                           # constant expression -> NA
    rng = np.random.default_rng(4)
    zidx = rng.integers(-1, L.shape[1], size=Y.shape[0]).astype(np.int32)
    names = ["A", "B", "C"]
    clones = ["unassigned" if z < 0 else names[z] for z in zidx]
    want = compute_correlations(Y, L, clones, names)
    for store in ("u8", "f32"):
        with _session(Y, np.minimum(L, 6.0), np.zeros((Y.shape[0], 1)), np.ones(Y.shape[1]), y_store=store) as sess:
            got = sess.correlations(zidx, L)
            sat = sess.correlations(zidx)                   # NULL -> the session's saturated copy number
        assert np.isnan(got[3]) and np.isnan(want[3])
        ok = ~np.isnan(want)
        assert (np.isnan(got) == np.isnan(want)).all() and np.abs(got[ok] - want[ok]).max() < 1e-9
        want_sat = compute_correlations(Y, np.minimum(L, 6.0), clones, names)
        oks = ~np.isnan(want_sat)
        assert np.abs(sat[oks] - want_sat[oks]).max() < 1e-6          # the session keeps L in fp32
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = clonealign(example_sce[0], example_sce[1], max_iter=3, verbose=False, seed=5, clone_names=names,
                       device_correlations=True)
        b = clonealign(example_sce[0], example_sce[1], max_iter=3, verbose=False, seed=5, clone_names=names)
    assert a["clone"] == b["clone"]
    np.testing.assert_allclose(a["correlations"], b["correlations"], atol=1e-9, equal_nan=True)


def test_device_pca_matches_host_svd(example_sce):
    """ca_core_pca_scores (power iteration on the resident Y) against prcomp's definition via a full SVD
    (R/inference-tflow.R:203-205), for every storage format; constant columns are rejected as R's scale() does."""
    from clonealign_b200._lib import CloneAlignLibraryError
    from clonealign_b200.inference import inference_tflow, pca_init
    Y, L = example_sce
    keep = Y.sum(0) > 0
    Y, L = Y[:, keep], L[keep]

    class NoNoise:
        def normal(self, *a, size=None, **k):
            return np.zeros(size)
    want = pca_init(Y, 1, NoNoise(), truncated=False)[:, 0]                 # scale(pcs), no noise
    for store in ("u8", "f32"):
        with _session(Y, L, np.zeros((Y.shape[0], 1)), np.ones(Y.shape[1]), y_store=store) as sess:
            got, iters = sess.pca_scores()
        got = (got - got.mean()) / got.std(ddof=1)
        err = min(np.abs(got - want).max(), np.abs(got + want).max())       # sign of a principal component is arbitrary
        assert err < 1e-5 and 1 < iters < 500, (err, iters)
    Yc = Y.copy()
    Yc[:, 2] = 3.0
    with _session(Yc, L, np.zeros((Y.shape[0], 1)), np.ones(Y.shape[1])) as sess:
        with pytest.raises(CloneAlignLibraryError, match="constant/zero column"):
            sess.pca_scores()
    # through the host mirror: same RNG stream positions as the host-PCA path -> same noise, same op seed
    a = inference_tflow(Y, L, max_iter=3, verbose=False, seed=11, device_pca=True)
    b = inference_tflow(Y, L, max_iter=3, verbose=False, seed=11)
    ea, eb = a["convergence_info"]["elbo"], b["convergence_info"]["elbo"]
    # W starts at 0, so the first ELBO does not see the (arbitrary) sign of the component; later iterates may differ
    # because the additive noise does not flip with it
    assert abs(ea[0] - eb[0]) / abs(eb[0]) < 1e-4 and np.all(np.isfinite(ea)) and ea[-1] > ea[0]


def test_sparse_input_stays_compressed(example_sce):
    """CA_Y_CSR ingest: a scipy.sparse cells x genes matrix -- the transposed dgCMatrix of a
    SingleCellExperiment -- gives bit-identical fits to the dense matrix, for every value type and through the host
    mirror (with the device PCA nothing is densified on the host); malformed index arrays fail loudly."""
    import scipy.sparse as sp
    from clonealign_b200._lib import CloneAlignLibraryError
    from clonealign_b200.inference import inference_tflow
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    run = lambda y: _run_trace(y, hi["L"], hi["psi_init"], hi["mu_guess"], n=2, seed=7)[0]
    dense = run(hi["Y"])
    for conv in (sp.csr_matrix, sp.csc_matrix, sp.coo_matrix, lambda y: sp.csr_matrix(y.astype(np.float32)),
                 lambda y: sp.csr_matrix(y.astype(np.int32))):
        assert run(conv(hi["Y"])).tobytes() == dense.tobytes()
    bad = sp.csr_matrix(hi["Y"])
    bad.indices = bad.indices.copy()
    bad.indices[5] = hi["Y"].shape[1] + 3
    bad.has_canonical_format = True                       # keep scipy from touching the broken index
    with pytest.raises(CloneAlignLibraryError, match="gene index outside"):
        from clonealign_b200.session import Session
        Session(bad, hi["L"], hi["psi_init"], O.safe_inverse_softplus(hi["mu_guess"]))
    a = inference_tflow(sp.csr_matrix(Y), L, max_iter=3, verbose=False, seed=11, device_pca=True)
    b = inference_tflow(Y, L, max_iter=3, verbose=False, seed=11, device_pca=True)
    assert a["convergence_info"]["elbo"].tobytes() == b["convergence_info"]["elbo"].tobytes()
    assert a["retained_genes"] == b["retained_genes"]
    c = inference_tflow(sp.csc_matrix(Y), L, max_iter=2, verbose=False, seed=11)        # host PCA: densified on the host
    d = inference_tflow(Y, L, max_iter=2, verbose=False, seed=11)
    assert c["convergence_info"]["elbo"].tobytes() == d["convergence_info"]["elbo"].tobytes()


@pytest.mark.parametrize("path", [("cudacore", ""), ("interp", "ypass2,epi2,lean"), ("cudacore", "p2p"),
                                  ("interp", "ypass2,epi2,lean,p2p"), ("interp", "ypass3,epi2,lean"),
                                  ("interp", "ypass3,epi2,lean,defer,overlap"), ("auto", ""), ("auto", "p2p")])
@pytest.mark.parametrize("world", [2, 3])
def test_cell_sharded_fit_matches_single_shard(example_sce, path, world):
    """SURVEY 8e through the REAL sharded code path of core.cu: `world` ranks (threads of this process; the emulation
    build swaps NCCL for an in-process rendezvous, tests/cuda_emul/nccl_emul.h) each hold a block of cells, sum the
    gene-level gradient partials with one all-reduce per step, and must reproduce the single-shard fit: ELBO trace within
    fp32 re-association, identical hard clone calls, per-cell parameters of the shards = rows of the global ones."""
    import threading
    from clonealign_b200 import dist as D
    from clonealign_b200.session import Session
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    Yk, Lk, psi, mu_guess = hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"]
    N = Yk.shape[0]
    loc = O.safe_inverse_softplus(mu_guess)
    kw = dict(mc_samples=2, K=1, seed=77, path=path[0], variants=path[1])

    def trace(sess, out):
        sess.init_gamma()
        tr = [sess.elbo()]
        for _ in range(3):
            sess.step()
            tr.append(sess.elbo())
        tr += list(sess.elbo_many(2))                   # collective like elbo(): one all-reduce per evaluation
        out["elbo"], out["prm"] = np.array(tr), sess.params()
        sess.close()

    ref = {}
    trace(Session(Yk, Lk, psi, loc, **dict(kw, variants=path[1].replace(",p2p", "").replace("p2p", ""))), ref)
    p2p = "p2p" in path[1]
    handles, gate = [None] * world, threading.Barrier(world)
    nid = Session.nccl_unique_id()
    colsum = Yk.sum(axis=0)
    outs, errs = [dict() for _ in range(world)], []

    def rank_main(r):
        try:
            a, b = D.shard_bounds(N, r, world)
            s = Session(Yk[a:b], Lk, psi[a:b], loc, rank=r, world=world, nccl_id=nid, n_total=N, colsum_total=colsum, **kw)
            if p2p:      # variant p2p: all-reduce kernel over "peer memory" (IPC handles are plain pointers here)
                handles[r] = s.p2p_export()
                gate.wait(timeout=120)
                s.p2p_connect(handles)
            trace(s, outs[r])
        except Exception as e:          # a dead rank would leave the others waiting in the rendezvous
            errs.append(e)
            raise
    ts = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errs and all("elbo" in o for o in outs), errs
    for o in outs[1:]:
        assert o["elbo"].tobytes() == outs[0]["elbo"].tobytes()                 # every rank reports the same ELBO
        assert o["prm"]["mu"].tobytes() == outs[0]["prm"]["mu"].tobytes()       # replicated gene-level state stays in step
    assert (np.abs(outs[0]["elbo"] - ref["elbo"]) / np.abs(ref["elbo"])).max() < 1e-6
    cp = np.concatenate([o["prm"]["clone_probs"] for o in outs])
    names = ["A", "B", "C"]
    assert O.clone_assignment(cp, names) == O.clone_assignment(ref["prm"]["clone_probs"], names)
    assert np.abs(cp - ref["prm"]["clone_probs"]).max() < 1e-4
    assert _relmax(np.concatenate([o["prm"]["psi"] for o in outs]), ref["prm"]["psi"]) < 1e-4
    assert _relmax(outs[0]["prm"]["W"], ref["prm"]["W"]) < 1e-3 and _relmax(outs[0]["prm"]["alpha"], ref["prm"]["alpha"]) < 1e-4


@pytest.mark.parametrize("layout", ["rowmajor_u8", "colmajor_f64", "csr"])
@pytest.mark.parametrize("world", [1, 3])
def test_single_process_multi_gpu_equals_one_rank_per_process(example_sce, world, layout):
    """ca_core_multi_* (one host thread drives all shards: what the R boundary needs) against the path the
    torchrun launch takes (one caller per shard, here Python threads): the same per-shard calls in the same order, so ELBO
    traces and parameters agree BIT FOR BIT; every input layout R can hand over is split by rows inside the library."""
    import threading
    import scipy.sparse as sp
    from clonealign_b200 import dist as D
    from clonealign_b200.session import MultiSession, Session
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    Yk, Lk, psi, mu_guess = hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"]
    N = Yk.shape[0]
    loc = O.safe_inverse_softplus(mu_guess)
    rng = np.random.default_rng(3)
    V = 7
    cn = rng.integers(1, 4, size=(V, Lk.shape[1])).astype(float)
    cov = rng.poisson(0.7, size=(N, V)).astype(float)
    alt = rng.binomial(cov.astype(int), 0.4).astype(float)
    kw = dict(mc_samples=2, K=1, seed=77, clone_allele=cn)
    Yin = {"rowmajor_u8": Yk.astype(np.uint8), "colmajor_f64": np.asfortranarray(Yk), "csr": sp.csr_matrix(Yk)}[layout]
    if world > 1 and layout == "csr":
        # variant p2p inside one process: the exchange buffers are wired up by ca_core_multi_create itself; the rank-ordered
        # reduction of the kernel gives the same bits as the (emulated, rank-ordered) NCCL all-reduce
        with MultiSession(Yin, Lk, psi, loc, devices=[0] * world, alt=alt, cov=cov, variants="p2p", **kw) as ms:
            e_p2p, p_p2p = None, None
            ms.init_gamma()
            e_p2p = [ms.elbo()]
            for _ in range(3):
                ms.step()
                e_p2p.append(ms.elbo())
            p_p2p = ms.params()

    def trace(sess):
        sess.init_gamma()
        tr = [sess.elbo()]
        for _ in range(3):
            sess.step()
            tr.append(sess.elbo())
        tr += list(sess.elbo_many(2))
        return np.array(tr), sess.params()

    with MultiSession(Yin, Lk, psi, loc, devices=[0] * world, alt=alt, cov=cov, **kw) as ms:
        d = ms.describe()
        assert d["path"] == "interp" and d["world"] == world and d["shards"] == [D.shard_bounds(N, r, world) for r in range(world)]
        e_multi, p_multi = trace(ms)
        assert ms.time_steps(1) >= 0.0
    outs, errs = [None] * world, []
    nid = Session.nccl_unique_id()

    def rank_main(r):
        try:
            a, b = D.shard_bounds(N, r, world)
            s = Session(Yk[a:b], Lk, psi[a:b], loc, rank=r, world=world, nccl_id=nid if world > 1 else None, n_total=N,
                        alt=alt[a:b], cov=cov[a:b], **kw)      # colsum_total = None: summed by the library (collective)
            outs[r] = trace(s)
            s.close()
        except Exception as e:
            errs.append(e)
            raise
    ts = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errs and all(o is not None for o in outs), errs
    assert e_multi.tobytes() == outs[0][0].tobytes()
    if world > 1 and layout == "csr":
        assert np.array(e_p2p).tobytes() == e_multi[:4].tobytes()
    for k in ("clone_probs", "psi", "s", "clone_probs_from_snv"):
        assert p_multi[k].tobytes() == np.concatenate([o[1][k] for o in outs]).tobytes(), k
    for k in ("mu", "W", "alpha", "chi"):
        assert p_multi[k].tobytes() == outs[0][1][k].tobytes(), k


@pytest.mark.parametrize("path", [("cudacore", "ypass2"), ("interp", "ypass2,epi2,lean"), ("interp", "ypass3,epi2,lean"),
                                  ("interp", "ypass4,epi2,lean,defer,cosched")])
def test_several_row_and_column_tiles(path):
    """N > 2 row blocks of the Y pass (RB = 512), G > one 2048-column tile, more cells than one sweep of the persistent
    per-cell / per-gene kernels: tile seams, partial-sum layouts and strided loops."""
    from clonealign_b200.synthetic import make_synthetic
    syn = make_synthetic(1100, 4500 if ("ypass3" in path[1] or "ypass4" in path[1]) else 2300, 4, seed=8)      # ypass3 / u8: 4096-column tiles
    d, p, mu_guess, _ = _case(syn["Y"].astype(np.float64), np.minimum(syn["L"], 6.0), K=1, seed=2)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=2, K=1, path=path[0], variants=path[1], seed=1) as sess:
        _load_params(sess, p)
        _check_grads(sess, d, p, 2)


@pytest.mark.parametrize("path", [("interp", ""), ("interp", "ypass2,epi2,lean")])
def test_c3_column_structure(path):
    """C = 12 clones, S = 8 samples (J = 192 columns): the template instantiations BASELINE config 3 runs on the device
    (6 columns per lane in the node kernels, 3 Clenshaw chains per lane in the fused per-cell kernel, 16-byte operand
    stores), on a matrix small enough for the emulation."""
    from clonealign_b200.synthetic import make_synthetic
    syn = make_synthetic(700, 2600, 12, seed=12)
    d, p, mu_guess, _ = _case(syn["Y"].astype(np.float64), np.minimum(syn["L"], 6.0), K=1, seed=4, scale=0.3)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=8, K=1, path=path[0], variants=path[1], seed=1) as sess:
        assert sess.describe()["J"] == 192
        _load_params(sess, p)
        errs = _check_grads(sess, d, p, 8)
        assert errs["Z"] < 5e-6


@pytest.mark.parametrize("path", [("interp", ""), ("interp", "ypass2,epi2,lean"), ("auto", "")])
def test_nan_parameters_give_nan_not_a_fault(example_sce, path):
    """A diverged fit (NaN in psi) must surface as a NaN ELBO, as in the reference (R/inference-tflow.R:411-412 lets NA
    ELBOs through after the first iteration) -- not as an out-of-range panel index in the interpolation tables."""
    Y, L = example_sce
    d, p, mu_guess, _ = _case(Y[:64], L, K=1, seed=3)
    with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=1, K=1, path=path[0], variants=path[1], seed=1) as sess:
        _load_params(sess, p)
        bad = p.psi.copy()
        bad[5, 0] = np.nan
        sess.set_array("psi", bad)
        assert np.isnan(sess.elbo())
        sess.step()
        assert np.isnan(sess.elbo())
        bad[:] = np.nan
        sess.set_array("psi", bad)
        assert np.isnan(sess.elbo())


def test_shared_device_inputs_for_restarts(example_sce):
    """ca_core_data_create / ca_core_create_shared.  Sessions built on shared, read-only device inputs are
    bit-identical to self-contained ones (also with the allele term and a non-default kernel set); run_clonealign with
    share_inputs uploads / preprocesses / decomposes once and returns the same best fit; the inputs cannot be destroyed
    while a session uses them."""
    from clonealign_b200 import run_clonealign
    from clonealign_b200._lib import CloneAlignLibraryError
    from clonealign_b200.session import DeviceData, Session
    Y, L = example_sce
    d, p, mu_guess, al = _case(Y, L, K=1, use_v=True, seed=21)
    loc = O.safe_inverse_softplus(mu_guess)

    def trace(**kw):
        with Session(d.Y, d.L, p.psi, loc, mc_samples=2, K=1, seed=5, **kw) as s:
            s.init_gamma()
            out = [s.elbo()]
            for _ in range(2):
                s.step()
                out.append(s.elbo())
            return np.array(out), s.params()["clone_probs"], s.params().get("clone_probs_from_snv")
    with DeviceData(d.Y, d.L, **al) as data:
        for kw in (dict(path="cudacore"), dict(path="interp", variants="ypass2,epi2,lean")):
            a = trace(**kw, **al)
            b = trace(data=data, **kw)
            c = trace(data=data, **kw)                                   # a second session on the same inputs
            assert a[0].tobytes() == b[0].tobytes() == c[0].tobytes() and a[1].tobytes() == b[1].tobytes()
            assert a[2].tobytes() == b[2].tobytes()
        s = Session(None, None, p.psi, loc, data=data)
        with pytest.raises(CloneAlignLibraryError, match="still use these inputs"):
            data.close()
        s.close()
        with pytest.raises(ValueError):                                 # wrong number of cells for these inputs
            Session(None, None, p.psi[:10], loc, data=data)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(initial_shrinks=(0,), n_repeats=3, print_elbos=False, max_iter=3, verbose=False, seed=3)
        f1 = run_clonealign(Y, L, share_inputs=True, **kw)
        f2 = run_clonealign(Y, L, share_inputs=False, **kw)
    assert f1["multirun_info"]["elbos"].tobytes() == f2["multirun_info"]["elbos"].tobytes()
    assert f1["clone"] == f2["clone"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f3 = run_clonealign(Y, L, share_inputs=True, restarts_in_flight=3, **kw)       # concurrent restarts on one device
    assert f3["multirun_info"]["elbos"].tobytes() == f2["multirun_info"]["elbos"].tobytes() and f3["clone"] == f2["clone"]
    # lock-step restarts with one batched Y pass per iteration: same fits when every restart uses the packed Y pass
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kv = dict(kw, path="interp", variants="ypass2,epi2,lean", max_iter=11, rel_tol=1e-3)      # fits stop at different iterations
        f4 = run_clonealign(Y, L, batch_y_pass=True, **kv)
        f5 = run_clonealign(Y, L, **kv)
    assert f4["multirun_info"]["elbos"].tobytes() == f5["multirun_info"]["elbos"].tobytes() and f4["clone"] == f5["clone"]


def test_clonealign_accepts_sparse_counts(example_sce):
    """clonealign() on the sparse counts of a SingleCellExperiment (scipy.sparse here): same fit and correlations as on
    the dense matrix, with and without the device-side PCA / correlations."""
    import scipy.sparse as sp
    from clonealign_b200 import clonealign
    Y, L = example_sce
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(max_iter=2, verbose=False, seed=1)
        a = clonealign(sp.csr_matrix(Y), L, device_pca=True, device_correlations=True, **kw)
        b = clonealign(Y, L, device_pca=True, **kw)
        c = clonealign(sp.csc_matrix(Y), L, **kw)
        d = clonealign(Y, L, **kw)
    assert a["convergence_info"]["elbo"].tobytes() == b["convergence_info"]["elbo"].tobytes() and a["clone"] == b["clone"]
    np.testing.assert_allclose(a["correlations"], b["correlations"], atol=1e-9, equal_nan=True)
    assert c["convergence_info"]["elbo"].tobytes() == d["convergence_info"]["elbo"].tobytes()
    np.testing.assert_allclose(c["correlations"], d["correlations"], atol=0, equal_nan=True)


@pytest.mark.parametrize("n_fits,store", [(2, "u8"), (3, "u16"), (5, "f32")])
def test_batched_y_pass_for_restarts(example_sce, n_fits, store):
    """ca_core_ypass_many: one stream over the shared count matrix yields every fit's (Y W, Y^T psi) partials; fits
    stepped in lock-step with it are bit-identical to fits that each run their own (packed) Y pass."""
    from clonealign_b200.session import DeviceData, Session, ypass_many
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    loc = O.safe_inverse_softplus(hi["mu_guess"])
    rng = np.random.default_rng(3)
    psis = [hi["psi_init"] + rng.normal(0, 0.05, size=hi["psi_init"].shape) for _ in range(n_fits)]
    kw = dict(mc_samples=2, K=1, path="interp", variants="ypass2,epi2,lean")

    def solo(i):
        with Session(hi["Y"], hi["L"], psis[i], loc, seed=10 + i, y_store=store, **kw) as s:
            s.init_gamma()
            tr = [s.elbo()]
            for _ in range(3):
                s.step()
                tr.append(s.elbo())
            return np.array(tr), s.params()["psi"]
    want = [solo(i) for i in range(n_fits)]
    with DeviceData(hi["Y"], hi["L"], y_store=store) as data:
        ss = [Session(None, None, psis[i], loc, seed=10 + i, data=data, **kw) for i in range(n_fits)]
        assert ss[0].describe()["y_store"] == store
        for s in ss:
            s.init_gamma()
        ypass_many(ss)
        trs = [[s.elbo()] for s in ss]
        for _ in range(3):
            for s in ss:
                s.step()                     # uses the partial sums of the batched pass, then changes the parameters
            ypass_many(ss)                   # one read of Y for all fits
            for s, tr in zip(ss, trs):
                tr.append(s.elbo())
        got = [(np.array(tr), s.params()["psi"]) for s, tr in zip(ss, trs)]
        assert all(s.describe()["launches_last_step"] > 0 for s in ss)
        for s in ss:
            s.close()
    for (a, pa), (b, pb) in zip(want, got):
        assert a.tobytes() == b.tobytes() and pa.tobytes() == pb.tobytes()


def test_device_pca_and_correlations_under_cell_sharding(example_sce):
    """SURVEY 8f-1 / 8f-3 combined with 8e: with the cells sharded over ranks, ca_core_pca_scores and ca_core_correlations
    are collective calls (column statistics, X^T t and the correlation sums are all-reduced) and reproduce the
    single-shard results."""
    import threading
    from clonealign_b200 import compute_correlations, dist as D
    from clonealign_b200.session import Session
    Y, L = example_sce
    keep = Y.sum(0) > 0
    Y, L = Y[:, keep], L[keep]
    N, G = Y.shape
    loc = np.ones(G)
    zidx = np.random.default_rng(4).integers(-1, L.shape[1], size=N).astype(np.int32)
    with Session(Y, L, np.zeros((N, 1)), loc) as s:
        want_pcs, _ = s.pca_scores()
        want_cor = s.correlations(zidx, L * 1.3)
    world, nid, colsum = 3, Session.nccl_unique_id(), Y.sum(axis=0)
    outs, errs = [None] * world, []

    def rank_main(r):
        try:
            a, b = D.shard_bounds(N, r, world)
            with Session(Y[a:b], L, np.zeros((b - a, 1)), loc, rank=r, world=world, nccl_id=nid, n_total=N, colsum_total=colsum) as s:
                outs[r] = (s.pca_scores()[0], s.correlations(zidx[a:b], L * 1.3))
        except Exception as e:
            errs.append(e)
            raise
    ts = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errs and all(o is not None for o in outs), errs
    pcs = np.concatenate([o[0] for o in outs])
    assert min(np.abs(pcs - want_pcs).max(), np.abs(pcs + want_pcs).max()) < 1e-8 * np.abs(want_pcs).max()
    for o in outs:
        np.testing.assert_allclose(o[1], want_cor, atol=1e-12, equal_nan=True)
    names = ["A", "B", "C"]
    host = compute_correlations(Y, L * 1.3, ["unassigned" if z < 0 else names[z] for z in zidx], names)
    np.testing.assert_allclose(want_cor, host, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("path,variants,V", [("auto", "", 0), ("cudacore", "", 0), ("interp", "epi2", 25)])
def test_bench_flow_on_the_emulation(monkeypatch, capsys, path, variants, V):
    """bench.py's whole `ours` arm (session, parity block with the sampled-cell fp64 check, timed blocks, per-kernel profile,
    roofline, late-training timing, fp32-storage measurement, e2e from host buffers, JSON line) executed on the emulated
    library with a small workload: the contract keys are present and consistent.  (Numbers are meaningless here; the point
    is that the code path the driver runs cannot raise.)"""
    import argparse
    import importlib.util
    import json
    import os
    import torch
    from clonealign_b200 import synthetic
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    monkeypatch.setattr(bench, "MIN_TIMED_MS", 0.0)
    monkeypatch.setattr(bench, "LATE_STEPS", 4)
    monkeypatch.setattr(bench, "PARITY_STEPS", 3)

    def fake_cuda(N, G, C, seed=2345234, device="cpu", rows=None, literal=False):
        a, b = rows if rows is not None else (0, N)
        syn = synthetic.make_synthetic(N, G, C, seed=seed)
        return dict(Y=torch.from_numpy(syn["Y"][a:b].astype(np.float32)), L=syn["L"], z=syn["z"][a:b], s=syn["s"][a:b])
    monkeypatch.setattr(synthetic, "make_synthetic_cuda", fake_cuda)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=3, warmup=1, impl="ours", config="c1", y_store="auto", path=path, variants=variants,
                              no_e2e=False, no_cpu_baseline=True, watchdog=0, quick=False)
    cfg = dict(N=300, G=900 if path == "cudacore" else 260, C=4, S=2, name="emulated mini workload")   # G = 900: u8 counts
    if V:
        cfg["V"] = V                         # the allele-specific configuration (BASELINE config 4 in miniature)
    bench.run_ours(args, cfg)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "roofline", "step_hbm", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["value"] > 0 and line["gpu_launches"] > 0
    want = "interp" if path == "auto" else path
    assert line["config"]["path"] == want and line["config"]["y_store"] in ("u8", "u16")
    assert line["config"]["timing"]["blocks"] >= 5 and len(line["config"]["timing"]["ms_blocks"]) == line["config"]["timing"]["blocks"]
    par = line["config"]["parity"]
    assert par["steps"] == 3 and np.isfinite(par["elbo_start"]) and par["elbo_after"] > par["elbo_start"] and len(par["hard_calls_sha256"]) == 16
    assert par["sampled_cell_check"]["ok"], par["sampled_cell_check"]
    assert line["roofline"]["kernel"] == "ypass" and line["roofline"]["bound"] == "hbm" if want == "interp" else True
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(line["e2e"]["seconds"]) == {"upload_and_setup", "gamma_init_and_first_elbo", "loop", "params_download"}
    if line["config"]["y_store"] == "u8":
        assert line["e2e"]["host_dtype"] == "uint8" and line["e2e_f32_host"]["final_elbo"] == line["e2e"]["final_elbo"]
    assert line["alt_fp32_storage"]["y_store"] == "f32" and line["alt_fp32_storage"]["step_hbm"]["bytes_per_step"] > line["step_hbm"]["bytes_per_step"]
    assert line["late_training"]["ms_per_step"] > 0 and line["late_training"]["elbo"] > par["elbo_start"]
    if want == "interp":
        assert line["config"]["panels"]["nb"] >= 1 and line["late_training"]["panels"]["nb"] >= 1


@pytest.mark.parametrize("path", [("cudacore", ""), ("interp", ""), ("interp", "ypass2,epi2,lean")])
def test_long_loop_stays_within_north_star_tolerances(example_sce, path):
    """40 iterations of the reference loop (train + fresh-draw ELBO) against the float64 oracle on identical draws:
    per-iteration ELBO within 1e-4 relative, ML parameters within 1e-3 relative, hard clone assignments identical
    (BASELINE.json north_star) -- i.e. fp32 storage, the fp32 clone softmax of the fused kernel and Adam do not drift."""
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(5))
    d = O.Data(hi["Y"], hi["L"])
    S, n_iter = 2, 40
    eps = np.random.default_rng(9).standard_normal((2 + 2 * n_iter, S, d.Y.shape[1])).astype(np.float32)
    with _session(hi["Y"], hi["L"], hi["psi_init"], hi["mu_guess"], mc_samples=S, K=1, path=path[0], variants=path[1], seed=3) as sess:
        sess.set_eps(eps)
        sess.init_gamma()
        elbos = [sess.elbo()]
        for _ in range(n_iter):
            sess.step()
            elbos.append(sess.elbo())
        prm = sess.params()
    it = iter(eps)
    p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
    r = O.fit(d, p0, lambda: next(it).astype(np.float64), max_iter=n_iter, rel_tol=0.0, n_final=0)
    rel = np.abs(np.array(elbos) - r["elbos"]) / np.abs(r["elbos"])
    assert rel.max() <= ELBO_RTOL, rel.max()
    names = ["A", "B", "C"]
    assert O.clone_assignment(prm["clone_probs"], names) == O.clone_assignment(r["clone_probs"], names)
    assert _relmax(prm["mu"], r["mu"]) <= PARAM_RTOL
    assert _relmax(prm["psi"], r["params"].psi) <= PARAM_RTOL
    assert _relmax(prm["W"], r["params"].W) <= 5e-3
    assert np.abs(prm["clone_probs"] - r["clone_probs"]).max() <= 2e-3


def test_random_shapes_and_variants():
    """Seeded fuzz over (cells, genes, clones, samples, exponent range, kernel set): gradients, Z, F and the ELBO of the
    interp path and its variants against the closed-form oracle, including tiny and degenerate shapes."""
    from clonealign_b200.synthetic import make_synthetic
    rng = np.random.default_rng(20261017)
    for it in range(16):
        small = it % 4 == 0
        N = int(rng.integers(2, 40 if small else 700))
        G = int(rng.integers(2, 70 if small else 3000))
        C = int(rng.integers(1, 6 if small else 33))
        S = max(1, min(int(rng.integers(1, 9)), 128 // C))
        syn = make_synthetic(N, G, C, seed=int(rng.integers(1e6)))
        Y, L = syn["Y"].astype(np.float64), np.minimum(syn["L"], 6.0)
        Y[:, Y.sum(0) == 0] += 1.0
        Y[Y.sum(1) == 0, 0] += 1.0
        d, p, mu_guess, _ = _case(Y, L, K=1, seed=it, scale=float(rng.choice([0.05, 0.3, 1.0])))
        p.psi *= float(rng.choice([0.5, 1.0, 3.0]))
        var = str(rng.choice(["", "ypass2", "epi2", "ypass2,epi2,lean", "ypass2,epi2,lean,overlap", "ypass3", "ypass3,epi2,lean",
                              "ypass2,epi2,lean,defer", "ypass3,epi2,lean,defer,overlap", "ypass4", "ypass4,epi2,lean,defer,cosched",
                              "ypass4,epi2,lean,defer,cosched,cell2", "ypass2,epi2,lean,defer,cell2"]))
        with _session(d.Y, d.L, p.psi, mu_guess, mc_samples=S, K=1, path="interp", seed=1, variants=var) as sess:
            _load_params(sess, p)
            try:
                _check_grads(sess, d, p, S)
            except AssertionError as e:
                raise AssertionError(f"N={N} G={d.Y.shape[1]} C={C} S={S} variants={var!r}: {e}") from None


def test_device_data_stats(example_sce):
    """ca_core_data_stats: s_init, colSums and mu_guess (R/inference-tflow.R:210,117,222) from the resident matrix."""
    import scipy.sparse as sp
    from clonealign_b200.session import DeviceData
    Y, L = example_sce
    Y = Y[:, Y.sum(0) > 0]
    L = L[: Y.shape[1]]
    want_mu = (Y / Y.mean(axis=1, keepdims=True)).mean(axis=0)
    for inp, store in ((Y, "u8"), (sp.csr_matrix(Y), "f32"), (Y * 300.0, "u16")):
        with DeviceData(inp, L, y_store=store) as data:
            st = data.stats()
        scale = 300.0 if store == "u16" else 1.0
        np.testing.assert_array_equal(st["rowsum"], Y.sum(1) * scale)
        np.testing.assert_array_equal(st["colsum"], Y.sum(0) * scale)
        np.testing.assert_allclose(st["mu_guess"], want_mu, rtol=1e-12)


def test_fit_without_host_passes_over_the_matrix(example_sce):
    """Sparse counts + device PCA + device statistics + device correlations + shared inputs: after the (sparse) gene filter
    the host never touches the N x G matrix; the fit agrees with the all-host-initialised one (same RNG stream) to the
    accuracy of the initial values."""
    import scipy.sparse as sp
    from clonealign_b200 import clonealign
    Y, L = example_sce
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(max_iter=4, verbose=False, seed=3, clone_names=["A", "B", "C"])
        cache = {}
        a = clonealign(sp.csr_matrix(Y), L, device_pca=True, device_stats=True, device_correlations=True, cache=cache, **kw)
        cache[("data", 0)].close()
        b = clonealign(Y, L, device_pca=True, **kw)
    ea, eb = a["convergence_info"]["elbo"], b["convergence_info"]["elbo"]
    assert np.abs(ea - eb).max() / np.abs(eb).max() < 1e-6 and a["clone"] == b["clone"]
    np.testing.assert_allclose(a["correlations"], b["correlations"], atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(a["ml_params"]["s"], b["ml_params"]["s"], rtol=0, atol=0)
    with pytest.raises(ValueError, match="device_stats=True needs a cache"):
        clonealign(Y, L, device_stats=True, **kw)


def test_smoke_cases_on_the_emulation(capsys):
    """__graft_entry__.smoke()'s check (one tiny fit step against the oracle) on the emulated library, for the default kernel
    set and the kernel sets that need no tensor core."""
    import __graft_entry__ as g
    for path, variants in (("auto", ""), ("cudacore", ""), ("interp", ""), ("interp", "ypass2,epi2,lean")):
        g._smoke_case(path, variants)
    out = capsys.readouterr().out
    assert out.count(" OK") == 4 and "smoke[auto]" in out and "smoke[interp+ypass2,epi2,lean]" in out


def test_clonealign_batched_final_elbo_is_identical(example_sce):
    """clonealign(batch_final_elbo=True): final_elbo / sd_final_elbo from ca_core_elbo_many equal the default (20 separate
    ca_core_elbo calls) bit for bit, and nothing else of the fit changes."""
    import warnings
    from clonealign_b200 import clonealign
    Y, L = example_sce
    kw = dict(max_iter=4, verbose=False, seed=11, path="interp", variants="ypass3,epi2,lean")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = clonealign(Y, L, **kw)
        b = clonealign(Y, L, batch_final_elbo=True, **kw)
    ca, cb = a["convergence_info"], b["convergence_info"]
    assert ca["final_elbo"] == cb["final_elbo"] and ca["sd_final_elbo"] == cb["sd_final_elbo"] and ca["sd_final_elbo"] > 0
    assert ca["elbo"].tobytes() == cb["elbo"].tobytes() and a["clone"] == b["clone"]


def test_device_preprocess_matches_host_mirror(example_sce):
    """preprocess_for_clonealign (R/preprocess.R:93-147) with its two passes over the matrix as device reductions
    (ca_core_data_stats, ca_core_data_masked_rowsums): same retained genes / cells and matrices as the host
    mirror, on the bundled data (vignette: 6 cells x 67 genes) and on a sparse synthetic matrix."""
    import scipy.sparse as sp
    from clonealign_b200.preprocess import preprocess_for_clonealign
    from clonealign_b200.synthetic import make_synthetic
    Y, L = example_sce
    cases = [(Y, L, {}), (Y.astype(np.uint8), L, dict(min_counts_per_cell=60, nmads=3))]
    syn = make_synthetic(400, 300, 5, seed=4)
    Ls = syn["L"].copy()
    Ls[::7] = 2.0                      # genes with the same copy number in every clone
    Ls[5] = 9.0                        # above max_copy_number
    cases.append((sp.csr_matrix(syn["Y"]), Ls, dict(min_counts_per_gene=150, min_counts_per_cell=2500)))
    for Yc, Lc, kw in cases:
        host = preprocess_for_clonealign(Yc.toarray() if sp.issparse(Yc) else Yc, Lc, **kw)
        dev = preprocess_for_clonealign(Yc, Lc, device=0, **kw)
        assert np.array_equal(dev["retained_genes"], host["retained_genes"]) and np.array_equal(dev["retained_cells"], host["retained_cells"])
        got = dev["gene_expression_data"]
        got = got.toarray() if sp.issparse(got) else np.asarray(got)
        assert np.array_equal(got, host["gene_expression_data"]) and np.array_equal(dev["copy_number_data"], host["copy_number_data"])
        assert 0 < len(host["retained_genes"]) < Yc.shape[1] and 0 < len(host["retained_cells"]) < Yc.shape[0]


def test_bench_restart_mode_on_the_emulation(monkeypatch, capsys):
    """bench.py --restarts (BASELINE config 5: run_clonealign restarts over the devices of one process, replicas and batched
    Y pass) on the emulated library with a miniature workload: the line is produced and both modes fit the same restarts."""
    import argparse
    import importlib.util
    import json
    import os
    import torch
    from clonealign_b200 import synthetic
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)

    def fake_cuda(N, G, C, seed=2345234, device="cpu", rows=None, literal=False):
        syn = synthetic.make_synthetic(N, G, C, seed=seed)
        return dict(Y=torch.from_numpy(syn["Y"].astype(np.float32)), L=syn["L"], z=syn["z"], s=syn["s"])
    monkeypatch.setattr(synthetic, "make_synthetic_cuda", fake_cuda)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=3, restarts=4)
    bench.run_restarts(args, dict(N=150, G=900, C=3, S=1, name="emulated mini restarts"))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["unit"] == "restarts/s" and line["value"] > 0 and line["config"]["restarts"] == 4
    m = line["modes"]
    assert m["replicas"]["restarts"] == m["batched_y_pass"]["restarts"] == 4
    assert abs(m["replicas"]["best_final_elbo"] / m["batched_y_pass"]["best_final_elbo"] - 1.0) < 1e-5
    assert m["batched_y_pass"]["y_bytes_streamed_per_fit_iteration"] * 4 == m["replicas"]["y_bytes_streamed_per_fit_iteration"]
