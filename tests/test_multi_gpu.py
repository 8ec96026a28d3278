"""GPU tests of the multi-GPU entry points: one fit sharded by cells over the GPUs of ONE process (ca_core_multi_*,
`MultiSession`, `clonealign(devices=...)`: what `options(clonealign.gpus=)` gives an R user) and restarts
spread over devices (`run_clonealign(devices=...)`, R/clonealign.R:50-56).  Tests that need two devices are skipped on a
one-GPU box; the one-device cases run everywhere.
"""
import warnings

import numpy as np
import pytest

from oracle import clonealign_oracle as O

pytestmark = pytest.mark.gpu


def _ndev():
    from clonealign_b200.session import Session
    return Session.device_count()


def _trace(sess, n=4):
    sess.init_gamma()
    tr = [sess.elbo()]
    for _ in range(n):
        sess.step()
        tr.append(sess.elbo())
    tr += list(sess.elbo_many(2))
    return np.array(tr), sess.params()


def test_multisession_on_one_device_is_the_plain_session(example_sce):
    """n = 1: the worker-thread plumbing adds nothing -- bit-identical to `Session`, for every input layout."""
    import scipy.sparse as sp
    from clonealign_b200.session import MultiSession, Session
    Y, L = example_sce
    hi = O.host_init(Y, L, K=1, rng=np.random.default_rng(0))
    loc = O.safe_inverse_softplus(hi["mu_guess"])
    with Session(hi["Y"], hi["L"], hi["psi_init"], loc, mc_samples=2, seed=5) as s:
        e_ref, p_ref = _trace(s)
    for Yin in (hi["Y"].astype(np.uint8), np.asfortranarray(hi["Y"]), sp.csr_matrix(hi["Y"])):
        with MultiSession(Yin, hi["L"], hi["psi_init"], loc, devices=[0], mc_samples=2, seed=5) as m:
            e, p = _trace(m)
            assert m.describe()["shards"] == [(0, hi["Y"].shape[0])]
        assert e.tobytes() == e_ref.tobytes()
        for k in p_ref:
            assert p[k].tobytes() == p_ref[k].tobytes(), k


@pytest.mark.parametrize("variants", ["", "p2p"])
def test_cells_sharded_over_two_gpus_match_one_gpu(variants):
    """SURVEY 8e on hardware: the fit with its cells on two GPUs (one all-reduce of the gene-level gradients per step) against
    the same fit on one GPU, same data, seeds and draws: ELBO trace within fp32 re-association (1e-6), identical hard clone
    calls, per-cell parameters = rows of the one-GPU ones.  `p2p`: the all-reduce as one kernel over NVLink peer memory."""
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    from clonealign_b200.session import MultiSession, Session
    from clonealign_b200.synthetic import make_synthetic
    N, G, C, S = 6000, 1500, 6, 4
    syn = make_synthetic(N, G, C, seed=11)
    Y = syn["Y"].astype(np.float64)
    L = np.minimum(syn["L"], 6.0)
    rng = np.random.default_rng(3)
    psi = rng.standard_normal((N, 1))
    mu_guess = (Y / Y.mean(axis=1, keepdims=True)).mean(axis=0)
    loc = O.safe_inverse_softplus(mu_guess)
    kw = dict(mc_samples=S, seed=9)
    with Session(Y, L, psi, loc, **kw) as s:
        e1, p1 = _trace(s, n=12)
    with MultiSession(Y, L, psi, loc, devices=[0, 1], variants=variants, **kw) as m:
        d = m.describe()
        assert d["world"] == 2 and d["path"] == "interp" and d["shards"] == [(0, N // 2), (N // 2, N)]
        e2, p2 = _trace(m, n=12)
    assert np.all(np.isfinite(e2)) and (np.abs(e2 - e1) / np.abs(e1)).max() < 1e-6
    call = lambda cp: np.where(cp.max(axis=1) < 0.95, -1, cp.argmax(axis=1))
    assert np.array_equal(call(p2["clone_probs"]), call(p1["clone_probs"]))
    assert np.abs(p2["clone_probs"] - p1["clone_probs"]).max() < 1e-4
    assert np.abs(p2["psi"] - p1["psi"]).max() <= 1e-4 * np.abs(p1["psi"]).max()
    assert np.abs(p2["mu"] - p1["mu"]).max() <= 1e-4 * np.abs(p1["mu"]).max()
    assert np.abs(p2["W"] - p1["W"]).max() <= 1e-3 * np.abs(p1["W"]).max() + 1e-6


def test_clonealign_with_devices_and_restarts_over_two_gpus(example_sce):
    """clonealign(devices=[0, 1]) (one fit on two GPUs) and run_clonealign(devices=[0, 1]) (restarts spread over GPUs):
    same clone calls / the same best restart as on one device."""
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    from clonealign_b200 import clonealign, run_clonealign
    Y, L = example_sce
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = clonealign(Y, L, max_iter=8, clone_names=["A", "B", "C"], verbose=False, seed=1)
        b = clonealign(Y, L, max_iter=8, clone_names=["A", "B", "C"], verbose=False, seed=1, devices=[0, 1])
        kw = dict(initial_shrinks=(0, 5), n_repeats=2, print_elbos=False, max_iter=6, verbose=False, seed=3)
        r1 = run_clonealign(Y, L, **kw)
        r2 = run_clonealign(Y, L, devices=[0, 1], **kw)
        r3 = run_clonealign(Y, L, devices=[0, 1], share_inputs=True, batch_y_pass=True, **kw)
    assert a["clone"] == b["clone"]
    assert np.abs(b["convergence_info"]["elbo"] / a["convergence_info"]["elbo"] - 1.0).max() < 1e-6
    assert r1["multirun_info"]["elbos"].tobytes() == r2["multirun_info"]["elbos"].tobytes() and r1["clone"] == r2["clone"]
    assert np.abs(r3["multirun_info"]["elbos"] / r1["multirun_info"]["elbos"] - 1.0).max() < 1e-5 and r3["clone"] == r1["clone"]
