"""CPU model of the tensor-operand precision choices (no GPU): the reference loop is run with the contraction
operands rounded exactly as the tcgen05 kernels round them (scripts/emulate_precision.py) and compared with the
float64 loop.  It documents WHY the forward kernel splits every column into bf16 hi/lo terms and why one fp16 x fp16
term is enough for the backward kernel ("Precision"); the GPU parity tests check the real kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import emulate_precision as E  # noqa: E402
from clonealign_b200.synthetic import make_synthetic  # noqa: E402
from oracle import clonealign_oracle as O  # noqa: E402


def _setup(N=700, G=400, C=4, S=2, n_iter=5):
    syn = make_synthetic(N, G, C, seed=2345234)
    rng = np.random.default_rng(12345)
    hi = O.host_init(syn["Y"], syn["L"], K=1, rng=rng)
    d = O.Data(hi["Y"], hi["L"])
    eps = [rng.standard_normal((S, d.Y.shape[1])) for _ in range(2 + 2 * n_iter)]
    p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
    return d, p0, eps, n_iter


def _dev(p, ref):
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    return dict(psi=rel(p.psi, ref.psi), W=rel(p.W, ref.W), mu=rel(O.softplus(p.loc), O.softplus(ref.loc)))


def test_chosen_operand_formats_meet_the_parameter_tolerance():
    d, p0, eps, n_iter = _setup()
    exact = {k: "f32" for k in ("Ez", "Mz", "Ezp", "Mzp", "Eb", "Rb")}
    ref = E.loop(d, p0, eps, exact, n_iter)
    # shipped: FWD = 3-term bf16 split on Z and Z' columns, BWD = fp16 x fp16 single term
    shipped = dict(Ez="bf16x2", Mz="bf16x2", Ezp="bf16x2", Mzp="bf16x2", Eb="f16", Rb="f16")
    dv = _dev(E.loop(d, p0, eps, shipped, n_iter), ref)
    assert dv["psi"] < 1e-3 and dv["W"] < 1e-3 and dv["mu"] < 1e-3, dv
    # rejected: plain bf16 on the gradient-only forward columns breaks psi (difference of two ~s_n |w| terms)
    cheap = dict(Ez="bf16x2", Mz="bf16x2", Ezp="bf16", Mzp="bf16", Eb="bf16", Rb="bf16")
    dv2 = _dev(E.loop(d, p0, eps, cheap, n_iter), ref)
    assert dv2["psi"] > 1e-3, dv2
    assert dv2["psi"] > 20 * dv["psi"]
