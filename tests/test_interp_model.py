"""CPU model of the "interp" path (clonealign_b200/csrc/kernels_interp.cuh): with K = 1 both contractions are
univariate functions of psi_n (forward) and w_g (backward); piecewise Chebyshev interpolation with 24 nodes per
panel reproduces the direct float64 contraction to ~1e-13.  scripts/interp_prototype.py uses the same node, DCT and
Clenshaw formulas as the kernels."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import interp_prototype as IP  # noqa: E402


@pytest.mark.parametrize("sd_psi,sd_w", [(1.0, 0.3), (1.5, 1.0), (2.0, 2.0), (1.0, 0.0)])
def test_univariate_collapse_matches_direct_contraction(sd_psi, sd_w):
    rng = np.random.default_rng(0)
    N, G, J = 1500, 1200, 16
    psi = rng.standard_normal(N) * sd_psi
    w = rng.standard_normal(G) * sd_w
    Mx = rng.uniform(0.1, 5.0, size=(G, J))
    Mx[:, J // 2:] *= w[:, None]                      # Z' columns carry w_g
    Rx = rng.uniform(0.0, 1.0, size=(N, J))
    Rx[:, J // 2:] *= psi[:, None]                    # dM' columns carry psi_n
    m = np.maximum(psi * w.max(), psi * w.min())
    E = np.exp(psi[:, None] * w[None, :] - m[:, None])
    Z_ref, dM_ref = E @ Mx, E.T @ Rx
    Z, nf = IP.forward_interp(psi, w, Mx, P=24)
    dM, nb = IP.backward_interp(psi, w, Rx, P=24)
    assert np.abs(Z[:, :J // 2] / Z_ref[:, :J // 2] - 1.0).max() < 1e-12
    scale = np.abs(Z_ref[:, :J // 2]).max(axis=1, keepdims=True) * max(np.abs(w).max(), 1e-300)
    assert (np.abs(Z - Z_ref)[:, J // 2:] / scale).max() < 1e-12
    assert (np.abs(dM - dM_ref) / np.abs(dM_ref).max(axis=0, keepdims=True).clip(1e-300)).max() < 1e-12
    # the whole point: a few hundred node rows instead of N (or G) rows
    assert nf <= 24 * 64 and nb <= 24 * 64
    if sd_w <= 0.3:
        assert nf <= 96 and nb <= 48


def test_one_sided_and_degenerate_ranges():
    rng = np.random.default_rng(1)
    G, J = 300, 8
    w = rng.standard_normal(G) * 0.5
    Mx = rng.uniform(0.1, 2.0, size=(G, J))
    for psi in (np.abs(rng.standard_normal(200)), -np.abs(rng.standard_normal(200)), np.zeros(50)):
        m = np.maximum(psi * w.max(), psi * w.min())
        ref = np.exp(psi[:, None] * w[None, :] - m[:, None]) @ Mx
        Z, _ = IP.forward_interp(psi, w, Mx, P=24)
        assert np.abs(Z / ref - 1.0).max() < 1e-12


@pytest.mark.parametrize("dist", ["normal", "uniform", "spiky"])
def test_sixteen_nodes_at_half_range_four(dist):
    """The kernels use kIP = 16 nodes per panel with exponent half-range kIAmax = 4: worst relative error of the
    interpolated normaliser over a panel, for gene-weight distributions from near point masses to uniform."""
    rng = np.random.default_rng(0)
    G, P, A, D = 5000, 16, 4.0, 2.0
    if dist == "normal":
        w = rng.standard_normal(G) * D / 6
    elif dist == "uniform":
        w = rng.uniform(-D / 2, D / 2, G)
    else:
        w = np.concatenate([rng.standard_normal(G - 5) * 0.01, np.array([1.0, -1.0, 0.9, -0.8, 0.7]) * D / 2])
    M = rng.uniform(0.5, 5, G)
    h = A / (w.max() - w.min())
    worst = 0.0
    for pan in range(5):
        mid = pan * 2 * h + h
        nodes = mid + h * np.cos(np.pi * (np.arange(P) + 0.5) / P)
        f = np.exp(nodes[:, None] * (w - w.max())[None, :]) @ M
        k, p = np.arange(P)[:, None], np.arange(P)[None, :]
        c = (2.0 / P) * (f[None, :] * np.cos(np.pi * k * (p + 0.5) / P)).sum(1)
        c[0] *= 0.5
        xs = np.linspace(mid - h, mid + h, 101)
        val = np.polynomial.chebyshev.chebval((xs - mid) / h, c)
        ref = np.exp(xs[:, None] * (w - w.max())[None, :]) @ M
        worst = max(worst, np.abs(val / ref - 1.0).max())
    assert worst < 2e-10, worst
