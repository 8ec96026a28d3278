"""Synthetic benchmark inputs: vectorised port of the generator at inst/create_model3_synthetic.R:3-39.

    rho_g ~ Bernoulli(.9/1.1) (:8, R normalises prob=c(.2,.9));  z_n ~ U{1..C} (:10);  mu_g ~ U(1,2) (:12)
    phi_g ~ Gamma(4,1) (:14);  L_gc ~ U{1..C} (:16);  Lp = L / colMeans(L) (:17);  s_n ~ U(500,10000) (:19)
    y_ng ~ NegBin(mean = m_ng, size = phi_g),  relative rate r_ng = (1-rho_g) mu_g + rho_g mu_g Lp[g, z_n] (:26-27)

Documented deviation: the script uses m_ng = s_n * r_ng *per gene* (row totals ~1e8 at
G = 20k); the benchmark uses m_ng = s_n * r_ng / sum_g r_ng so that s_n is the library size (mean
~0.26 counts/entry at G = 20k, ~80 % zeros).  `literal=True` gives the script's own scaling.
The model input is the integer copy-number matrix L (what `clonealign()` expects), not Lp.
"""
from __future__ import annotations

import numpy as np


def gene_level(N, G, C, seed=2345234):
    rng = np.random.default_rng(seed)
    rho = (rng.random(G) < 0.9 / 1.1).astype(np.float64)
    z = rng.integers(0, C, size=N)
    mu = rng.uniform(1.0, 2.0, size=G)
    phi = rng.gamma(4.0, 1.0, size=G)
    L = rng.integers(1, C + 1, size=(G, C)).astype(np.float64)
    Lp = L / L.mean(axis=0, keepdims=True)
    s = rng.uniform(500.0, 10000.0, size=N)
    return dict(rho=rho, z=z, mu=mu, phi=phi, L=L, Lp=Lp, s=s, rng=rng)


def make_synthetic(N, G, C, seed=2345234, literal=False):
    """numpy generator (small / medium shapes).  Returns dict(Y float32 (N,G), L (G,C), z, s)."""
    p = gene_level(N, G, C, seed)
    rng = p["rng"]
    Y = np.empty((N, G), dtype=np.float32)
    step = max(1, min(N, (1 << 24) // max(G, 1)))
    for a in range(0, N, step):
        b = min(N, a + step)
        r = (1.0 - p["rho"])[None] * p["mu"][None] + p["rho"][None] * p["mu"][None] * p["Lp"][:, p["z"][a:b]].T
        m = p["s"][a:b, None] * (r if literal else r / r.sum(axis=1, keepdims=True))
        lam = rng.gamma(p["phi"][None], m / p["phi"][None])
        Y[a:b] = rng.poisson(lam).astype(np.float32)
    _fix_empty(Y)
    return dict(Y=Y, L=p["L"], z=p["z"], s=p["s"])


def _fix_empty(Y):
    # shapes must be exact: the gene filter (R/inference-tflow.R:117) must retain every gene, and no cell
    # may be empty (:212).  Vanishingly rare at benchmark sizes; matters for tiny test shapes.
    cs = Y.sum(axis=0)
    for g in np.nonzero(cs == 0)[0]:
        Y[g % Y.shape[0], g] = 1.0
    rs = Y.sum(axis=1)
    for n in np.nonzero(rs == 0)[0]:
        Y[n, n % Y.shape[1]] = 1.0


SYN_BLOCK = 1024   # rows per independently seeded block of the CUDA generator


def make_synthetic_cuda(N, G, C, seed=2345234, device="cuda:0", rows=None, literal=False):
    """Same model drawn on the GPU with torch (plumbing only: the 100k x 20k matrix would take minutes on
    the host).  `rows=(a, b)` draws only that cell shard.  The draw is SHARD-INVARIANT: gene-level draws are identical
    for every shard and the counts come in blocks of SYN_BLOCK rows, each from its own generator seeded by
    (seed, global block index), so the matrix a rank holds is exactly rows [a, b) of the matrix one GPU would draw --
    a sharded fit and the unsharded one see the same data (1-vs-k-GPU parity runs).
    Returns dict(Y torch.float32 cuda (b-a, G), L numpy (G,C), z, s)."""
    import torch
    p = gene_level(N, G, C, seed)
    a, b = rows if rows is not None else (0, N)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    t = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    rho, mu, phi, Lp = t(p["rho"]), t(p["mu"]), t(p["phi"]), t(p["Lp"])
    z_all = torch.tensor(p["z"], device=dev)
    s_all = t(p["s"])
    Y = torch.empty((b - a, G), dtype=torch.float32, device=dev)
    phi_b = phi[None].expand(SYN_BLOCK, G).contiguous()
    for blk in range(a // SYN_BLOCK, (b + SYN_BLOCK - 1) // SYN_BLOCK):
        r0 = blk * SYN_BLOCK
        r1 = min(N, r0 + SYN_BLOCK)
        gen.manual_seed(int(seed) * 1000003 + 7919 * (blk + 1))
        zz = torch.zeros(SYN_BLOCK, dtype=z_all.dtype, device=dev)
        ss = torch.ones(SYN_BLOCK, dtype=torch.float32, device=dev)
        zz[: r1 - r0] = z_all[r0:r1]
        ss[: r1 - r0] = s_all[r0:r1]
        r = ((1.0 - rho) * mu)[None] + (rho * mu)[None] * Lp[:, zz].T
        m = ss[:, None] * (r if literal else r / r.sum(dim=1, keepdim=True))
        # Gamma(shape=phi, scale=m/phi) via standard gamma; then Poisson.  Always a full block: the draws of a row do not
        # depend on where the shard boundaries fall.
        g0 = torch._standard_gamma(phi_b, generator=gen)
        lam = g0 * (m / phi[None])
        yb = torch.poisson(lam, generator=gen)
        lo, hi = max(a, r0), min(b, r1)
        Y[lo - a:hi - a] = yb[lo - r0:hi - r0]
        del r, m, g0, lam, yb
    # exact shapes (see _fix_empty): no empty cell (row-local rule, shard-invariant)
    rs = Y.sum(dim=1)
    bad = torch.nonzero(rs == 0).flatten().tolist()
    for n in bad:
        Y[n, (n + a) % G] = 1.0
    return dict(Y=Y, L=p["L"], z=p["z"][a:b], s=p["s"][a:b])
