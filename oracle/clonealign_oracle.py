"""CPU oracle for clonealign's variational hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this module.  The product (`clonealign_b200/`) never does.

PARITY UNPINNED.  The reference (kieranrcampbell/clonealign v1.99.3) is an R package that
builds a TensorFlow-1 graph through `reticulate`; its arithmetic lives in the un-vendored
third-party wheels `tensorflow==2.1.0` (README.md:35, .travis.yml:38) and
`tensorflow-probability` (unpinned, 0.9.x line).  Neither R nor TensorFlow exists in this
image, the reference ships no golden vectors for this path (tests/testthat/test_clonealign.R
checks shapes and seed determinism only), so this file is a *restatement* of the graph at
R/inference-tflow.R:238-346 using the published semantics of the TF/TFP ops it calls.
The only recorded reference numbers (rendered vignette, docs/introduction_to_clonealign.html
:816-819,908) are used as a loose sanity regime in tests/test_oracle.py.

Two independent implementations live here and are asserted equal in the tests:

* `elbo_tfgraph`      literal op-for-op restatement of the TF graph (einsum chain that
                      materialises the (S,G,C,N) tensors), in torch so that autograd plays
                      the role of `optimizer$minimize(-elbo)`'s autodiff.
* `elbo_grads_closed` numpy float64, factorised form + closed-form gradients (the algebra
                      the CUDA kernels implement).

plus the TF1 Adam update, the reference optimisation loop and the host-side initialisation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------
# small helpers  (R/inference-tflow.R:2-29, R/clonealign.R:394-397)
# ----------------------------------------------------------------------------------------
def saturate(x, threshold=4.0):
    """R/clonealign.R:394-397 — clip values above `threshold`."""
    x = np.array(x, dtype=np.float64, copy=True)
    x[x > threshold] = threshold
    return x


def softplus(x):
    """R/inference-tflow.R:13-15 (computed stably)."""
    x = np.asarray(x, dtype=np.float64)
    return np.logaddexp(0.0, x)


def safe_inverse_softplus(x):
    """R/inference-tflow.R:6-11."""
    x = np.asarray(x, dtype=np.float64)
    if np.any(x < 0):
        raise ValueError("Inverse softplus only takes positive values")
    return np.log(1.0 - np.exp(-np.abs(x))) + np.maximum(x, 0.0)


def clone_assignment(gamma, clone_names, clone_assignment_probability=0.95):
    """R/inference-tflow.R:22-29 — argmax clone, "unassigned" when max prob < threshold."""
    gamma = np.asarray(gamma)
    out = []
    for r in gamma:
        if r.max() < clone_assignment_probability:
            out.append("unassigned")
        else:
            out.append(clone_names[int(np.argmax(r))])
    return out


def _lgamma(x):
    from scipy.special import gammaln
    return gammaln(x)


# ----------------------------------------------------------------------------------------
# allele-specific likelihood  (R/allele-specific.R:17-58)
# ----------------------------------------------------------------------------------------
def beta_binomial_log_prob(k, n, alpha, beta):
    """R/allele-specific.R:52-58."""
    ll = _lgamma(n + 1) - _lgamma(k + 1) - _lgamma(n - k + 1)
    ll = ll + _lgamma(k + alpha) + _lgamma(n - k + beta) - _lgamma(alpha + beta + n)
    ll = ll - _lgamma(alpha) - _lgamma(beta) + _lgamma(alpha + beta)
    return ll


def construct_ai_likelihood(clone_allele, alt, cov):
    """R/allele-specific.R:17-48, literal (C,V,N) stacking.

    clone_allele: (V,C) copy number at each variant; alt, cov: (V,N).  Returns (N,C).
    """
    clone_allele = np.asarray(clone_allele, dtype=np.float64)
    alt = np.asarray(alt, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    V, C = clone_allele.shape
    N = alt.shape[1]
    p1_low = math.log(0.5) + beta_binomial_log_prob(alt, cov, 0.1, 1.9)
    p1_high = math.log(0.5) + beta_binomial_log_prob(alt, cov, 1.9, 0.1)
    p1 = np.logaddexp(p1_low, p1_high)                      # reduce_logsumexp over the stack
    p2 = beta_binomial_log_prob(alt, cov, 2.0, 2.0)
    p1c = np.broadcast_to(p1[None], (C, V, N))              # clone by variant by cell
    p2c = np.broadcast_to(p2[None], (C, V, N))
    is2 = (clone_allele == 2)                               # (V,C)
    is2 = np.broadcast_to(is2[None], (N, V, C)).transpose(2, 1, 0)   # (C,V,N)
    Lm = np.where(is2, p2c, p1c)
    return Lm.sum(axis=1).T                                 # (N,C)


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
PARAM_NAMES = ("W", "chi_raw", "psi", "beta", "alpha_unconstr", "loc", "lsd", "gamma_logits")


@dataclass
class Params:
    """Trainable variables of R/inference-tflow.R:240-272 (float64 numpy arrays)."""
    W: np.ndarray               # (G,K)   :240
    chi_raw: np.ndarray         # (K,)    :241   chi = exp(chi_raw)
    psi: np.ndarray             # (N,K)   :242
    beta: np.ndarray            # (G,P)   :245
    alpha_unconstr: np.ndarray  # (C,)    :254
    loc: np.ndarray             # (G,)    :262   qmu Normal loc (pre-softplus)
    lsd: np.ndarray             # (G,)    :263   qmu Normal log-scale
    gamma_logits: np.ndarray    # (N,C)   :272

    def copy(self):
        return Params(**{k: getattr(self, k).copy() for k in PARAM_NAMES})

    def asdict(self):
        return {k: getattr(self, k) for k in PARAM_NAMES}


@dataclass
class Data:
    Y: np.ndarray               # (N,G) counts
    L: np.ndarray               # (G,C) copy number (already saturated)
    s: np.ndarray = None        # (N,) library sizes rowSums(Y)   :210
    X: np.ndarray = None        # (N,P) covariates or None
    v: np.ndarray = None        # (N,C) allele log-lik (construct_ai_likelihood) or None

    def __post_init__(self):
        self.Y = np.asarray(self.Y, dtype=np.float64)
        self.L = np.asarray(self.L, dtype=np.float64)
        if self.s is None:
            self.s = self.Y.sum(axis=1)
        if self.X is not None:
            self.X = np.asarray(self.X, dtype=np.float64)
            if self.X.ndim == 1:
                self.X = self.X[:, None]


def init_params(Y, L, psi_init, mu_guess, K=1, P=0):
    """Initial values of R/inference-tflow.R:240-272."""
    N, G = Y.shape
    C = L.shape[1]
    return Params(
        W=np.zeros((G, K)), chi_raw=np.zeros(K), psi=np.array(psi_init, dtype=np.float64).reshape(N, K),
        beta=np.zeros((G, P)), alpha_unconstr=np.zeros(C),
        loc=safe_inverse_softplus(mu_guess), lsd=np.zeros(G), gamma_logits=np.zeros((N, C)))


# ----------------------------------------------------------------------------------------
# host-side initialisation  (R/inference-tflow.R:117-144, 204-235)
# ----------------------------------------------------------------------------------------
def host_init(Y_dat, L_dat, K=1, gene_filter_threshold=0, do_saturate=True, saturation_threshold=6,
              data_init_mu=True, rng=None, psi_noise_sd=0.05):
    """Gene filter, saturation, PCA psi init (+N(0,.05^2) noise), s, mu_guess.

    Returns dict(Y, L, retained (bool mask), psi_init (N,K), s, mu_guess).
    The noise comes from `rng` (numpy Generator) instead of R's RNG (:208).
    """
    Y = np.asarray(Y_dat, dtype=np.float64)
    L = np.asarray(L_dat, dtype=np.float64)
    zero_gene_means = Y.sum(axis=0) <= gene_filter_threshold          # :117
    Y = Y[:, ~zero_gene_means]                                        # :123
    L = L[~zero_gene_means, :]                                        # :124
    if do_saturate:
        L = saturate(L, saturation_threshold)                         # :142-144
    N, G = Y.shape
    s = Y.sum(axis=1)                                                 # :210
    if np.any(s == 0):
        raise ValueError("Some cells have no counts mapping")         # :212-214
    psi = np.zeros((N, 0))
    if K > 0:
        X = np.log2(Y + 1.0)                                          # :204
        X = X - X.mean(axis=0)
        sd = X.std(axis=0, ddof=1)
        if np.any(sd == 0):
            raise ValueError("cannot rescale a constant/zero column to unit variance")
        X = X / sd
        U, S, Vt = np.linalg.svd(X, full_matrices=False)
        pcs = (U * S)[:, :K]                                          # :205
        pcs = (pcs - pcs.mean(axis=0)) / pcs.std(axis=0, ddof=1)      # :206 scale()
        if rng is not None and psi_noise_sd > 0:
            pcs = pcs + rng.normal(0.0, psi_noise_sd, size=pcs.shape)  # :208
        psi = pcs
    if data_init_mu is True:
        mu_guess = (Y / Y.mean(axis=1, keepdims=True)).mean(axis=0)   # :222
    elif data_init_mu is False:
        mu_guess = np.ones(G)                                         # :224
    else:
        d = np.asarray(data_init_mu, dtype=np.float64)
        mu_guess = d / d.mean()                                       # :232
    return dict(Y=Y, L=L, retained=~zero_gene_means, psi_init=psi, s=s, mu_guess=mu_guess)


# ----------------------------------------------------------------------------------------
# (1) literal TF-graph restatement, torch + autograd
# ----------------------------------------------------------------------------------------
def elbo_tfgraph(params, data, eps, dtype=None, want_grads=True, want_gamma_init=False):
    """Restates R/inference-tflow.R:240-346 op for op (the (S,G,C,N) einsum chain).

    eps: (S,G) standard-normal draws standing in for `qmu$sample(S, seed)` (:269), which is
    reparameterised: x = loc + exp(lsd) * eps, mu = softplus(x).
    Returns dict(elbo=float, grads={name: ndarray of d(elbo)/d(param)}, gamma_init=(N,C)).
    """
    import torch
    dtype = dtype or torch.float64
    tt = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
    p = {k: tt(v).requires_grad_(want_grads) for k, v in params.asdict().items()}
    Y, L, s = tt(data.Y), tt(data.L), tt(data.s)
    eps_t = tt(eps)
    N, G = Y.shape
    C = L.shape[1]
    S = eps_t.shape[0]
    K = p["W"].shape[1]
    P = p["beta"].shape[1]

    chi = torch.exp(p["chi_raw"])                                            # :241
    log_alpha = torch.log_softmax(p["alpha_unconstr"], dim=0)                # :255
    x = p["loc"] + torch.exp(p["lsd"]) * eps_t                               # :260-269 (reparam. sample)
    mu_samples = torch.nn.functional.softplus(x, threshold=60.0)             # (S,G); no linear shortcut
    gamma = torch.softmax(p["gamma_logits"], dim=1)                          # :273

    if K > 0 and P == 0:                                                     # :279-285
        rfe = torch.exp(p["psi"] @ p["W"].T)
    elif K > 0 and P > 0:
        rfe = torch.exp(p["psi"] @ p["W"].T + tt(data.X) @ p["beta"].T)
    else:
        rfe = torch.ones((N, G), dtype=dtype)

    mu_scg = torch.einsum("sg,gc->scg", mu_samples, L)                       # :288
    mu_sgcn = torch.einsum("scg,ng->sgcn", mu_scg, rfe)                      # :289
    norm = 1.0 / mu_sgcn.sum(dim=1)                                          # :290 (S,C,N)
    mu_sgcn_norm = torch.einsum("sgcn,scn->sgcn", mu_sgcn, norm)             # :291
    mu_scng = mu_sgcn_norm.permute(0, 2, 3, 1)                               # :292

    # tfd$Multinomial(total_count = s, probs)$log_prob(Y)   :294-296
    #   = sum(counts * log_softmax(log probs)) + lgamma(n+1) - sum(lgamma(counts+1))
    logp = torch.log_softmax(torch.log(mu_scng), dim=-1)
    log_unnorm = (Y * logp).sum(dim=-1)                                      # (S,C,N)
    log_comb = torch.lgamma(s + 1.0) - torch.lgamma(Y + 1.0).sum(dim=-1)     # (N,)
    p_y_on_c = log_unnorm + log_comb
    if data.v is not None:                                                   # :302-304
        p_y_on_c = p_y_on_c + tt(data.v).T
    E_p_y_on_c = p_y_on_c.mean(dim=0)                                        # :306 (C,N)
    EE_p_y = (gamma * E_p_y_on_c.T).sum()                                    # :308

    def normal_lp(xv, scale):
        return -0.5 * (xv / scale) ** 2 - torch.log(scale) - 0.5 * LOG2PI

    one = torch.ones(1, dtype=dtype)
    conc = torch.full((C,), 1.0 / C, dtype=dtype)
    dir_x = torch.exp(log_alpha) + 1e-3
    dirichlet_lp = ((conc - 1.0) * torch.log(dir_x)).sum() - (torch.lgamma(conc).sum() - torch.lgamma(conc.sum()))
    E_log_p_p = (log_alpha * gamma).sum() \
        + normal_lp(torch.log(mu_samples), one).sum() / float(S) \
        + dirichlet_lp                                                       # :322-324
    if K > 0:                                                                # :311-328
        W_log_prob = normal_lp(p["W"], torch.sqrt(one / chi)).sum()
        chi_log_prob = (torch.log(chi) - chi).sum()                          # Gamma(2,1): (a-1)log x - b x - lgamma(a) + a log b
        p_psi = normal_lp(p["psi"], one).sum()
        E_log_p_p = E_log_p_p + W_log_prob + chi_log_prob + p_psi

    # qmu$log_prob(mu_samples): Normal(loc, scale).log_prob(x) - log|d softplus/dx|(x),  x the pre-softplus value
    scale = torch.exp(p["lsd"])
    q_lp = -0.5 * ((x - p["loc"]) / scale) ** 2 - torch.log(scale) - 0.5 * LOG2PI \
        - torch.nn.functional.logsigmoid(x)
    log_gamma = torch.log_softmax(p["gamma_logits"], dim=1)
    ent = torch.where(gamma == 0, torch.zeros_like(gamma), gamma * log_gamma)
    E_log_q = q_lp.mean(dim=0).sum() + ent.sum()                             # :332-333
    elbo = EE_p_y + E_log_p_p - E_log_q                                      # :336

    out = {"elbo": float(elbo.detach())}
    if want_gamma_init:                                                      # :338-340
        gi = p_y_on_c.sum(dim=0)
        gi = gi - torch.logsumexp(gi, dim=0)
        out["gamma_init"] = gi.T.detach().numpy().astype(np.float64)
    if want_grads:
        elbo.backward()
        out["grads"] = {k: (v.grad.detach().numpy().astype(np.float64) if v.grad is not None
                            else np.zeros(tuple(v.shape))) for k, v in p.items()}
    return out


# ----------------------------------------------------------------------------------------
# (2) factorised form + closed-form gradients, numpy float64  
# ----------------------------------------------------------------------------------------
def precompute(data):
    """Y-only constants: B = Y log L (N,C), const_n, colsum_g."""
    Y, L = data.Y, data.L
    with np.errstate(divide="ignore", invalid="ignore"):
        logL = np.log(L)
        # 0 * log 0 = NaN, as in the reference: keep IEEE semantics
        B = Y @ logL if np.all(np.isfinite(logL)) else (Y[:, :, None] * logL[None]).sum(axis=1)
    const = _lgamma(data.s + 1.0) - _lgamma(Y + 1.0).sum(axis=1)
    return dict(B=B, const=const, colsum=Y.sum(axis=0))


def elbo_grads_closed(params, data, eps, pre=None, want_grads=True):
    """ELBO, its gradients (ascent direction, d ELBO / d param) and gamma_init in closed form."""
    pre = pre or precompute(data)
    Y, L, s = data.Y, data.L, data.s
    N, G = Y.shape
    C = L.shape[1]
    eps = np.asarray(eps, dtype=np.float64)
    S = eps.shape[0]
    W, psi, beta = params.W, params.psi, params.beta
    K, P = W.shape[1], beta.shape[1]

    sig_q = np.exp(params.lsd)
    x = params.loc + sig_q * eps                       # (S,G)
    mu = softplus(x)
    logmu = np.log(mu)
    sgm = 1.0 / (1.0 + np.exp(-x))                     # sigmoid(x)
    log_sgm = -np.logaddexp(0.0, -x)

    eta = np.zeros((N, G))
    if K > 0:
        eta = eta + psi @ W.T
    if P > 0 and K > 0:   # reference quirk (:279-285): with K == 0 the covariates are silently ignored
        eta = eta + data.X @ beta.T
    m = eta.max(axis=1)                                # (N,) shift
    E = np.exp(eta - m[:, None])
    M = mu[:, :, None] * L[None]                       # (S,G,C)
    Z = np.einsum("ng,sgc->scn", E, M)                 # shifted normaliser
    logZ = np.log(Z) + m[None, None, :]

    v = data.v if data.v is not None else np.zeros((N, C))
    ell_c = pre["B"].T[None] - s[None, None, :] * logZ + v.T[None]       # clone-dependent part (S,C,N)
    F = ell_c.mean(axis=0).T                                              # (N,C)

    t = params.gamma_logits
    tmax = t.max(axis=1, keepdims=True)
    log_gamma = t - tmax - np.log(np.exp(t - tmax).sum(axis=1, keepdims=True))
    gamma = np.exp(log_gamma)
    u = params.alpha_unconstr
    log_alpha = u - u.max() - np.log(np.exp(u - u.max()).sum())
    alpha = np.exp(log_alpha)
    chi = np.exp(params.chi_raw)

    ent = np.where(gamma == 0, 0.0, gamma * log_gamma)
    y_eta = (Y * eta).sum()                                              # sum_ng y eta
    y_logmu = (pre["colsum"][None] * logmu).sum() / S
    elbo = (gamma * F).sum() + y_eta + y_logmu + pre["const"].sum()
    elbo += (gamma * log_alpha[None]).sum()
    elbo += (-0.5 * logmu ** 2 - 0.5 * LOG2PI).sum() / S
    elbo += (1.0 / C - 1.0) * np.log(alpha + 1e-3).sum() - (C * math.lgamma(1.0 / C) - math.lgamma(1.0))
    if K > 0:
        elbo += (-0.5 * chi[None] * W ** 2 + 0.5 * np.log(chi)[None] - 0.5 * LOG2PI).sum()
        elbo += (np.log(chi) - chi).sum()
        elbo += (-0.5 * psi ** 2 - 0.5 * LOG2PI).sum()
    elbo -= (-0.5 * eps ** 2 - params.lsd[None] - 0.5 * LOG2PI - log_sgm).sum() / S
    elbo -= ent.sum()

    out = {"elbo": float(elbo), "gamma": gamma, "F": F, "Z": Z, "m": m, "mu": mu}
    # gamma_init (:338-340): SUM over s, clone-independent terms cancel in the log-softmax
    gi = ell_c.sum(axis=0)                                                # (C,N)
    gmax = gi.max(axis=0, keepdims=True)
    gi = gi - (gmax + np.log(np.exp(gi - gmax).sum(axis=0, keepdims=True)))
    out["gamma_init"] = gi.T
    if not want_grads:
        return out

    H = F + log_alpha[None] - np.where(gamma == 0, 0.0, log_gamma)
    g_t = gamma * (H - (gamma * H).sum(axis=1, keepdims=True))
    R = gamma.T[None] * s[None, None, :] / (S * Z)                        # (S,C,N)
    dM = np.einsum("ng,scn->sgc", E, R)
    Q = np.einsum("sgc,scn->ng", M, R)
    d_mu = pre["colsum"][None] / (S * mu) - (L[None] * dM).sum(axis=2) - logmu / (S * mu)
    d_x = sgm * d_mu + (1.0 - sgm) / S
    g_loc = d_x.sum(axis=0)
    g_lsd = (d_x * sig_q[None] * eps).sum(axis=0) + 1.0
    d_eta = Y - E * Q
    g_psi = d_eta @ W - psi if K > 0 else np.zeros((N, 0))
    g_W = d_eta.T @ psi - chi[None] * W if K > 0 else np.zeros((G, 0))
    g_beta = d_eta.T @ data.X if (P > 0 and K > 0) else np.zeros((G, P))
    g_chi = (-0.5 * chi * (W ** 2).sum(axis=0) + G / 2.0 + 1.0 - chi) if K > 0 else np.zeros(0)
    r = alpha / (alpha + 1e-3)
    g_u = gamma.sum(axis=0) - N * alpha + (1.0 / C - 1.0) * (r - alpha * r.sum())
    out["grads"] = dict(W=g_W, chi_raw=g_chi, psi=g_psi, beta=g_beta, alpha_unconstr=g_u,
                        loc=g_loc, lsd=g_lsd, gamma_logits=g_t)
    out["R"] = R
    out["dM"] = dM
    return out


# ----------------------------------------------------------------------------------------
# TF1 Adam  (tf.compat.v1.train.AdamOptimizer defaults)
# ----------------------------------------------------------------------------------------
@dataclass
class AdamTF1:
    lr: float = 0.1
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    t: int = 0
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)

    def step(self, params: Params, grads_of_loss: dict):
        """In-place update with g = d(-ELBO)/d theta.  epsilon is NOT bias-corrected (TF1 form)."""
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        for k in PARAM_NAMES:
            g = grads_of_loss[k]
            th = getattr(params, k)
            if th.size == 0:
                continue
            m = self.m.setdefault(k, np.zeros_like(th))
            v = self.v.setdefault(k, np.zeros_like(th))
            m += (1.0 - self.beta1) * (g - m)
            v += (1.0 - self.beta2) * (g * g - v)
            th -= lr_t * m / (np.sqrt(v) + self.epsilon)


# ----------------------------------------------------------------------------------------
# the reference optimisation loop  (R/inference-tflow.R:351-480)
# ----------------------------------------------------------------------------------------
def fit(data, params, eps_source, max_iter=100, rel_tol=1e-5, learning_rate=0.1, n_final=20,
        engine="closed"):
    """Mirror of the session loop.  `eps_source()` returns the next (S,G) draw; it is called
    once per `sess$run` that touches mu_samples, in the reference's order:
    gamma_init, ELBO_0, then (train, eval) per iteration, then `n_final` evals.
    Returns dict(params, elbos, final_elbo, sd_final_elbo, n_iter).
    """
    pre = precompute(data)

    def run(p, eps, want_grads):
        if engine == "closed":
            return elbo_grads_closed(p, data, eps, pre=pre, want_grads=want_grads)
        return elbo_tfgraph(p, data, eps, want_grads=want_grads, want_gamma_init=True)

    params = params.copy()
    gi = run(params, eps_source(), False)["gamma_init"]                    # :368
    params.gamma_logits = gi.copy()                                        # :369
    elbo_val = run(params, eps_source(), False)["elbo"]                    # :372
    if math.isnan(elbo_val):
        raise ValueError("Initial elbo is NA")                             # :374-376
    elbo_diffs = [1e3] * 10                                                # :379
    elbos = [elbo_val]
    adam = AdamTF1(lr=learning_rate)
    n_iter = 0
    for _ in range(max_iter):                                              # :394
        g = run(params, eps_source(), True)["grads"]                       # :401 train
        adam.step(params, {k: -g[k] for k in PARAM_NAMES})
        elbo_new = run(params, eps_source(), False)["elbo"]                # :403 fresh eps
        elbo_diff = (elbo_new - elbo_val) / abs(elbo_val)
        elbo_diffs = elbo_diffs[1:] + [elbo_diff]
        elbos.append(elbo_new)
        elbo_val = elbo_new
        n_iter += 1
        if np.mean(np.abs(elbo_diffs)) < rel_tol:                          # :414
            break
    final = [run(params, eps_source(), False)["elbo"] for _ in range(n_final)]   # :447-449
    t = params.gamma_logits - params.gamma_logits.max(axis=1, keepdims=True)
    gamma = np.exp(t) / np.exp(t).sum(axis=1, keepdims=True)                # :424 softmax(gamma_logits)
    return dict(params=params, elbos=np.array(elbos), final_elbo=float(np.mean(final)) if final else float("nan"),
                sd_final_elbo=float(np.std(final, ddof=1)) if len(final) > 1 else float("nan"),
                n_iter=n_iter, clone_probs=gamma, mu=softplus(params.loc),
                alpha=np.exp(params.alpha_unconstr - np.logaddexp.reduce(params.alpha_unconstr)))


# ----------------------------------------------------------------------------------------
# CPU baseline step (timed by bench.py): TF-graph-shaped, torch float32, autograd + TF1 Adam
# ----------------------------------------------------------------------------------------
def tfgraph_train_step_f32(params, data, eps, adam):
    """One `sess$run(train)` equivalent in float32 on the host cores (the restated reference)."""
    import torch
    g = elbo_tfgraph(params, data, eps, dtype=torch.float32, want_grads=True)["grads"]
    adam.step(params, {k: -g[k] for k in PARAM_NAMES})
    return params
