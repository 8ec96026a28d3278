"""CPU emulation of tensor-operand rounding schemes for the two contractions (decides kernel precision).
Runs the reference loop on the medium synthetic case with the closed-form oracle, but with the contraction
operands rounded as the kernels would, and reports parameter deviation from the exact fp64 loop."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clonealign_oracle as O
from clonealign_b200.synthetic import make_synthetic

def rnd(x, kind):
    t = torch.from_numpy(np.ascontiguousarray(x)).float()
    if kind == "f32": return t.double().numpy()
    if kind == "bf16": return t.bfloat16().double().numpy()
    if kind == "f16": return t.half().double().numpy()
    if kind == "bf16x2":
        hi = t.bfloat16().float(); lo = (t - hi).bfloat16().float(); return (hi + lo).double().numpy()
    if kind == "f16x2":
        hi = t.half().float(); lo = (t - hi).half().float(); return (hi + lo).double().numpy()
    raise ValueError(kind)

def grads(params, data, eps, pre, scheme):
    """closed form with rounded operands: scheme = dict(Ez, Mz, Ezp, Mzp, Eb, Rb)"""
    Y, L, s = data.Y, data.L, data.s
    N, G = Y.shape; C = L.shape[1]; S = eps.shape[0]
    W, psi = params.W, params.psi
    sig_q = np.exp(params.lsd); x = params.loc + sig_q * eps
    mu = O.softplus(x); logmu = np.log(mu); sgm = 1/(1+np.exp(-x))
    eta = psi @ W.T; m = eta.max(1); E = np.exp(eta - m[:, None])
    M = (mu[:, :, None] * L[None])                      # S,G,C
    Mf = M.transpose(1, 0, 2).reshape(G, S*C)
    Z = (rnd(E, scheme["Ez"]) @ rnd(Mf, scheme["Mz"])).reshape(N, S, C).transpose(1, 2, 0)
    Zp = (rnd(E, scheme["Ezp"]) @ rnd(W[:, :1] * Mf, scheme["Mzp"])).reshape(N, S, C).transpose(1, 2, 0)
    logZ = np.log(Z) + m[None, None]
    F = (pre["B"].T[None] - s[None, None] * logZ).mean(0).T
    t = params.gamma_logits; lg = t - np.logaddexp.reduce(t, axis=1, keepdims=True); gamma = np.exp(lg)
    u = params.alpha_unconstr; la = u - np.logaddexp.reduce(u); alpha = np.exp(la); chi = np.exp(params.chi_raw)
    H = F + la[None] - lg
    g_t = gamma * (H - (gamma*H).sum(1, keepdims=True))
    R = gamma.T[None] * s[None, None] / (S * Z)         # S,C,N
    Rf = R.transpose(2, 0, 1).reshape(N, S*C)
    Eb = rnd(E, scheme["Eb"])
    dM = (Eb.T @ rnd(Rf, scheme["Rb"])).reshape(G, S, C).transpose(1, 0, 2)
    dMp = (Eb.T @ rnd(psi[:, :1] * Rf, scheme["Rb"]))   # G, SC
    d_mu = pre["colsum"][None]/(S*mu) - (L[None]*dM).sum(2) - logmu/(S*mu)
    d_x = sgm*d_mu + (1-sgm)/S
    YV = Y @ W; YtU = Y.T @ psi
    g_psi = YV - (Rf * Zp.transpose(2, 0, 1).reshape(N, S*C)).sum(1, keepdims=True) - psi
    g_W = YtU - (Mf * dMp).sum(1, keepdims=True) - chi[None]*W
    r = alpha/(alpha+1e-3)
    return dict(W=g_W, chi_raw=-0.5*chi*(W**2).sum(0)+G/2+1-chi, psi=g_psi, beta=np.zeros((G,0)),
                alpha_unconstr=gamma.sum(0)-N*alpha+(1/C-1)*(r-alpha*r.sum()), loc=d_x.sum(0),
                lsd=(d_x*sig_q[None]*eps).sum(0)+1, gamma_logits=g_t)

def loop(data, p0, eps_list, scheme, n_iter):
    pre = O.precompute(data); p = p0.copy()
    it = iter(eps_list)
    p.gamma_logits = O.elbo_grads_closed(p, data, next(it), pre=pre, want_grads=False)["gamma_init"]; next(it)
    adam = O.AdamTF1(lr=0.1)
    for _ in range(n_iter):
        g = grads(p, data, next(it), pre, scheme); next(it)
        adam.step(p, {k: -g[k] for k in O.PARAM_NAMES})
    return p

if __name__ == "__main__":
    N, G, C, S, n_iter = 2000, 1000, 6, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 6
    syn = make_synthetic(N, G, C, seed=2345234)
    rng = np.random.default_rng(12345)
    hi = O.host_init(syn["Y"], syn["L"], K=1, rng=rng)
    d = O.Data(hi["Y"], hi["L"])
    eps = [rng.standard_normal((S, d.Y.shape[1])) for _ in range(2 + 2*n_iter)]
    p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
    exact = dict(Ez="f32", Mz="f32", Ezp="f32", Mzp="f32", Eb="f32", Rb="f32")
    ref = loop(d, p0, eps, {k: "f32" for k in exact}, n_iter)
    schemes = {
      "current: Z bf16x2; Z',bwd bf16": dict(Ez="bf16x2", Mz="bf16x2", Ezp="bf16", Mzp="bf16", Eb="bf16", Rb="bf16"),
      "A f16, B bf16": dict(Ez="bf16x2", Mz="bf16x2", Ezp="f16", Mzp="bf16", Eb="f16", Rb="bf16"),
      "A f16, B bf16x2": dict(Ez="bf16x2", Mz="bf16x2", Ezp="f16", Mzp="bf16x2", Eb="f16", Rb="bf16x2"),
      "A f16, B f16": dict(Ez="bf16x2", Mz="bf16x2", Ezp="f16", Mzp="f16", Eb="f16", Rb="f16"),
      "A bf16x2, B bf16x2": dict(Ez="bf16x2", Mz="bf16x2", Ezp="bf16x2", Mzp="bf16x2", Eb="bf16x2", Rb="bf16x2"),
      "A f16x2/B bf16x2 for Z too": dict(Ez="f16x2", Mz="bf16x2", Ezp="f16", Mzp="bf16x2", Eb="f16", Rb="bf16x2"),
    }
    rel = lambda a, b: np.abs(a-b).max()/np.abs(b).max()
    for name, sc in schemes.items():
        p = loop(d, p0, eps, sc, n_iter)
        print(f"{name:34s} psi {rel(p.psi, ref.psi):.2e}  W {rel(p.W, ref.W):.2e}  mu {rel(O.softplus(p.loc), O.softplus(ref.loc)):.2e}"
              f"  gamma {np.abs(p.gamma_logits-ref.gamma_logits).max():.2e}")
