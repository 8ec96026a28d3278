"""K = 1 collapses both contractions to UNIVARIATE functions:
     Z_n,j   = sum_g Mx[g,j] exp(psi_n w_g - m(psi_n)) = F_j(psi_n)      (forward,  j over Z and Z' columns)
     dM_g,j  = sum_n Rx[n,j] exp(psi_n w_g - m_n)       = H_j(w_g)        (backward)
   F_j and H_j are entire functions (sums of exponentials): piecewise Chebyshev interpolation on a few panels is
   spectrally accurate.  This prototype measures the accuracy against the direct fp64 contraction."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def cheb_nodes(P):
    return np.cos(np.pi * (np.arange(P) + 0.5) / P)            # first-kind nodes on [-1, 1]

def cheb_coeffs(fvals):
    """fvals: (P, J) values at the first-kind nodes -> (P, J) Chebyshev coefficients (c_0 halved convention applied)."""
    P = fvals.shape[0]
    k = np.arange(P)[:, None]; p = np.arange(P)[None, :]
    Tm = np.cos(np.pi * k * (p + 0.5) / P)                     # (k, p)
    c = (2.0 / P) * (Tm @ fvals)
    c[0] *= 0.5
    return c

def clenshaw(c, t):
    """c: (P, J), t: (n,) in [-1, 1] -> (n, J)"""
    P = c.shape[0]
    b1 = np.zeros((t.size, c.shape[1])); b2 = np.zeros_like(b1)
    for k in range(P - 1, 0, -1):
        b1, b2 = 2 * t[:, None] * b1 - b2 + c[k][None], b1
    return t[:, None] * b1 - b2 + c[0][None]

def panels(lo, hi, half_range_scale, amax=4.0):
    """split [lo, hi] into equal panels such that scale * width / 2 <= amax"""
    width = hi - lo
    if width <= 0: return np.array([lo, lo + 1e-30])
    n = max(1, int(np.ceil(half_range_scale * width / 2.0 / amax)))
    return np.linspace(lo, hi, n + 1)

def forward_interp(psi, w, Mx, P=24, amax=4.0):
    """returns Zx (N, J) with the row shift m_n = max(psi w_max, psi w_min) already divided out."""
    wmax, wmin = w.max(), w.min(); D = wmax - wmin
    out = np.empty((psi.size, Mx.shape[1])); nodes_used = 0
    for sign in (+1, -1):
        sel = psi >= 0 if sign > 0 else psi < 0
        if not sel.any(): continue
        x = psi[sel]
        lo, hi = (0.0, x.max()) if sign > 0 else (x.min(), 0.0)
        wref = wmax if sign > 0 else wmin
        edges = panels(lo, hi, D, amax)
        res = np.empty((x.size, Mx.shape[1]))
        for a, b in zip(edges[:-1], edges[1:]):
            inp = (x >= a) & (x <= b)
            if not inp.any(): continue
            mid, half = 0.5 * (a + b), 0.5 * (b - a) if b > a else 1.0
            xn = mid + half * cheb_nodes(P)
            fv = np.exp(xn[:, None] * (w[None, :] - wref)) @ Mx         # (P, J) node evaluations: the only contraction over G
            nodes_used += P
            c = cheb_coeffs(fv)
            res[inp] = clenshaw(c, (x[inp] - mid) / half)
        out[sel] = res
    return out, nodes_used

def backward_interp(psi, w, Rx, P=24, amax=4.0):
    """dMx (G, J) = sum_n Rx[n, j] exp(psi_n w_g - m_n)"""
    wmax, wmin = w.max(), w.min()
    m = np.maximum(psi * wmax, psi * wmin)
    A = np.abs(psi).max()
    edges = panels(wmin, wmax, A, amax)
    out = np.empty((w.size, Rx.shape[1])); nodes_used = 0
    for a, b in zip(edges[:-1], edges[1:]):
        ing = (w >= a) & (w <= b)
        if not ing.any(): continue
        mid, half = 0.5 * (a + b), 0.5 * (b - a) if b > a else 1.0
        yn = mid + half * cheb_nodes(P)
        hv = np.exp(yn[:, None] * psi[None, :] - m[None, :]) @ Rx        # (P, J): the only contraction over N
        nodes_used += P
        c = cheb_coeffs(hv)
        out[ing] = clenshaw(c, (w[ing] - mid) / half)
    return out, nodes_used

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    N, G, J = 3000, 2000, 32
    for (sp, sw) in [(1.0, 0.3), (1.5, 1.0), (2.0, 2.0), (1.0, 0.0)]:
        psi = rng.standard_normal(N) * sp; w = rng.standard_normal(G) * sw
        Mx = rng.uniform(0.1, 5.0, size=(G, J)); Mx[:, J // 2:] *= w[:, None]      # Z' columns carry w_g (signed)
        Rx = rng.uniform(0.0, 1.0, size=(N, J)); Rx[:, J // 2:] *= psi[:, None]
        m = np.maximum(psi * w.max(), psi * w.min())
        E = np.exp(psi[:, None] * w[None, :] - m[:, None])
        Z_ref, dM_ref = E @ Mx, E.T @ Rx
        for P in (12, 16, 24, 32):
            Z, nf = forward_interp(psi, w, Mx, P); dM, nb = backward_interp(psi, w, Rx, P)
            ez = np.abs(Z / Z_ref - 1)[:, :J // 2].max()
            ezp = (np.abs(Z - Z_ref)[:, J // 2:] / np.abs(Z_ref[:, :J // 2] * np.abs(w).max() + 1e-300)).max()
            ed = (np.abs(dM - dM_ref) / (np.abs(dM_ref).max(0, keepdims=True) + 1e-300)).max()
            print(f"sd_psi {sp} sd_w {sw} range a={np.abs(psi).max() * (w.max() - w.min()) / 2:5.1f} P={P:2d} nodes fwd {nf:4d} bwd {nb:4d} "
                  f"| Z rel {ez:.1e}  Z' {ezp:.1e}  dM {ed:.1e}")
