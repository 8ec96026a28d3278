"""Regenerates the committed fixtures under tests/golden/.  Run HERE (needs /root/reference):

    python tests/golden/make_fixtures.py

1. example_sce_counts.npy / example_sce_cn.npy — the bundled `example_sce` (BASELINE config 1) decoded from
   /root/reference/data/example_sce.rda with the pure-Python RDX2 reader (no R in this image); checksums
   from SURVEY.md Appendix C are asserted.
2. golden_c1.npz — float64 oracle outputs (oracle/clonealign_oracle.py) on that fixture with fixed MC draws.
   PARITY UNPINNED: these come from the restatement, not from the reference (no R / TensorFlow here).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.setrecursionlimit(100000)

from rdx2 import read_rda  # noqa: E402
from oracle import clonealign_oracle as O  # noqa: E402


def extract():
    o = read_rda("/root/reference/data/example_sce.rda")
    sce = o["example_sce"]
    m = sce.attr["assays"].attr[".xData"].value[".->data"].attr["listData"].value[0]
    assert m.attr["dim"].value == [100, 200]
    Y = np.array(m.value).reshape(200, 100)          # genes x cells column-major == cells x genes C-order
    ld = sce.attr["rowRanges"].attr["elementMetadata"].attr["listData"]
    names = ld.attr["names"].value
    L = np.stack([np.array(ld.value[names.index(c)].value) for c in "ABC"], 1)
    assert hashlib.sha256(Y.astype("<i4").tobytes()).hexdigest()[:16] == "77e85a201511aa0e"
    assert hashlib.sha256(L.astype("<i4").tobytes()).hexdigest()[:16] == "ba1bb019198ca6af"
    assert Y.sum() == 16090 and (Y != 0).sum() == 5845 and Y.max() == 163
    return Y.astype(np.int32), L.astype(np.int32)


def golden(Y, L):
    hi = O.host_init(Y, L, K=1, rng=None)            # no psi noise: deterministic fixture
    d = O.Data(hi["Y"], hi["L"])
    out = {"psi_init": hi["psi_init"], "mu_guess": hi["mu_guess"]}
    for S in (1, 3):
        rng = np.random.default_rng(1000 + S)
        n_iter = 5
        eps = rng.standard_normal((2 + 2 * n_iter + 3, S, d.Y.shape[1])).astype(np.float32)
        it = iter(eps)
        p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
        r = O.fit(d, p0, lambda: next(it).astype(np.float64), max_iter=n_iter, rel_tol=0.0, n_final=3)
        out[f"eps_S{S}"] = eps
        out[f"elbos_S{S}"] = r["elbos"]
        out[f"final_elbo_S{S}"] = r["final_elbo"]
        out[f"clone_probs_S{S}"] = r["clone_probs"]
        out[f"mu_S{S}"] = r["mu"]
        out[f"W_S{S}"] = r["params"].W
        out[f"psi_S{S}"] = r["params"].psi
        out[f"alpha_S{S}"] = r["alpha"]
    # SURVEY Appendix D self-check values (eps = 0)
    p0 = O.init_params(d.Y, d.L, hi["psi_init"], hi["mu_guess"])
    z = np.zeros((1, d.Y.shape[1]))
    r0 = O.elbo_grads_closed(p0, d, z, want_grads=False)
    p1 = p0.copy()
    p1.gamma_logits = r0["gamma_init"]
    out["elbo_t0"] = r0["elbo"]
    out["elbo_after_init"] = O.elbo_grads_closed(p1, d, z, want_grads=False)["elbo"]
    return out


if __name__ == "__main__":
    Y, L = extract()
    np.save(os.path.join(HERE, "example_sce_counts.npy"), Y)
    np.save(os.path.join(HERE, "example_sce_cn.npy"), L)
    g = golden(Y, L)
    np.savez_compressed(os.path.join(HERE, "golden_c1.npz"), **g)
    print("elbo_t0", g["elbo_t0"], "elbo_after_init", g["elbo_after_init"])
    print("elbos S1", g["elbos_S1"])
