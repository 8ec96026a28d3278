"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls without a
GPU), the ctypes mirror of ca_config matches the C layout, the host-side mirror logic, the synthetic generator,
and the world_size-2 (gloo) check of the sharding algebra."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clonealign_b200.h")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "clonealign_b200", "libclonealign_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "clonealign_b200", "csrc")])
    from clonealign_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from clonealign_b200 import _lib
    declared = re.findall(r"CA_API\s+int\s+(ca_core_\w+)\s*\(", open(HEADER).read())
    assert len(declared) >= 15
    assert sorted(declared) == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ca_core_abi_version() == _lib.ABI_VERSION


def test_ca_config_layout_matches_c():
    """Compile a tiny C program against the header and compare sizeof / offsetof with the ctypes mirror."""
    from clonealign_b200._lib import CaConfig
    fields = [f[0] for f in CaConfig._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "clonealign_b200.h"\nint main(){printf("%zu", sizeof(ca_config));' + \
        "".join(f'printf(" %zu", offsetof(ca_config, {f}));' for f in fields) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = [int(x) for x in subprocess.check_output([exe]).split()]
    assert out[0] == ctypes.sizeof(CaConfig)
    assert out[1:] == [getattr(CaConfig, f).offset for f in fields]


def _build_c_probe(td):
    exe = os.path.join(td, "c_abi_probe")
    libdir = os.path.join(ROOT, "clonealign_b200")
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_probe.c"),
                           "-o", exe, "-L", libdir, "-lclonealign_b200", "-lm", f"-Wl,-rpath,{libdir}"])
    return exe


def test_plain_c_client_of_the_abi(lib):
    """The boundary is a C ABI (plain pointers and sizes): a C program with R-layout inputs links and runs against it.
    Without a GPU it must fail loudly inside ca_core_create; with one it must complete a train step."""
    with tempfile.TemporaryDirectory() as td:
        out = subprocess.run([_build_c_probe(td)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    from clonealign_b200 import _lib as L
    assert f"abi={L.ABI_VERSION}" in out.stdout
    assert ("create_failed" in out.stdout and "msg_len=0" not in out.stdout) or "ok elbo0=" in out.stdout


def test_fails_loudly_without_gpu(lib, example_sce):
    """No CPU fallback: on a box without CUDA the product path must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clonealign_b200 import inference_tflow
    from clonealign_b200._lib import CloneAlignLibraryError
    Y, L = example_sce
    with pytest.raises(CloneAlignLibraryError):
        inference_tflow(Y, L, max_iter=1, verbose=False, seed=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "clonealign_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


def test_host_helpers_match_reference_semantics():
    from clonealign_b200 import clone_assignment, safe_inverse_softplus, saturate, softplus
    x = np.array([0.1, 1.0, 5.0, 30.0])
    np.testing.assert_allclose(softplus(safe_inverse_softplus(x)), x, rtol=1e-12)
    with pytest.raises(ValueError):
        safe_inverse_softplus(np.array([-1.0]))
    assert saturate(np.array([[1, 7], [6, 3]]), 6).tolist() == [[1, 6], [6, 3]]
    assert clone_assignment(np.array([[.96, .04], [.5, .5]]), ["A", "B"]) == ["A", "unassigned"]


def test_pca_init_properties(example_sce):
    from clonealign_b200.inference import pca_init
    Y, _ = example_sce
    Yf = Y[:, Y.sum(0) > 0]
    p = pca_init(Yf, 1, np.random.default_rng(0))
    assert p.shape == (Yf.shape[0], 1)
    assert abs(p.mean()) < 0.02 and abs(p.std(ddof=1) - 1.0) < 0.02     # scale(pcs) + N(0, .05^2)
    from oracle import clonealign_oracle as O
    q = O.host_init(Y, np.ones((Y.shape[1], 2)), K=1, rng=np.random.default_rng(0))["psi_init"]
    assert np.allclose(np.abs(np.corrcoef(p[:, 0], q[:, 0])[0, 1]), 1.0, atol=1e-3)


def test_pca_init_truncated_matches_full():
    """Large inputs use a truncated SVD for the K leading PCs; same components as the full prcomp up to sign."""
    from clonealign_b200.inference import pca_init
    from clonealign_b200.synthetic import make_synthetic
    Y = make_synthetic(400, 300, 3, seed=5)["Y"]
    class NoNoise:                                         # the N(0, .05^2) jitter (:208) is not what is compared
        def normal(self, loc, scale, size):
            return np.zeros(size)
    for K in (1, 2):
        a = pca_init(Y, K, NoNoise(), truncated=False)
        b = pca_init(Y, K, NoNoise(), truncated=True)
        assert a.shape == b.shape == (400, K)
        for k in range(K):
            assert abs(np.corrcoef(a[:, k], b[:, k])[0, 1]) > 0.999


def test_compute_correlations_matches_per_gene_pearson():
    """compute_correlations (R/clonealign.R:318-334) against an explicit per-gene loop."""
    from clonealign_b200.api import compute_correlations
    rng = np.random.default_rng(0)
    N, G = 60, 25
    Y = rng.poisson(3.0, size=(N, G)).astype(float)
    Y[:, 4] = 2.0                                         # constant expression -> NA
    L = rng.integers(1, 4, size=(G, 3)).astype(float)
    L[7] = 2.0                                            # same copy number in every clone -> NA
    names = ["A", "B", "C"]
    clones = [names[i] for i in rng.integers(0, 3, size=N)]
    clones[3] = clones[10] = "unassigned"
    got = compute_correlations(Y, L, clones, names)
    keep = np.array([c != "unassigned" for c in clones])
    idx = np.array([names.index(c) for c in np.array(clones)[keep]])
    for g in range(G):
        x, y = L[g, idx], Y[keep, g]
        if x.std() == 0 or y.std() == 0:
            assert np.isnan(got[g])
        else:
            assert abs(got[g] - np.corrcoef(x, y)[0, 1]) < 1e-12


def test_preprocess_matches_vignette(example_sce):
    from clonealign_b200.preprocess import preprocess_for_clonealign
    Y, L = example_sce
    pp = preprocess_for_clonealign(Y, L)
    assert pp["gene_expression_data"].shape == (6, 67) and pp["copy_number_data"].shape == (67, 3)


def test_synthetic_generator():
    from clonealign_b200.synthetic import make_synthetic
    a = make_synthetic(300, 200, 4, seed=1)
    b = make_synthetic(300, 200, 4, seed=1)
    assert a["Y"].shape == (300, 200) and a["L"].shape == (200, 4)
    assert np.array_equal(a["Y"], b["Y"])
    assert a["Y"].min() >= 0 and np.all(a["Y"] == np.round(a["Y"]))
    assert np.all(a["Y"].sum(0) > 0) and np.all(a["Y"].sum(1) > 0)
    assert set(np.unique(a["L"])) <= {1, 2, 3, 4}
    rs = a["Y"].sum(1)
    assert 0.5 < np.corrcoef(rs, a["s"])[0, 1]            # s_n is the library size in the benchmark variant


def test_device_generator_is_shard_invariant():
    """make_synthetic_cuda draws the counts in 1024-row blocks with their own seeded generators: the rows a rank draws are
    exactly the rows of the matrix one device would draw, wherever the shard boundaries fall (run on the torch CPU
    generator here; the bench's 1-vs-2/4/8-GPU parity block relies on it)."""
    from clonealign_b200 import dist as D
    from clonealign_b200.synthetic import make_synthetic_cuda
    N, G, C = 2500, 64, 4
    full = make_synthetic_cuda(N, G, C, seed=5, device="cpu")
    assert tuple(full["Y"].shape) == (N, G) and float(full["Y"].sum(dim=1).min()) > 0
    for world in (2, 3, 8):
        parts = [make_synthetic_cuda(N, G, C, seed=5, device="cpu", rows=D.shard_bounds(N, r, world)) for r in range(world)]
        import torch
        assert torch.equal(torch.cat([p["Y"] for p in parts]), full["Y"])
        assert np.array_equal(np.concatenate([p["z"] for p in parts]), full["z"])
    other = make_synthetic_cuda(N, G, C, seed=6, device="cpu")
    assert not np.array_equal(other["Y"].numpy(), full["Y"].numpy())


def test_run_clonealign_spreads_restarts_over_devices(monkeypatch):
    """run_clonealign (R/clonealign.R:35-75): restarts pinned round-robin to `devices`, run from one host thread per
    GPU, results kept in the serial order and the max-ELBO fit returned.  The fit itself is faked (no GPU here)."""
    import threading
    from clonealign_b200 import api
    seen = []
    lock = threading.Lock()

    def fake_clonealign(gex, cnv, **kw):
        import time
        time.sleep(0.05)          # a real fit takes seconds: keeps the per-device workers alive side by side
        with lock:
            seen.append((kw["device"], kw["initial_shrink"], kw["seed"], threading.get_ident()))
        return api.CloneAlignFit(convergence_info={"final_elbo": -1000.0 + kw["seed"] % 97}, clone=["A", "B"],
                                 correlations=np.array([0.1, 0.2]), ml_params={}, seed=kw["seed"], device=kw["device"])

    monkeypatch.setattr(api, "clonealign", fake_clonealign)
    Y, L = np.ones((2, 3)), np.ones((3, 2))
    a = api.run_clonealign(Y, L, initial_shrinks=(0, 5, 10), n_repeats=2, print_elbos=False, seed=7, devices=[0, 1, 2])
    b = api.run_clonealign(Y, L, initial_shrinks=(0, 5, 10), n_repeats=2, print_elbos=False, seed=7)
    assert len(a["multirun_info"]["elbos"]) == 6
    np.testing.assert_array_equal(a["multirun_info"]["elbos"], b["multirun_info"]["elbos"])   # same seeds, same order
    assert a["convergence_info"]["final_elbo"] == a["multirun_info"]["elbos"].max()
    par = seen[:6]
    assert sorted(d for d, *_ in par) == [0, 0, 1, 1, 2, 2]                                   # round-robin pinning
    assert len({t for *_, t in par}) >= 2                                                       # more than one host thread


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU, oracle port) prints one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_shard_bounds_cover_exactly():
    from clonealign_b200.dist import shard_bounds
    for n, w in [(10, 3), (100000, 8), (7, 8), (64, 2)]:
        segs = [shard_bounds(n, r, w) for r in range(w)]
        assert segs[0][0] == 0 and segs[-1][1] == n
        assert all(segs[i][1] == segs[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in segs]
        assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["CA_ROOT"])
from clonealign_b200 import dist as D
from oracle import clonealign_oracle as O
rank, _, world = D.init_process_group("gloo")
Y = np.load(os.path.join(os.environ["CA_ROOT"], "tests/golden/example_sce_counts.npy")).astype(float)
L = np.load(os.path.join(os.environ["CA_ROOT"], "tests/golden/example_sce_cn.npy")).astype(float)
hi = O.host_init(Y, L, K=1, rng=None)
N, G = hi["Y"].shape
rng = np.random.default_rng(0)
S = 2
eps = rng.standard_normal((S, G))
full = O.Data(hi["Y"], hi["L"])
p = O.init_params(full.Y, full.L, hi["psi_init"], hi["mu_guess"])
p.W = rng.normal(size=p.W.shape) * 0.1
p.gamma_logits = rng.normal(size=p.gamma_logits.shape)
ref = O.elbo_grads_closed(p, full, eps)
a, b = D.shard_bounds(N, rank, world)
# this rank's shard: per-cell parameters are local, gene-level ones replicated
colsum = D.allreduce_sum(hi["Y"][a:b].sum(0))
assert np.array_equal(colsum, hi["Y"].sum(0))
loc = O.Data(hi["Y"][a:b], hi["L"])
pl = O.Params(W=p.W, chi_raw=p.chi_raw, psi=p.psi[a:b], beta=p.beta, alpha_unconstr=p.alpha_unconstr, loc=p.loc,
              lsd=p.lsd, gamma_logits=p.gamma_logits[a:b])
pre = O.precompute(loc)
r = O.elbo_grads_closed(pl, loc, eps, pre=pre)
# rank-local LINEAR parts of the gene-level gradients (what the library puts in its allreduce buffer)
mu = r["mu"]; sg = 1 / (1 + np.exp(-(p.loc + np.exp(p.lsd) * eps)))
dmu_lin = -(hi["L"][None] * r["dM"]).sum(axis=2)
dx_lin = sg * dmu_lin
E = np.exp(loc.Y * 0 + pl.psi @ p.W.T - r["m"][:, None])
Q = np.einsum("sgc,scn->ng", mu[:, :, None] * hi["L"][None], r["R"])
part = np.concatenate([dx_lin.sum(0), (dx_lin * np.exp(p.lsd)[None] * eps).sum(0),
                       (loc.Y - E * Q).T @ pl.psi[:, 0], r["gamma"].sum(0)])
tot = D.allreduce_sum(part)
# replicated terms added after the sum, identically on every rank
d_loc = (colsum[None] - np.log(mu)) / (S * mu)
dx_rep = sg * d_loc + (1 - sg) / S
g_loc = tot[:G] + dx_rep.sum(0)
g_lsd = tot[G:2 * G] + (dx_rep * np.exp(p.lsd)[None] * eps).sum(0) + 1.0
g_W = tot[2 * G:3 * G] - np.exp(p.chi_raw[0]) * p.W[:, 0]
al = np.exp(p.alpha_unconstr - np.logaddexp.reduce(p.alpha_unconstr)); rr = al / (al + 1e-3); C = len(al)
g_u = tot[3 * G:] - N * al + (1.0 / C - 1.0) * (rr - al * rr.sum())
for got, want in [(g_loc, ref["grads"]["loc"]), (g_lsd, ref["grads"]["lsd"]), (g_W, ref["grads"]["W"][:, 0]),
                  (g_u, ref["grads"]["alpha_unconstr"])]:
    assert np.abs(got - want).max() <= 1e-9 * (np.abs(want).max() + 1), np.abs(got - want).max()
# per-cell gradients are purely local
assert np.abs(r["grads"]["psi"] - ref["grads"]["psi"][a:b]).max() < 1e-9
assert np.abs(r["grads"]["gamma_logits"] - ref["grads"]["gamma_logits"][a:b]).max() < 1e-9
# the 128-byte id broadcast used for ncclCommInitRank
payload = bytes(range(128)) if rank == 0 else bytes(128)
assert D.broadcast_bytes(payload, 128) == bytes(range(128))
assert D.max_over_ranks(float(rank)) == world - 1
# the 64-byte CUDA IPC handles of variant p2p are gathered in rank order
got = D.allgather_bytes(bytes([rank + 1]) * 64, 64)
assert got == [bytes([r + 1]) * 64 for r in range(world)]
sys.stdout.write("RANK_OK_%d\n" % rank); sys.stdout.flush()
'''


def test_sharding_algebra_gloo_world2():
    """World size 2 over gloo: gene-level gradient partials summed across cell shards + replicated terms ==
    the unsharded gradient; per-cell gradients need no communication."""
    with tempfile.TemporaryDirectory() as td:
        w = os.path.join(td, "worker.py")
        open(w, "w").write(_GLOO_WORKER)
        env = dict(os.environ, CA_ROOT=ROOT, OMP_NUM_THREADS="2")
        port = 29000 + (os.getpid() % 900)
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", str(port), w],
                             env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "RANK_OK_0" in out.stdout and "RANK_OK_1" in out.stdout


def test_r_shim_compiles_against_stub_headers():
    """r/src/ca_shim.c (the .Call binding a clonealign maintainer adds) is type-checked against include/clonealign_b200.h
    with stub declarations of the R API (tests/r_stub/): there is no R in this image, but argument counts / types of every
    ca_core_* call and the registered arities must stay in step with the header."""
    src = os.path.join(ROOT, "r", "src", "ca_shim.c")
    out = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
                          "-Wno-cast-function-type", "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
                          src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    text = open(src).read()
    for m in re.finditer(r'\{"(ca_\w+)", \(DL_FUNC\)&(\w+), (\d+)\}', text):          # registered arity == definition
        name, fn, n = m.group(1), m.group(2), int(m.group(3))
        sig = re.search(r"SEXP %s\(([^)]*)\)" % fn, text).group(1)
        assert name == fn and sig.count("SEXP") == n, (name, n, sig)


def test_product_has_no_emulation_or_cpu_path(lib):
    """The CPU emulation (tests/cuda_emul/) is test infrastructure: the product package never mentions it, the product
    build never defines CA_EMULATE, and the shipped library contains none of its symbols and links the CUDA runtime."""
    pkg = os.path.join(ROOT, "clonealign_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "cuda_emul" not in src and "CA_EMULATE" not in src and "oracle" not in src.replace("the oracle", ""), fn
    mk = open(os.path.join(pkg, "csrc", "Makefile")).read()
    assert "CA_EMULATE" not in mk and "cuda_emul" not in mk
    so = os.path.join(pkg, "libclonealign_b200.so")
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    assert "ca_emul" not in syms
    allsyms = subprocess.check_output(["nm", "-C", so], text=True)
    assert "ca_emul" not in allsyms and "cudaLaunchKernel" in allsyms       # real launches, no fiber launcher
