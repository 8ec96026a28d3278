"""Functional test of the `interp` kernels WITHOUT a GPU: clonealign_b200/csrc/kernels_interp.cuh is compiled for the
host against a minimal CUDA-execution-model emulation (tests/cuda_emul/: one OS thread per CUDA thread, barriers for
__syncthreads / warp shuffles) and run in the same order and with the same arguments as core.cu launches them.
This checks the kernels' index math, panel logic, reductions and Clenshaw evaluation against a direct float64
contraction; GPU-only hazards (memory model, occupancy) are of course not covered."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "cuda_emul")


@pytest.fixture(scope="module")
def exe():
    td = tempfile.mkdtemp()
    out = os.path.join(td, "run_interp")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-I", EMUL, os.path.join(EMUL, "run_interp.cpp"), "-o", out])
    return out


def _run(exe, psi, w, Mx, Rx, smem_panels):
    N, G, J = psi.size, w.size, Mx.shape[1]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            np.array([N, G, J, smem_panels], dtype=np.int32).tofile(f)
            for a in (psi, w, Mx, Rx):
                np.ascontiguousarray(a, dtype=np.float32).tofile(f)
        subprocess.check_call([exe, fin, fout], timeout=600)
        raw = open(fout, "rb").read()
    hdr = np.frombuffer(raw[:12], dtype=np.int32)
    body = np.frombuffer(raw[12:], dtype=np.float32)
    return hdr, body[:N * J].reshape(N, J).astype(np.float64), body[N * J:].reshape(G, J).astype(np.float64)


@pytest.mark.parametrize("sd_psi,sd_w,smem_panels,sign", [(1.0, 0.4, 16, 0), (2.0, 1.5, 2, 0), (1.0, 0.0, 16, 0),
                                                          (1.0, 0.5, 16, +1), (1.0, 0.5, 16, -1)])
def test_interp_kernels_match_direct_contraction(exe, sd_psi, sd_w, smem_panels, sign):
    rng = np.random.default_rng(3)
    N, G, J = 150, 90, 8
    psi = (rng.standard_normal(N) * sd_psi).astype(np.float32)
    if sign:
        psi = (sign * np.abs(psi)).astype(np.float32)
    w = (rng.standard_normal(G) * sd_w).astype(np.float32)
    Mx = rng.uniform(0.1, 5.0, size=(G, J)).astype(np.float32)
    Mx[:, J // 2:] *= w[:, None]
    Rx = rng.uniform(0.0, 1.0, size=(N, J)).astype(np.float32)
    Rx[:, J // 2:] *= psi[:, None]
    hdr, Z, dM = _run(exe, psi, w, Mx, Rx, smem_panels)
    p, ww = psi.astype(np.float64), w.astype(np.float64)
    m = np.maximum(p * ww.max(), p * ww.min())
    E = np.exp(p[:, None] * ww[None, :] - m[:, None])
    Z_ref, dM_ref = E @ Mx.astype(np.float64), E.T @ Rx.astype(np.float64)
    h = J // 2
    # fp32 expf at the nodes + fp32 output: ~1e-6 relative
    assert np.abs(Z[:, :h] / Z_ref[:, :h] - 1.0).max() < 5e-6
    zscale = np.abs(Z_ref[:, :h]).max(axis=1, keepdims=True) * max(np.abs(ww).max(), 1e-30)
    assert (np.abs(Z - Z_ref)[:, h:] / zscale).max() < 5e-6
    assert (np.abs(dM - dM_ref) / np.abs(dM_ref).max(axis=0, keepdims=True).clip(1e-30)).max() < 5e-6
    nf_neg, nf_pos, nb = hdr
    assert (nf_neg > 0) == bool((psi < 0).any()) and (nf_pos > 0) == bool((psi >= 0).any())
    if sd_psi == 2.0:
        assert nf_neg + nf_pos > smem_panels          # exercises the coefficients-through-L2 branch of k_interp_eval


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_results_do_not_depend_on_thread_scheduling(order):
    """Race check on the emulation: the threads of a block are visited last-to-first / in a fresh random permutation every
    scheduling round (CA_EMUL_ORDER).  A kernel that is missing a __syncthreads / __syncwarp, or whose reductions are not
    in a fixed order, fails the parity or the bitwise-determinism tests under one of these orders."""
    import sys
    env = dict(os.environ, CA_EMUL_ORDER=order)
    sel = "gradients_and_elbo or same_seed or c3_column or cell_sharded"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emul_parity.py"), "-x", "-q", "-k", sel,
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
