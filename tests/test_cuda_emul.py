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
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-DCA_EMULATE", "-I", EMUL, os.path.join(EMUL, "run_interp.cpp"), "-o", out])
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
    Z, dM = body[:N * J].reshape(N, J), body[N * J:(N + G) * J].reshape(G, J)
    Zh, dMh = body[(N + G) * J:(2 * N + G) * J].reshape(N, J), body[(2 * N + G) * J:].reshape(G, J)
    return hdr, Z.astype(np.float64), dM.astype(np.float64), Zh.astype(np.float64), dMh.astype(np.float64)


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
    hdr, Z, dM, Zh, dMh = _run(exe, psi, w, Mx, Rx, smem_panels)
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
    # the monomial tables (Horner: what the per-cell / per-gene kernels of the cell2 set evaluate), the backward one through the
    # ticketed two-level reduction of k_interp_coeffs3: same interpolants as the Chebyshev tables
    assert np.abs(Zh[:, :h] / Z_ref[:, :h] - 1.0).max() < 5e-6
    assert (np.abs(Zh - Z_ref)[:, h:] / zscale).max() < 5e-6
    assert (np.abs(dMh - dM_ref) / np.abs(dM_ref).max(axis=0, keepdims=True).clip(1e-30)).max() < 5e-6
    assert np.abs(Zh - Z).max() <= 2e-6 * np.abs(Z).max() and np.abs(dMh - dM).max() <= 2e-6 * np.abs(dM).max()
    nf_neg, nf_pos, nb = hdr
    assert (nf_neg > 0) == bool((psi < 0).any()) and (nf_pos > 0) == bool((psi >= 0).any())
    if sd_psi == 2.0:
        assert nf_neg + nf_pos > smem_panels          # exercises the coefficients-through-L2 branch of k_interp_eval


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_results_do_not_depend_on_thread_scheduling(order):
    """Race check on the emulation: the threads of a block are visited last-to-first / in a fresh random permutation every
    scheduling round (CA_EMUL_ORDER).  A kernel that is missing a __syncthreads / __syncwarp, or whose reductions are not
    in a fixed order, fails the parity or the bitwise-determinism tests under one of these orders.  The two runs also report
    1 and 148 multiprocessors (the default of the emulation is 3): the persistent kernels and the wave-fitted row blocks of
    the Y pass size their grids and partial-sum buffers from that number."""
    import sys
    env = dict(os.environ, CA_EMUL_ORDER=order, CA_EMUL_SMS="1" if order == "reverse" else "148")
    sel = ("same_seed or c3_column or cell_sharded or several_row" if order == "reverse" else
           "gradients_and_elbo or same_seed or c3_column or cell_sharded")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emul_parity.py"), "-x", "-q", "-k", sel,
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def test_emulator_rejects_invalid_launch_configurations():
    """The emulation enforces the hardware's launch limits (grid / block dimensions, dynamic shared memory above 48 KB
    only after cudaFuncSetAttribute, 227 KB maximum): what the driver would refuse aborts with a message here."""
    src = r'''
#include "cuda_emul.h"
__global__ void k(int* p) { CA_DYNAMIC_SMEM(int, s); if (threadIdx.x == 0) s[0] = 1; p[blockIdx.x] = 1; }
int main(int argc, char** argv) {
  int* p; cudaMalloc(&p, 1024 * sizeof(int));
  int mode = atoi(argv[1]);
  if (mode == 0) { CA_LAUNCH(k, 4, 64, 1024, nullptr)(p); }                                 // fine
  if (mode == 1) { CA_LAUNCH(k, 4, 64, 64 * 1024, nullptr)(p); }                            // > 48 KB without the attribute
  if (mode == 2) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                   CA_LAUNCH(k, 4, 64, 64 * 1024, nullptr)(p); }                            // fine after opting in
  if (mode == 3) { CA_LAUNCH(k, dim3(1, 70000), 64, 0, nullptr)(p); }                       // grid.y > 65535
  if (mode == 4) { CA_LAUNCH(k, 1, 2048, 0, nullptr)(p); }                                  // > 1024 threads
  if (mode == 5) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 300 * 1024) ? 0 : 1; }
  if (mode == 6) { CA_LAUNCH(k, 0, 64, 0, nullptr)(p); }                                    // empty grid
  cudaFree(p);
  puts("ok");
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        cpp, exe_path = os.path.join(td, "t.cpp"), os.path.join(td, "t")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-DCA_EMULATE", "-I", EMUL, cpp, "-o", exe_path])
        for mode, ok in ((0, True), (1, False), (2, True), (3, False), (4, False), (5, True), (6, False)):
            r = subprocess.run([exe_path, str(mode)], capture_output=True, text=True, timeout=120)
            assert (r.returncode == 0) == ok, (mode, r.returncode, r.stderr)
            if not ok:
                assert "invalid launch configuration" in r.stderr, (mode, r.stderr)


def test_parity_subset_under_alignment_and_bounds_sanitizer():
    """A slice of the emulated parity suite on a build with UBSan's alignment / bounds / shift checks
    (CA_EMUL_SANITIZE=1): the emulated float4 / uint4 / double2 carry the device's alignment, so a vector access that
    would raise cudaErrorMisalignedAddress on hardware (and is silently fine on x86) aborts with file:line.  Covers the
    template instantiations of BASELINE config 3 (C = 12, S = 8: 16-byte operand stores), several row / column tiles of the
    Y pass, and every count storage format."""
    env = dict(os.environ, CA_EMUL_SANITIZE="1")
    r = subprocess.run(["python", "-m", "pytest", os.path.join(ROOT, "tests", "test_emul_parity.py"), "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "c3_column_structure or several_row_and_column_tiles or storage_formats_and_input_layouts or batched_y_pass"],
                       env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout and "runtime error" not in r.stderr
