"""Builds clonealign_b200/csrc/core.cu for the HOST against the CUDA-execution-model emulation in this directory
(cuda_emul.h) -> tests/cuda_emul/_build/libclonealign_emul.so, exporting the same C-ABI as the real library.

TEST INFRASTRUCTURE ONLY.  The product never loads this library: clonealign_b200/_lib.py knows one path
(clonealign_b200/libclonealign_b200.so) and fails loudly without it or without a CUDA device.  Tests that want to
exercise the host code + the non-tensor kernels without a GPU use the `emulated_library` fixture of tests/conftest.py,
which points the ctypes loader at the emulated build for the duration of one test module and restores it afterwards.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "clonealign_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")


def _sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".inl"))]
    files += [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith((".h", ".inl"))]
    files.append(os.path.join(ROOT, "include", "clonealign_b200.h"))
    return files


def build(verbose=False) -> str:
    h = hashlib.sha256()
    for f in _sources():
        h.update(f.encode())
        h.update(open(f, "rb").read())
    # CA_EMUL_SANITIZE=1: UBSan alignment / bounds / shift checks.  The emulated float4 / uint4 / double types carry the
    # device's alignment, so a 16-byte vector access at an address that is only 4-byte aligned (cudaErrorMisalignedAddress
    # on hardware, silently fine on x86) aborts with file:line.
    sanitize = os.environ.get("CA_EMUL_SANITIZE", "") not in ("", "0")
    h.update(b"sanitize" if sanitize else b"")
    tag = h.hexdigest()[:16]
    os.makedirs(OUT_DIR, exist_ok=True)
    prefix = "libclonealign_emulsan_" if sanitize else "libclonealign_emul_"
    out = os.path.join(OUT_DIR, f"{prefix}{tag}.so")
    if os.path.exists(out):
        return out
    for old in os.listdir(OUT_DIR):
        if old.startswith(prefix):
            os.unlink(os.path.join(OUT_DIR, old))
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-x", "c++", "-DCA_EMULATE", "-Wno-unknown-pragmas",
           "-I", HERE, "-I", CSRC, os.path.join(CSRC, "core.cu"), os.path.join(CSRC, "multi.cu"), "-o", out + ".tmp", "-ldl", "-pthread"]
    if sanitize:
        cmd[1:1] = ["-fsanitize=alignment,bounds,shift,integer-divide-by-zero,vla-bound,null", "-fno-sanitize-recover=all"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build(verbose=True))
