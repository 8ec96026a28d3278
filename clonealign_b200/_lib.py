"""ctypes binding of the C-ABI in include/clonealign_b200.h (the same symbols an R `.Call` shim binds).

There is no fallback: if the shared library is missing, or no CUDA device is usable, the calls fail.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclonealign_b200.so")

# enums of include/clonealign_b200.h
Y_F64, Y_F32, Y_I32, Y_U8, Y_U16 = 0, 1, 2, 3, 4
Y_COLMAJOR, Y_ROWMAJOR, Y_CSR = 0, 1, 2
Y_HOST, Y_DEVICE = 0, 1
STORE_AUTO, STORE_F32, STORE_U16, STORE_U8 = 0, 1, 2, 3
PATH_AUTO, PATH_CUDACORE, PATH_TENSOR, PATH_INTERP = 0, 1, 2, 3
VAR_YPASS2, VAR_EPI2, VAR_LEAN, VAR_P2P, VAR_OVERLAP, VAR_YPASS3, VAR_DEFER, VAR_YPASS4, VAR_COSCHED, VAR_CELL2, VAR_YPASS5 = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024
ABI_VERSION = 6

EXPORTS = (
    "ca_core_abi_version", "ca_core_device_count", "ca_core_nccl_unique_id", "ca_core_create",
    "ca_core_destroy", "ca_core_init_gamma", "ca_core_step", "ca_core_elbo", "ca_core_params",
    "ca_core_set_eps", "ca_core_get_eps", "ca_core_grads", "ca_core_get_array", "ca_core_set_array",
    "ca_core_time_steps", "ca_core_profile_step", "ca_core_describe", "ca_core_correlations", "ca_core_pca_scores", "ca_core_p2p_export", "ca_core_p2p_connect", "ca_core_data_create", "ca_core_data_destroy",
    "ca_core_create_shared", "ca_core_ypass_many", "ca_core_data_stats", "ca_core_elbo_many", "ca_core_shutdown",
    "ca_core_multi_create", "ca_core_multi_destroy", "ca_core_multi_init_gamma", "ca_core_multi_step", "ca_core_multi_elbo",
    "ca_core_multi_elbo_many", "ca_core_multi_params", "ca_core_multi_time_steps", "ca_core_multi_shard", "ca_core_multi_size",
    "ca_core_p2p_base", "ca_core_p2p_connect_ptrs", "ca_core_data_masked_rowsums",
)


class CaConfig(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("N_total", C.c_int64),
        ("G", C.c_int32), ("C", C.c_int32), ("S", C.c_int32), ("K", C.c_int32), ("P", C.c_int32), ("V", C.c_int32),
        ("learning_rate", C.c_double), ("seed", C.c_uint64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("y_dtype", C.c_int32), ("y_layout", C.c_int32), ("y_mem", C.c_int32), ("y_store", C.c_int32),
        ("path", C.c_int32), ("y_ld", C.c_int64), ("nccl_id", C.c_void_p), ("variants", C.c_uint32),
        ("y_indptr", C.c_void_p), ("y_indices", C.c_void_p),
    ]


class CloneAlignLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load libclonealign_b200.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CloneAlignLibraryError(
            f"{LIB_PATH} not found: build it with `make -C clonealign_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, dp, cp, sz = C.c_void_p, C.POINTER(C.c_double), C.c_char_p, C.c_size_t
    lib.ca_core_abi_version.restype = C.c_int
    lib.ca_core_abi_version.argtypes = []
    lib.ca_core_device_count.argtypes = [C.POINTER(C.c_int), cp, sz]
    lib.ca_core_nccl_unique_id.argtypes = [vp, cp, sz]
    lib.ca_core_create.argtypes = [C.POINTER(vp), C.POINTER(CaConfig), vp, vp, vp, vp, vp, vp, vp, vp, vp, cp, sz]
    lib.ca_core_destroy.argtypes = [vp]
    lib.ca_core_init_gamma.argtypes = [vp, cp, sz]
    lib.ca_core_step.argtypes = [vp, cp, sz]
    lib.ca_core_grads.argtypes = [vp, cp, sz]
    lib.ca_core_elbo.argtypes = [vp, dp, cp, sz]
    lib.ca_core_elbo_many.argtypes = [vp, C.c_int32, vp, cp, sz]
    lib.ca_core_params.argtypes = [vp] + [vp] * 9 + [cp, sz]
    lib.ca_core_set_eps.argtypes = [vp, vp, C.c_int64, cp, sz]
    lib.ca_core_get_eps.argtypes = [vp, vp, cp, sz]
    lib.ca_core_get_array.argtypes = [vp, cp, vp, C.c_int64, cp, sz]
    lib.ca_core_set_array.argtypes = [vp, cp, vp, C.c_int64, cp, sz]
    lib.ca_core_time_steps.argtypes = [vp, C.c_int32, C.c_int32, dp, cp, sz]
    lib.ca_core_profile_step.argtypes = [vp, cp, sz, dp, C.c_int32, C.POINTER(C.c_int32), cp, sz]
    lib.ca_core_describe.argtypes = [vp, cp, sz]
    lib.ca_core_correlations.argtypes = [vp, vp, vp, vp, cp, sz]
    lib.ca_core_data_create.argtypes = [C.POINTER(vp), C.POINTER(CaConfig), vp, vp, vp, vp, vp, vp, cp, sz]
    lib.ca_core_data_destroy.argtypes = [vp, cp, sz]
    lib.ca_core_create_shared.argtypes = [C.POINTER(vp), C.POINTER(CaConfig), vp, vp, vp, vp, cp, sz]
    lib.ca_core_data_stats.argtypes = [vp, vp, vp, vp, cp, sz]
    lib.ca_core_data_masked_rowsums.argtypes = [vp, vp, vp, cp, sz]
    lib.ca_core_ypass_many.argtypes = [C.POINTER(vp), C.c_int32, cp, sz]
    lib.ca_core_p2p_export.argtypes = [vp, vp, cp, sz]
    lib.ca_core_p2p_connect.argtypes = [vp, vp, cp, sz]
    lib.ca_core_p2p_base.argtypes = [vp, C.POINTER(vp), cp, sz]
    lib.ca_core_p2p_connect_ptrs.argtypes = [vp, vp, vp, cp, sz]
    lib.ca_core_pca_scores.argtypes = [vp, C.c_int32, C.c_double, vp, C.POINTER(C.c_int32), cp, sz]
    lib.ca_core_shutdown.argtypes = []
    lib.ca_core_multi_create.argtypes = [C.POINTER(vp), C.POINTER(CaConfig), vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, cp, sz]
    lib.ca_core_multi_destroy.argtypes = [vp]
    lib.ca_core_multi_init_gamma.argtypes = [vp, cp, sz]
    lib.ca_core_multi_step.argtypes = [vp, cp, sz]
    lib.ca_core_multi_elbo.argtypes = [vp, dp, cp, sz]
    lib.ca_core_multi_elbo_many.argtypes = [vp, C.c_int32, vp, cp, sz]
    lib.ca_core_multi_params.argtypes = [vp] + [vp] * 9 + [cp, sz]
    lib.ca_core_multi_time_steps.argtypes = [vp, C.c_int32, C.c_int32, dp, cp, sz]
    lib.ca_core_multi_shard.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ca_core_multi_size.argtypes = [vp]
    for n in EXPORTS:
        getattr(lib, n).restype = C.c_int
    if lib.ca_core_abi_version() != ABI_VERSION:
        raise CloneAlignLibraryError("libclonealign_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(status: int, errbuf) -> None:
    if status != 0:
        raise CloneAlignLibraryError(errbuf.value.decode("utf-8", "replace") or f"status {status}")
