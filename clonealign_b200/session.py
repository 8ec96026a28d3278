"""`Session`: the CUDA replacement for the TensorFlow session used by `inference_tflow()`.

Lifecycle mirrors R/inference-tflow.R:351-457 of the reference:

    sess <- tf$Session(); sess$run(init)          -> Session(...)
    sess$run(gamma_init) / sess$run(init_gamma)   -> Session.init_gamma()
    sess$run(train)                               -> Session.step()
    sess$run(elbo)                                -> Session.elbo()
    sess$run(list(softplus(loc), gamma, ...))     -> Session.params()
    sess$close()                                  -> Session.close()

Everything numeric happens in libclonealign_b200.so (hand-written sm_100a kernels) through the
C-ABI of include/clonealign_b200.h; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _lib

_STORE = {"auto": _lib.STORE_AUTO, "f32": _lib.STORE_F32, "u16": _lib.STORE_U16, "u8": _lib.STORE_U8}
_PATH = {"auto": _lib.PATH_AUTO, "cudacore": _lib.PATH_CUDACORE, "tensor": _lib.PATH_TENSOR, "interp": _lib.PATH_INTERP}
_VARIANT = {"ypass2": _lib.VAR_YPASS2, "epi2": _lib.VAR_EPI2, "lean": _lib.VAR_LEAN, "p2p": _lib.VAR_P2P,
            "overlap": _lib.VAR_OVERLAP, "ypass3": _lib.VAR_YPASS3, "defer": _lib.VAR_DEFER, "ypass4": _lib.VAR_YPASS4,
            "cosched": _lib.VAR_COSCHED, "cell2": _lib.VAR_CELL2, "ypass5": _lib.VAR_YPASS5}


def variant_mask(variants) -> int:
    """Kernel variants (include/clonealign_b200.h, enum ca_variant): iterable / comma-separated string of names."""
    if not variants:
        return 0
    if isinstance(variants, str):
        variants = [v for v in variants.split(",") if v]
    mask = 0
    for v in variants:
        if v not in _VARIANT:
            raise ValueError(f"unknown kernel variant {v!r}; known: {sorted(_VARIANT)}")
        mask |= _VARIANT[v]
    return mask

_SHAPES = {  # name -> lambda(session) -> (rows, cols)
    "W": lambda s: (s.G, s.K), "beta": lambda s: (s.G, s.P), "psi": lambda s: (s.N, s.K),
    "chi_raw": lambda s: (s.K, 1), "alpha_unconstr": lambda s: (s.C, 1), "loc": lambda s: (s.G, 1),
    "lsd": lambda s: (s.G, 1), "gamma_logits": lambda s: (s.N, s.C),
    "Z": lambda s: (s.N, s.S * s.C), "R": lambda s: (s.N, s.S * s.C), "dM": lambda s: (s.G, s.S * s.C),
    "F": lambda s: (s.N, s.C), "YV": lambda s: (s.N, s.K + s.P), "YtU": lambda s: (s.G, s.K + s.P),
    "B": lambda s: (s.N, s.C), "v": lambda s: (s.N, s.C), "s": lambda s: (s.N, 1), "colsum": lambda s: (s.G, 1),
    "shift": lambda s: (s.N, 1), "mu_samples": lambda s: (s.S, s.G),
}


def _f64_colmajor(a, shape=None):
    if a is None:
        return None
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _marshal_y(Y, cfg, keep):
    """Describe the count matrix in `cfg` (dtype / layout / memory space) and return (N, G, pointer)."""
    if hasattr(Y, "data_ptr") and getattr(Y, "is_cuda", False):      # a CUDA tensor already in HBM
        if Y.dim() != 2 or not Y.is_contiguous() or str(Y.dtype) != "torch.float32":
            raise ValueError("device Y must be a contiguous 2-D float32 tensor (cells x genes)")
        N, G = int(Y.shape[0]), int(Y.shape[1])
        yptr = C.c_void_p(Y.data_ptr())
        cfg.y_dtype, cfg.y_layout, cfg.y_mem = _lib.Y_F32, _lib.Y_ROWMAJOR, _lib.Y_DEVICE
        keep.append(Y)
    elif hasattr(Y, "tocsr") and hasattr(Y, "nnz"):                   # scipy.sparse: cells x genes, kept compressed
        Ys = Y.tocsr()
        Ys.sum_duplicates()
        N, G = Ys.shape
        if Ys.nnz >= 2 ** 31:
            raise ValueError("sparse Y with >= 2^31 stored values is not supported (int32 offsets, as R's dgCMatrix)")
        vals = np.ascontiguousarray(Ys.data, dtype=Ys.data.dtype if Ys.data.dtype in (np.float32, np.int32, np.uint8, np.uint16)
                                    else np.float64)
        indptr = np.ascontiguousarray(Ys.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(Ys.indices, dtype=np.int32)
        cfg.y_dtype = {np.dtype(np.float64): _lib.Y_F64, np.dtype(np.float32): _lib.Y_F32, np.dtype(np.int32): _lib.Y_I32,
                       np.dtype(np.uint8): _lib.Y_U8, np.dtype(np.uint16): _lib.Y_U16}[vals.dtype]
        cfg.y_layout, cfg.y_mem = _lib.Y_CSR, _lib.Y_HOST
        cfg.y_indptr, cfg.y_indices = _ptr(indptr), _ptr(indices)
        yptr = _ptr(vals)
        keep += [vals, indptr, indices]
    else:
        Y = np.asarray(Y)
        if Y.ndim != 2:
            raise ValueError("Y must be cells x genes")
        if Y.dtype == np.float64:
            cfg.y_dtype = _lib.Y_F64
        elif Y.dtype == np.float32:
            cfg.y_dtype = _lib.Y_F32
        elif Y.dtype == np.int32:
            cfg.y_dtype = _lib.Y_I32
        elif Y.dtype == np.uint8:
            cfg.y_dtype = _lib.Y_U8
        elif Y.dtype == np.uint16:
            cfg.y_dtype = _lib.Y_U16
        else:
            Y = Y.astype(np.float64)
            cfg.y_dtype = _lib.Y_F64
        if Y.flags.f_contiguous and not Y.flags.c_contiguous:
            cfg.y_layout = _lib.Y_COLMAJOR                            # an R matrix
        else:
            Y = np.ascontiguousarray(Y)
            cfg.y_layout = _lib.Y_ROWMAJOR
        cfg.y_mem = _lib.Y_HOST
        N, G = Y.shape
        yptr = _ptr(Y)
        keep.append(Y)
    return N, G, yptr


def _marshal_data(Y, L, clone_allele, alt, cov, cfg, keep):
    """Y, L and the allele inputs -> (N, G, C, V, pointers...) with `cfg` filled in for them."""
    N, G, yptr = _marshal_y(Y, cfg, keep)
    Lm = _f64_colmajor(L)
    if Lm.ndim != 2 or Lm.shape[0] != G:
        raise ValueError("copy_number_data must have same number of genes (rows) as gene_expression_data")
    Cn = Lm.shape[1]
    V = 0
    ca = al = cv = None
    if clone_allele is not None:
        ca = _f64_colmajor(clone_allele)
        V = ca.shape[0]
        if ca.shape[1] != Cn:
            raise ValueError("clone_allele must be variants x clones")
        al = _f64_colmajor(alt, (N, V))
        cv = _f64_colmajor(cov, (N, V))
    return N, G, Cn, V, yptr, Lm, ca, al, cv


class DeviceData:
    """Inputs of a fit that do not depend on the restart, resident on one GPU: the count matrix as stored plus everything
    derived from it once (ca_core_data_create).  Sessions built with `Session(..., data=this)` share it read-only, so the
    restarts of `run_clonealign` (R/clonealign.R:50-56) upload and preprocess Y once per device."""

    def __init__(self, Y, L, *, device=0, clone_allele=None, alt=None, cov=None, y_store="auto"):
        self._d = None
        lib = _lib.load()
        self._lib = lib
        err = C.create_string_buffer(1024)
        cfg = _lib.CaConfig()
        keep = []
        N, G, Cn, V, yptr, Lm, ca, al, cv = _marshal_data(Y, L, clone_allele, alt, cov, cfg, keep)
        cfg.N, cfg.N_total, cfg.G, cfg.C, cfg.S, cfg.K, cfg.P, cfg.V = N, N, G, Cn, 1, 0, 0, V
        cfg.device, cfg.rank, cfg.world = int(device), 0, 1
        cfg.y_store, cfg.path, cfg.y_ld = _STORE[y_store], _lib.PATH_AUTO, 0
        d = C.c_void_p()
        _lib.check(lib.ca_core_data_create(C.byref(d), C.byref(cfg), yptr, _ptr(Lm), None, _ptr(ca), _ptr(al), _ptr(cv), err,
                                           len(err)), err)
        self._d = d
        self.N, self.G, self.C, self.V, self.device = N, G, Cn, V, int(device)

    def stats(self):
        """rowSums(Y), colSums(Y) and mu_guess = colMeans(Y / rowMeans(Y)) (R/inference-tflow.R:210,117,222) from the
        resident matrix (ca_core_data_stats)."""
        rs, cs, mg = (np.zeros(self.N), np.zeros(self.G), np.zeros(self.G))
        err = C.create_string_buffer(1024)
        _lib.check(self._lib.ca_core_data_stats(self._d, _ptr(rs), _ptr(cs), _ptr(mg), err, len(err)), err)
        return {"rowsum": rs, "colsum": cs, "mu_guess": mg}

    def masked_rowsums(self, gene_keep):
        """rowSums(Y[, keep]) from the resident matrix (ca_core_data_masked_rowsums): the cell filter of
        preprocess_for_clonealign (R/preprocess.R:138-139)."""
        k = np.ascontiguousarray(np.asarray(gene_keep).astype(bool), dtype=np.uint8)
        if k.shape != (self.G,):
            raise ValueError("gene_keep must have one entry per gene")
        rs = np.zeros(self.N)
        err = C.create_string_buffer(1024)
        _lib.check(self._lib.ca_core_data_masked_rowsums(self._d, _ptr(k), _ptr(rs), err, len(err)), err)
        return rs

    def close(self):
        if self._d is not None:
            err = C.create_string_buffer(1024)
            st = self._lib.ca_core_data_destroy(self._d, err, len(err))
            _lib.check(st, err)
            self._d = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ypass_many(sessions):
    """One pass over the shared count matrix for several sessions built on the same `DeviceData` (ca_core_ypass_many)."""
    sessions = list(sessions)
    if not sessions:
        return
    lib = sessions[0]._lib
    arr = (C.c_void_p * len(sessions))(*[s._h for s in sessions])
    err = C.create_string_buffer(1024)
    _lib.check(lib.ca_core_ypass_many(arr, len(sessions), err, len(err)), err)


class Session:
    """One cell shard of one fit on one GPU."""

    def __init__(self, Y, L, psi_init, loc_init, *, mc_samples=1, K=1, x=None, learning_rate=0.1, seed=0,
                 device=0, clone_allele=None, alt=None, cov=None, rank=0, world=1, nccl_id=None, n_total=None,
                 colsum_total=None, y_store="auto", path="auto", variants=None, data=None):
        self._h = None
        if path == "auto":   # operator override, e.g. CLONEALIGN_B200_PATH=cudacore
            path = os.environ.get("CLONEALIGN_B200_PATH", "auto")
        if variants is None:
            variants = os.environ.get("CLONEALIGN_B200_VARIANTS", "")
        lib = _lib.load()
        self._lib = lib
        self._err = C.create_string_buffer(1024)

        cfg = _lib.CaConfig()
        keep = []
        if data is not None:            # inputs already resident (DeviceData): Y / L / allele arguments are ignored
            if data._d is None:
                raise ValueError("DeviceData is closed")
            N, G, Cn, V, device = data.N, data.G, data.C, data.V, data.device
            yptr = Lm = ca = al = cv = None
        else:
            N, G, Cn, V, yptr, Lm, ca, al, cv = _marshal_data(Y, L, clone_allele, alt, cov, cfg, keep)
        K = int(K)
        psi = _f64_colmajor(psi_init, (N, K)) if K > 0 else None
        loc = _f64_colmajor(loc_init, (G,))
        X = None
        P = 0
        if x is not None:
            X = np.asarray(x, dtype=np.float64)
            if X.ndim == 1:
                X = X[:, None]
            if X.shape[0] != N:
                raise ValueError("x must have one row per cell")
            P = X.shape[1]
            X = np.asfortranarray(X)
        cs = _f64_colmajor(colsum_total, (G,)) if colsum_total is not None else None
        idbuf = None
        if world > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("world > 1 needs the 128-byte nccl_id from Session.nccl_unique_id()")
            idbuf = C.create_string_buffer(bytes(nccl_id), 128)

        cfg.N, cfg.N_total = N, int(n_total) if n_total is not None else N
        cfg.G, cfg.C, cfg.S, cfg.K, cfg.P, cfg.V = G, Cn, int(mc_samples), K, P, V
        cfg.learning_rate, cfg.seed = float(learning_rate), int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.device, cfg.rank, cfg.world = int(device), int(rank), int(world)
        cfg.y_store, cfg.path, cfg.y_ld = _STORE[y_store], _PATH[path], 0
        cfg.nccl_id = C.cast(idbuf, C.c_void_p) if idbuf is not None else None
        cfg.variants = variant_mask(variants)

        self.N, self.G, self.C, self.S, self.K, self.P, self.V = N, G, Cn, int(mc_samples), K, P, V
        h = C.c_void_p()
        if data is not None:
            st = lib.ca_core_create_shared(C.byref(h), C.byref(cfg), data._d, _ptr(psi), _ptr(loc), _ptr(X), self._err,
                                           len(self._err))
        else:
            st = lib.ca_core_create(C.byref(h), C.byref(cfg), yptr, _ptr(Lm), _ptr(psi), _ptr(loc), _ptr(X), _ptr(cs),
                                    _ptr(ca), _ptr(al), _ptr(cv), self._err, len(self._err))
        _lib.check(st, self._err)
        self._h = h
        self._data = data               # keep the shared inputs alive as long as this session
        del keep

    # -- discovery --------------------------------------------------------------------------------
    @staticmethod
    def device_count() -> int:
        lib = _lib.load()
        n = C.c_int(0)
        err = C.create_string_buffer(512)
        _lib.check(lib.ca_core_device_count(C.byref(n), err, len(err)), err)
        return n.value

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        err = C.create_string_buffer(512)
        _lib.check(lib.ca_core_nccl_unique_id(buf, err, len(err)), err)
        return buf.raw

    # -- the session operations ---------------------------------------------------------------------
    def _chk(self, st):
        _lib.check(st, self._err)

    def init_gamma(self):
        self._chk(self._lib.ca_core_init_gamma(self._h, self._err, len(self._err)))

    def step(self):
        self._chk(self._lib.ca_core_step(self._h, self._err, len(self._err)))

    def grads(self):
        self._chk(self._lib.ca_core_grads(self._h, self._err, len(self._err)))

    def elbo(self) -> float:
        out = C.c_double(0.0)
        self._chk(self._lib.ca_core_elbo(self._h, C.byref(out), self._err, len(self._err)))
        return out.value

    def elbo_many(self, n: int) -> np.ndarray:
        """n fresh-draw ELBO evaluations with one host round trip (replicate(20, sess$run(elbo)), R/inference-tflow.R:447-449)."""
        out = np.zeros(int(n), dtype=np.float64)
        self._chk(self._lib.ca_core_elbo_many(self._h, int(n), out.ctypes.data_as(C.c_void_p), self._err, len(self._err)))
        return out

    def params(self) -> dict:
        """mu, clone_probs, s, alpha [, beta] [, psi, W, chi] [, clone_probs_from_snv] (R/inference-tflow.R:424-440)."""
        N, G, Cn, K, P = self.N, self.G, self.C, self.K, self.P
        f = lambda *shape: np.zeros(shape, dtype=np.float64, order="F")
        mu, cp, s, alpha = f(G), f(N, Cn), f(N), f(Cn)
        psi = f(N, K) if K > 0 else None
        W = f(G, K) if K > 0 else None
        chi = f(K) if K > 0 else None
        beta = f(G, P) if P > 0 else None
        snv = f(N, Cn) if self.V > 0 else None
        self._chk(self._lib.ca_core_params(self._h, _ptr(mu), _ptr(cp), _ptr(s), _ptr(alpha), _ptr(psi), _ptr(W),
                                           _ptr(chi), _ptr(beta), _ptr(snv), self._err, len(self._err)))
        out = {"mu": mu, "clone_probs": cp, "s": s, "alpha": alpha}
        if P > 0:
            out["beta"] = beta
        if K > 0:
            out.update(psi=psi, W=W, chi=chi)
        if snv is not None:
            out["clone_probs_from_snv"] = snv
        return out

    # -- parity / test hooks ------------------------------------------------------------------------
    def set_eps(self, eps):
        """Queue host-fed N(0,1) draws, shape (n_draws, S, G) or (S, G)."""
        e = np.ascontiguousarray(eps, dtype=np.float32)
        if e.ndim == 2:
            e = e[None]
        if e.shape[1:] != (self.S, self.G):
            raise ValueError(f"eps must have shape (n, {self.S}, {self.G})")
        self._chk(self._lib.ca_core_set_eps(self._h, _ptr(e), e.shape[0], self._err, len(self._err)))

    def get_eps(self):
        e = np.zeros((self.S, self.G), dtype=np.float32)
        self._chk(self._lib.ca_core_get_eps(self._h, _ptr(e), self._err, len(self._err)))
        return e

    def get_array(self, name: str):
        base = name[5:] if name.startswith("grad_") else name
        rows, cols = _SHAPES[base](self)
        out = np.zeros((rows, cols), dtype=np.float64, order="F")
        if out.size:
            self._chk(self._lib.ca_core_get_array(self._h, name.encode(), _ptr(out), out.size, self._err, len(self._err)))
        return np.ascontiguousarray(out)

    def set_array(self, name: str, value):
        rows, cols = _SHAPES[name](self)
        v = _f64_colmajor(value, (rows, cols))
        if v.size:
            self._chk(self._lib.ca_core_set_array(self._h, name.encode(), _ptr(v), v.size, self._err, len(self._err)))

    def correlations(self, clone_idx, L=None):
        """Per-gene Pearson correlation of expression with the assigned clone's copy number, computed on the resident Y
        (compute_correlations, R/clonealign.R:318-334).  clone_idx: N ints, negative = unassigned; L: G x C or None."""
        z = np.ascontiguousarray(clone_idx, dtype=np.int32)
        if z.shape != (self.N,):
            raise ValueError("clone_idx must have one entry per cell")
        Lm = _f64_colmajor(L, (self.G, self.C)) if L is not None else None
        out = np.zeros(self.G, dtype=np.float64)
        self._chk(self._lib.ca_core_correlations(self._h, _ptr(z), _ptr(Lm), _ptr(out), self._err, len(self._err)))
        return out

    def pca_scores(self, max_iter=500, tol=1e-12):
        """Scores of the leading principal component of the centred, scaled log2(Y + 1) (R/inference-tflow.R:203-204)
        by power iteration on the resident Y.  Returns (scores[N], iterations)."""
        out = np.zeros(self.N, dtype=np.float64)
        it = C.c_int32(0)
        self._chk(self._lib.ca_core_pca_scores(self._h, int(max_iter), float(tol), _ptr(out), C.byref(it), self._err,
                                               len(self._err)))
        return out, it.value

    # -- variant p2p: all-reduce over NVLink peer memory (kernels_p2p.cuh) ----------------------------
    def p2p_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange buffer."""
        buf = C.create_string_buffer(64)
        self._chk(self._lib.ca_core_p2p_export(self._h, buf, self._err, len(self._err)))
        return buf.raw

    def p2p_connect(self, handles):
        """`handles`: the world handles from p2p_export() of every rank, in rank order (collective)."""
        blob = b"".join(bytes(h) for h in handles)
        buf = C.create_string_buffer(blob, len(blob))
        self._chk(self._lib.ca_core_p2p_connect(self._h, buf, self._err, len(self._err)))

    # -- measurement hooks --------------------------------------------------------------------------
    def time_steps(self, n_steps: int, with_eval: bool = False) -> float:
        """Milliseconds (CUDA events on the library's stream) for n_steps train steps [+ ELBO evals]."""
        ms = C.c_double(0.0)
        self._chk(self._lib.ca_core_time_steps(self._h, int(n_steps), int(bool(with_eval)), C.byref(ms), self._err,
                                               len(self._err)))
        return ms.value

    def profile_step(self) -> list:
        names = C.create_string_buffer(4096)
        ms = (C.c_double * 64)()
        nk = C.c_int32(0)
        self._chk(self._lib.ca_core_profile_step(self._h, names, len(names), ms, 64, C.byref(nk), self._err,
                                                 len(self._err)))
        labels = names.value.decode().split(";") if names.value else []
        return [(labels[i], ms[i]) for i in range(nk.value)]

    def describe(self) -> dict:
        buf = C.create_string_buffer(1024)
        self._lib.ca_core_describe(self._h, buf, len(buf))
        return json.loads(buf.value.decode())

    def close(self):
        if self._h is not None:
            self._lib.ca_core_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiSession:
    """One fit with its cells sharded over several GPUs of THIS process (ca_core_multi_*): what an R user gets with
    `options(clonealign.gpus = ...)` -- the R interpreter is single-threaded, so the library owns one worker thread per
    device and builds the communicator itself.  Same lifecycle and results as `Session`; `params()` covers all cells.
    Y must be a host array (numpy, any layout / supported dtype) or a scipy.sparse matrix."""

    def __init__(self, Y, L, psi_init, loc_init, *, devices, mc_samples=1, K=1, x=None, learning_rate=0.1, seed=0,
                 clone_allele=None, alt=None, cov=None, y_store="auto", path="auto", variants=None):
        self._m = None
        if path == "auto":
            path = os.environ.get("CLONEALIGN_B200_PATH", "auto")
        if variants is None:
            variants = os.environ.get("CLONEALIGN_B200_VARIANTS", "")
        lib = _lib.load()
        self._lib = lib
        self._err = C.create_string_buffer(1024)
        cfg = _lib.CaConfig()
        keep = []
        N, G, Cn, V, yptr, Lm, ca, al, cv = _marshal_data(Y, L, clone_allele, alt, cov, cfg, keep)
        if cfg.y_mem != _lib.Y_HOST:
            raise ValueError("MultiSession needs Y in host memory")
        K = int(K)
        psi = _f64_colmajor(psi_init, (N, K)) if K > 0 else None
        loc = _f64_colmajor(loc_init, (G,))
        X, P = None, 0
        if x is not None:
            X = np.asarray(x, dtype=np.float64)
            if X.ndim == 1:
                X = X[:, None]
            if X.shape[0] != N:
                raise ValueError("x must have one row per cell")
            P = X.shape[1]
            X = np.asfortranarray(X)
        devs = [int(d) for d in devices]
        if not devs:
            raise ValueError("devices must name at least one GPU")
        cfg.N, cfg.N_total = N, N
        cfg.G, cfg.C, cfg.S, cfg.K, cfg.P, cfg.V = G, Cn, int(mc_samples), K, P, V
        cfg.learning_rate, cfg.seed = float(learning_rate), int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.device, cfg.rank, cfg.world = devs[0], 0, 1
        cfg.y_store, cfg.path, cfg.y_ld = _STORE[y_store], _PATH[path], 0
        cfg.variants = variant_mask(variants)
        darr = (C.c_int32 * len(devs))(*devs)
        m = C.c_void_p()
        _lib.check(lib.ca_core_multi_create(C.byref(m), C.byref(cfg), darr, len(devs), yptr, _ptr(Lm), _ptr(psi), _ptr(loc), _ptr(X),
                                            _ptr(ca), _ptr(al), _ptr(cv), self._err, len(self._err)), self._err)
        self._m = m
        self.N, self.G, self.C, self.S, self.K, self.P, self.V = N, G, Cn, int(mc_samples), K, P, V
        self.devices = devs
        del keep

    def _chk(self, st):
        _lib.check(st, self._err)

    def init_gamma(self):
        self._chk(self._lib.ca_core_multi_init_gamma(self._m, self._err, len(self._err)))

    def step(self):
        self._chk(self._lib.ca_core_multi_step(self._m, self._err, len(self._err)))

    def elbo(self) -> float:
        out = C.c_double(0.0)
        self._chk(self._lib.ca_core_multi_elbo(self._m, C.byref(out), self._err, len(self._err)))
        return out.value

    def elbo_many(self, n: int) -> np.ndarray:
        out = np.zeros(int(n), dtype=np.float64)
        self._chk(self._lib.ca_core_multi_elbo_many(self._m, int(n), out.ctypes.data_as(C.c_void_p), self._err, len(self._err)))
        return out

    def params(self) -> dict:
        N, G, Cn, K, P = self.N, self.G, self.C, self.K, self.P
        f = lambda *shape: np.zeros(shape, dtype=np.float64, order="F")
        mu, cp, s, alpha = f(G), f(N, Cn), f(N), f(Cn)
        psi = f(N, K) if K > 0 else None
        W = f(G, K) if K > 0 else None
        chi = f(K) if K > 0 else None
        beta = f(G, P) if P > 0 else None
        snv = f(N, Cn) if self.V > 0 else None
        self._chk(self._lib.ca_core_multi_params(self._m, _ptr(mu), _ptr(cp), _ptr(s), _ptr(alpha), _ptr(psi), _ptr(W),
                                                 _ptr(chi), _ptr(beta), _ptr(snv), self._err, len(self._err)))
        out = {"mu": mu, "clone_probs": cp, "s": s, "alpha": alpha}
        if P > 0:
            out["beta"] = beta
        if K > 0:
            out.update(psi=psi, W=W, chi=chi)
        if snv is not None:
            out["clone_probs_from_snv"] = snv
        return out

    def time_steps(self, n_steps: int, with_eval: bool = False) -> float:
        ms = C.c_double(0.0)
        self._chk(self._lib.ca_core_multi_time_steps(self._m, int(n_steps), int(bool(with_eval)), C.byref(ms), self._err,
                                                     len(self._err)))
        return ms.value

    def describe(self) -> dict:
        """describe() of shard 0 plus the shard layout."""
        h, a, b = C.c_void_p(), C.c_int64(0), C.c_int64(0)
        rows = []
        d0 = None
        for i in range(len(self.devices)):
            self._lib.ca_core_multi_shard(self._m, i, C.byref(h), C.byref(a), C.byref(b))
            rows.append((a.value, b.value))
            if i == 0:
                buf = C.create_string_buffer(1024)
                self._lib.ca_core_describe(h, buf, len(buf))
                d0 = json.loads(buf.value.decode())
        d0["shards"] = rows
        d0["devices"] = list(self.devices)
        return d0

    def close(self):
        if self._m is not None:
            self._lib.ca_core_multi_destroy(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
