"""Cell-sharded multi-GPU plumbing: one process per GPU, `torch.distributed` for rendezvous only.

Cells are the sharding unit: rank r owns a contiguous block of rows of Y together with its
per-cell variational parameters (psi, gamma_logits and their Adam state), which never leave the GPU.
Gene-level quantities are replicated; per train step the library sums G*(2+K+P)+C floats of gene-level
gradient partials with ONE ncclAllReduce on its own stream (clonealign_b200/csrc/core.cu, run_train).
torch.distributed is used for: the ncclUniqueId broadcast, the one-off global column sums at set-up,
barriers and the max-over-ranks of measured times.  On CPU (tests) the same helpers run over gloo.
"""
from __future__ import annotations

import os

import numpy as np


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous, balanced row ranges: the first (n_total % world) ranks get one extra cell."""
    base, extra = divmod(int(n_total), int(world))
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        import datetime
        # short collective timeout: a rank that dies or diverges must take the job down in minutes, not hang the box
        dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=600))
    return rank, local_rank, world


def _dev():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload, nbytes: int, src: int = 0) -> bytes:
    """Ship `nbytes` raw bytes (e.g. the 128-byte ncclUniqueId) from rank `src` to every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return bytes(payload)
    t = torch.zeros(nbytes, dtype=torch.uint8, device=_dev())
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def allreduce_sum(x):
    """Sum a float64 numpy array over ranks (set-up quantities such as colSums(Y), R/inference-tflow.R:117)."""
    import torch
    import torch.distributed as dist
    x = np.asarray(x, dtype=np.float64)
    if not dist.is_initialized():
        return x
    t = torch.from_numpy(x.copy()).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(v: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()


def sharded_session(Y_local, L, psi_local, loc_init, n_total, colsum_local, rank, world, device, **kw):
    """Build this rank's `Session` of a cell-sharded fit (collective: every rank must call it)."""
    from .session import Session
    colsum_total = allreduce_sum(colsum_local) if world > 1 else None
    nccl_id = None
    if world > 1:
        mine = Session.nccl_unique_id() if rank == 0 else bytes(128)
        nccl_id = broadcast_bytes(mine, 128, src=0)
    sess = Session(Y_local, L, psi_local, loc_init, device=device, rank=rank, world=world, nccl_id=nccl_id,
                   n_total=n_total, colsum_total=colsum_total, **kw)
    if world > 1 and "p2p" in _variant_names(kw.get("variants")):
        # variant p2p: gather the CUDA IPC handles of the exchange buffers in rank order and map the peers
        sess.p2p_connect(allgather_bytes(sess.p2p_export(), 64))
    return sess


def _variant_names(v):
    if not v:
        return []
    return [x for x in v.split(",") if x] if isinstance(v, str) else list(v)


def allgather_bytes(payload: bytes, nbytes: int):
    """Every rank contributes `nbytes` raw bytes; returns the list of all contributions in rank order."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return [bytes(payload)]
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(_dev())
    out = [torch.zeros(nbytes, dtype=torch.uint8, device=_dev()) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def gather_concat(x):
    """Concatenate a 1-D numpy array over the ranks in rank order (shards may differ in length); every rank gets the result."""
    import torch
    import torch.distributed as dist
    x = np.ascontiguousarray(x)
    if not dist.is_initialized():
        return x
    world = dist.get_world_size()
    n = torch.tensor([x.size], dtype=torch.int64, device=_dev())
    sizes = [torch.zeros(1, dtype=torch.int64, device=_dev()) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    buf = torch.zeros(cap * x.itemsize, dtype=torch.uint8, device=_dev())
    raw = torch.frombuffer(bytearray(x.tobytes()), dtype=torch.uint8)
    buf[: raw.numel()] = raw.to(_dev())
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [np.frombuffer(o.cpu().numpy().tobytes()[: sizes[r] * x.itemsize], dtype=x.dtype) for r, o in enumerate(out)]
    return np.concatenate(parts)


def shutdown():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
