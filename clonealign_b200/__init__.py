"""clonealign_b200 — B200-native (sm_100a) backend for clonealign's variational hot path.

Public surface mirrors the reference R package for that path only:
`clonealign`, `run_clonealign`, `inference_tflow`, `clone_assignment`, `saturate`.
Numerics live in libclonealign_b200.so (hand-written CUDA) behind include/clonealign_b200.h.
"""
from .api import CloneAlignFit, clonealign, compute_correlations, run_clonealign
from .inference import clone_assignment, inference_tflow, safe_inverse_softplus, saturate, softplus
from .session import Session

__all__ = ["clonealign", "run_clonealign", "inference_tflow", "clone_assignment", "saturate", "softplus",
           "safe_inverse_softplus", "compute_correlations", "CloneAlignFit", "Session"]
