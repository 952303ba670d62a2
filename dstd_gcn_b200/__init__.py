"""dstd_gcn_b200 — B200 (sm_100a) implementation of the DSTD-GC hot path of Jaakk0F/DSTD-GCN.

Public surface mirrors the reference's model package:
    from dstd_gcn_b200.model import get_model
    from dstd_gcn_b200.model.dstdgcn import DSTDGCN, DSTDGCB, DSTDGC, BatchNorm, ST_GCNN_layer
    from dstd_gcn_b200.model.dstdgcn_fast import DSTDGCN as DSTDGCNFast
The compute path is `libdstd_b200.so` (C ABI, include/dstd_b200.h) reached through torch.library ops
(`torch.ops.dstd_b200.*`).  There is no CPU fallback.
"""
from . import ops  # registers torch.ops.dstd_b200.*
from .model import get_model

__all__ = ["get_model", "ops"]
__version__ = "0.1.0"
