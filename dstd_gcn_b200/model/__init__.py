"""Model registry — the reference's plug-in seam (/root/reference/model/__init__.py:7-14).

``get_model(model_type, **model_opts)`` builds ``model_opts[model_type]`` exactly like the reference.  The
reference selects the fast variant by swapping an import; here both are registered ("dstdgcn" keeps the
reference default, "dstdgcn_fast" is the explicit name of the other one).
"""
from .dstdgcn import DSTDGCN
from .dstdgcn_fast import DSTDGCN as DSTDGCNFast

_REGISTRY = {"dstdgcn": DSTDGCN, "dstdgcn_fast": DSTDGCNFast}


def get_model(model_type, **model_opts):
    return _REGISTRY[model_type](**model_opts[model_type])
