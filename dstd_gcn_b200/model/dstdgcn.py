"""`model.dstdgcn` drop-in (reference: /root/reference/model/dstdgcn.py): channel-major [N,C,T,V] variant.

Same public names, constructor signatures, forward signatures and state_dict layout as the reference;
the computation runs on the sm_100a kernels (see ``_impl.py``).
"""
from ._impl import (_BatchNormBase, _ConvTemporalGraphicalBase, _DSTDGCBase, _DSTDGCBBase, _DSTDGCNBase,
                    _STLayerBase, bn_init, conv_init, weights_init)

__all__ = ["BatchNorm", "DSTDGC", "DSTDGCB", "ConvTemporalGraphical", "ST_GCNN_layer", "DSTDGCN", "conv_init",
           "bn_init", "weights_init"]


class BatchNorm(_BatchNormBase):
    _fast = False


class DSTDGC(_DSTDGCBase):
    _fast = False


class DSTDGCB(_DSTDGCBBase):
    _fast = False
    _gc_cls = DSTDGC
    _bn_cls = BatchNorm


class ConvTemporalGraphical(_ConvTemporalGraphicalBase):
    _fast = False


class ST_GCNN_layer(_STLayerBase):
    _fast = False
    _blk_cls = DSTDGCB
    _ctg_cls = ConvTemporalGraphical


class DSTDGCN(_DSTDGCNBase):
    _fast = False
    _layer_cls = ST_GCNN_layer
    _bn_cls = BatchNorm
