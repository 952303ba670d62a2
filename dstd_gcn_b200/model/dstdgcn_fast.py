"""`model.dstdgcn_fast` drop-in (reference: /root/reference/model/dstdgcn_fast.py): channels-last [N,T,V,C] variant.

Not the same function as `model.dstdgcn`: the dynamic adjacency is applied
transposed, BatchNorm channels are ordered v*C+c, `conv_f` / the block residual are nn.Linear, A_s is trainable
and there is no W_s / R_s.  Checkpoints are therefore not interchangeable between the two variants.
"""
from ._impl import (_BatchNormBase, _ConvTemporalGraphicalBase, _DSTDGCBase, _DSTDGCBBase, _DSTDGCNBase,
                    _STLayerBase, bn_init, conv_init, weights_init)

__all__ = ["BatchNorm", "DSTDGC", "DSTDGCB", "ConvTemporalGraphical", "ST_GCNN_layer", "DSTDGCN", "conv_init",
           "bn_init", "weights_init"]


class BatchNorm(_BatchNormBase):
    _fast = True


class DSTDGC(_DSTDGCBase):
    _fast = True


class DSTDGCB(_DSTDGCBBase):
    _fast = True
    _gc_cls = DSTDGC
    _bn_cls = BatchNorm


class ConvTemporalGraphical(_ConvTemporalGraphicalBase):
    _fast = True


class ST_GCNN_layer(_STLayerBase):
    _fast = True
    _blk_cls = DSTDGCB
    _ctg_cls = ConvTemporalGraphical


class DSTDGCN(_DSTDGCNBase):
    _fast = True
    _layer_cls = ST_GCNN_layer
    _bn_cls = BatchNorm
