"""ORACLE — test infrastructure only.  Never imported by the product path.

CPU restatement (plain PyTorch, any float dtype, autograd for the gradients) of
the reference's DSTD-GC hot path, written functionally over a ``state_dict`` so
that weights produced by the reference load unchanged:

* ``dstdgc``      <- `DSTDGC.forward`        /root/reference/model/dstdgcn.py:80-94
                     fast twin                /root/reference/model/dstdgcn_fast.py:108-155
* ``batchnorm``   <- `BatchNorm.forward`     model/dstdgcn.py:44-50  (fast: dstdgcn_fast.py:50-56)
* ``dstdgcb``     <- `DSTDGCB.forward`       model/dstdgcn.py:141-163 (fast: dstdgcn_fast.py:237-258)
* ``st_layer``    <- `ST_GCNN_layer.forward` model/dstdgcn.py:234-249
* ``dstdgcn``     <- `DSTDGCN.forward`       model/dstdgcn.py:293-317 (fast: dstdgcn_fast.py:548-614)
* ``mpjpe``       <- `mpjpe_error_3d`        engine/utils/loss.py:52-65
* ``train_step``  <- `PredictionEngine.train` body, engine/prediction.py:215-294

Everything is expressed with explicit index contractions (einsum over named
axes) instead of conv2d/Linear modules; only `torch.nn.functional.batch_norm`
is reused because its running-statistics update rule is the contract.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference
itself (in the build container, where /root/reference exists), runs it on seeded
inputs and commits inputs/weights/outputs/gradients under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against those vectors and against the
known-answer anchors of SURVEY.md Appendix D.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- operator
def _w2(w):
    """Conv2d 1x1 weight [O,I,1,1] or Linear weight [O,I] -> [O,I]."""
    return w.reshape(w.shape[0], w.shape[1])


def dstdgc(x, A, alpha, p, mode="spatial", fast=False):
    """One DSTD-GC.  ``p``: dict with conv_m1/conv_m2/conv_rm/conv_f ``.weight``/``.bias``.

    std : x [N,C,T,V] -> [N,O,T,V]      (model/dstdgcn.py:80-94)
    fast: x [N,T,V,C] -> [N,T,V,O]      (model/dstdgcn_fast.py:108-155)
    ``A`` is [1,K,K] (K=V spatial, K=T temporal); ``alpha`` scalar / 1-element tensor.
    """
    assert mode in {"spatial", "temporal"}
    if fast:
        x = x.permute(0, 3, 1, 2)  # work in [N,C,T,V]; pure index relabelling
    wf, bf = _w2(p["conv_f.weight"]), p["conv_f.bias"]
    w1, b1 = _w2(p["conv_m1.weight"]), p["conv_m1.bias"]
    w2, b2 = _w2(p["conv_m2.weight"]), p["conv_m2.bias"]
    wrm, brm = _w2(p["conv_rm.weight"]), p["conv_rm.bias"]
    xf = torch.einsum("oc,nctv->notv", wf, x) + bf.view(1, -1, 1, 1)
    m1 = torch.einsum("rc,nctv->nrtv", w1, x) + b1.view(1, -1, 1, 1)
    m2 = torch.einsum("rc,nctv->nrtv", w2, x) + b2.view(1, -1, 1, 1)
    n, r, t, v = m1.shape
    if mode == "spatial":
        # k = r*T + t ; D[n,k,v,w] = tanh(m1[n,k,v] - m2[n,k,w])
        d = torch.tanh(m1.reshape(n, r * t, v, 1) - m2.reshape(n, r * t, 1, v))
        xm = (torch.einsum("pk,nkvw->npvw", wrm, d) + brm.view(1, -1, 1, 1)) * alpha + A.reshape(1, 1, v, v)
        if not fast:
            out = torch.einsum("notv,ntvw->notw", xf, xm)
        else:
            out = torch.einsum("ntvw,notw->notv", xm, xf)   # matmul(xm, xf): adjacency used transposed
    else:
        # k = r*V + v ; D[n,k,t,u] = tanh(m1[n,r,t,v] - m2[n,r,u,v])
        m1t = m1.permute(0, 1, 3, 2).reshape(n, r * v, t, 1)
        m2t = m2.permute(0, 1, 3, 2).reshape(n, r * v, 1, t)
        d = torch.tanh(m1t - m2t)
        xm = (torch.einsum("qk,nktu->nqtu", wrm, d) + brm.view(1, -1, 1, 1)) * alpha + A.reshape(1, 1, t, t)
        if not fast:
            out = torch.einsum("notv,nvtu->nouv", xf, xm)
        else:
            out = torch.einsum("nvtu,nouv->notv", xm, xf)
    if fast:
        out = out.permute(0, 2, 3, 1)
    return out


def batchnorm(x, p, prefix, training, fast=False, momentum=0.1, eps=1e-5):
    """BatchNorm1d over C*V channels, statistics over (N,T).

    std : x [N,C,T,V], channel index c*V+v   (model/dstdgcn.py:44-50)
    fast: x [N,T,V,C], channel index v*C+c   (model/dstdgcn_fast.py:50-56)
    In training mode mutates running_mean / running_var / num_batches_tracked in ``p``
    exactly like nn.BatchNorm1d (momentum 0.1, unbiased running variance).
    """
    w, b = p[prefix + "bn.weight"], p[prefix + "bn.bias"]
    rm, rv = p[prefix + "bn.running_mean"], p[prefix + "bn.running_var"]
    if not fast:
        n, c, t, v = x.shape
        h = x.permute(0, 1, 3, 2).reshape(n, c * v, t)
    else:
        n, t, v, c = x.shape
        h = x.permute(0, 2, 3, 1).reshape(n, v * c, t)
    if training and (prefix + "bn.num_batches_tracked") in p:
        p[prefix + "bn.num_batches_tracked"] += 1
    h = F.batch_norm(h, rm, rv, w, b, training, momentum, eps)
    if not fast:
        return h.reshape(n, c, v, t).permute(0, 1, 3, 2)
    return h.reshape(n, v, c, t).permute(0, 3, 1, 2)


def prelu(x, a):
    return torch.where(x > 0, x, a * x)


def _sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def dstdgcb(x, p, prefix, training, fast=False):
    """One DSTD-GC block (model/dstdgcn.py:141-163; fast: dstdgcn_fast.py:237-258)."""
    g = lambda k: p[prefix + k]
    has_res = (prefix + "residual.0.weight") in p
    if has_res:
        wr, br = _w2(g("residual.0.weight")), g("residual.0.bias")
        if not fast:
            r = torch.einsum("oc,nctv->notv", wr, x) + br.view(1, -1, 1, 1)
        else:
            r = torch.einsum("oc,ntvc->ntvo", wr, x) + br
        r = batchnorm(r, p, prefix + "residual.1.", training, fast)
    else:
        r = x
    y = None
    for i in range(g("A_s").shape[0]):
        if not fast:
            a = g("A_s")[i:i + 1] * g("W_s")[i:i + 1] + g("R_s")[i:i + 1]
        else:
            a = g("A_s")[i:i + 1]
        z = dstdgc(x, a, g("alpha_sm"), _sub(p, f"{prefix}conv_s.{i}."), "spatial", fast)
        y = z if y is None else y + z
    x = prelu(batchnorm(y, p, prefix + "bn.", training, fast) + r, g("prelu.weight"))
    y = None
    for i in range(g("A_t").shape[0]):
        a = g("A_t")[i:i + 1] + g("R_t")[i:i + 1]
        z = dstdgc(x, a, g("alpha_tm"), _sub(p, f"{prefix}conv_t.{i}."), "temporal", fast)
        y = z if y is None else y + z
    return y


def st_layer(x, p, prefix, training, residual, fast=False):
    """ST_GCNN_layer with refine=True (model/dstdgcn.py:234-249).  ``residual``: identity skip or none."""
    y = dstdgcb(x, p, prefix + "stgcn.0.0.", training, fast)
    return y + x if residual else y


def dstdgcn(x, p, training=True, fast=False, dropout_mask=None, prefix=""):
    """Whole network, x [N,T,V,3] -> [N,T,V,3] (model/dstdgcn.py:293-317; fast :548-614).

    ``dropout_mask``: optional tensor already scaled by 1/(1-p) in the layout of the
    activation it multiplies ([N,C,T,V] std, [N,T,V,C] fast); None == dropout off.
    """
    n_layers = 0
    while f"{prefix}encoders.{n_layers}.2.weight" in p:
        n_layers += 1
    last = x[:, -1:]
    h = torch.cat((x, x - last), dim=-1)
    if not fast:
        h = h.permute(0, 3, 1, 2)
    h = st_layer(h, p, prefix + "conv_st_in.", training, False, fast)
    h = prelu(batchnorm(h, p, prefix + "bn_in.", training, fast), p[prefix + "prelu.weight"])
    if dropout_mask is not None:
        h = h * dropout_mask
    for i in range(n_layers):
        h = st_layer(h, p, f"{prefix}encoders.{i}.0.", training, True, fast)
        h = prelu(batchnorm(h, p, f"{prefix}encoders.{i}.1.", training, fast), p[f"{prefix}encoders.{i}.2.weight"])
    h = st_layer(h, p, prefix + "conv_st_out.", training, False, fast)
    if not fast:
        h = h.permute(0, 2, 3, 1)
    return h + last


# --------------------------------------------------------------------------- engine glue
def mpjpe(outputs, targets):
    """Mean per-joint position error over every (n,t,v) (engine/utils/loss.py:52-65, unit joint weights)."""
    d = outputs.reshape(-1, 3) - targets.reshape(-1, 3)
    return torch.sqrt((d * d).sum(dim=1)).mean()


def train_loss(p, inputs, inputs_inv, targets, fast=False, inverse=True, dropout_masks=(None, None)):
    """Loss of one engine step on raw [N,T,3V] batches (engine/prediction.py:223-287)."""
    n, t, vc = inputs.shape
    out = dstdgcn(inputs.view(n, t, vc // 3, 3), p, True, fast, dropout_masks[0]).reshape(n, t, vc)
    loss = mpjpe(out, targets)
    if inverse:
        out_i = dstdgcn(inputs_inv.view(n, t, vc // 3, 3), p, True, fast, dropout_masks[1]).reshape(n, t, vc)
        loss = (loss + mpjpe(out_i, torch.flip(targets, dims=[1]))) / 2
    return loss


def state_from_module(module, dtype=None, detach=True):
    """Own-storage copy of a module's state_dict (breaks the reference's A_s/R_s alias);
    floating entries that are parameters get requires_grad as in the module."""
    req = {k: v.requires_grad for k, v in module.named_parameters()}
    out = {}
    for k, v in module.state_dict().items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        if req.get(k, False):
            t.requires_grad_(True)
        out[k] = t
    return out
