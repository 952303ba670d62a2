"""ORACLE — test infrastructure only.  CPU emulation of the libdstd_b200 C-ABI contract.

Every method restates, with explicit index contractions and hand-derived
backward formulas, what the corresponding entry point of
``include/dstd_b200.h`` must compute — including the saved tensors ``m``,
``pd``, ``xa`` — so that

  * the Python host layer (``dstd_gcn_b200/ops.py`` + modules) can be tested on a
    CPU-only box by injecting this backend (tests only: the product path has no
    CPU fallback and raises without the CUDA library), and
  * every CUDA entry point can be checked buffer-by-buffer on the GPU.

The formulas themselves are validated against autograd of
``oracle/dstd_oracle.py`` (which is pinned to the reference) in
``tests/test_abi_emul.py``.
"""
from __future__ import annotations

import torch


def _aeff(br):
    a = br["adj"]
    if br.get("adj_w") is not None:
        a = a * br["adj_w"]
    if br.get("adj_r") is not None:
        a = a + br["adj_r"]
    return a


def _like_layout(ref, channels):
    """Empty tensor with ref's dim order but `channels` channels (dim 1)."""
    shape = list(ref.shape)
    shape[1] = channels
    order = sorted(range(ref.dim()), key=lambda d: (ref.stride(d), ref.shape[d]), reverse=True)
    strides = [0] * ref.dim()
    acc = 1
    for d in reversed(order):
        strides[d] = acc
        acc *= shape[d]
    return torch.empty_strided(shape, strides, dtype=ref.dtype, device=ref.device)


class EmulBackend:
    """Same python-level interface as dstd_gcn_b200._lib.CudaBackend."""

    name = "emul"
    launches = 0

    # ------------------------------------------------------------------ DSTD-GC unit
    def gc_forward(self, xu, alpha, brs, skip_u, adj_t):
        n, cin, p_, k_ = xu.shape
        nb = len(brs)
        cout = brs[0]["w_f"].shape[0]
        al = 1.0 if alpha is None else alpha.reshape(())
        m = xu.new_empty((n, nb, 4, p_, k_))
        pd = xu.new_empty((n, nb, p_, k_, k_))
        xa = xu.new_empty((n, nb, cin + 1, p_, k_))
        xaug = torch.cat((xu, xu.new_ones((n, 1, p_, k_))), dim=1)
        out = xu.new_zeros((n, cout, p_, k_))
        for b, br in enumerate(brs):
            wm = torch.cat((br["w_m1"].reshape(2, cin), br["w_m2"].reshape(2, cin)), 0)
            bm = torch.cat((br["b_m1"], br["b_m2"]), 0)
            mb = torch.einsum("jc,ncpk->njpk", wm, xu) + bm.view(1, 4, 1, 1)
            d = torch.tanh(mb[:, 0:2, :, :, None] - mb[:, 2:4, :, None, :])            # [n,r,q,v,w]
            w3 = br["w_rm"].reshape(p_, 2, p_)
            pdb = torch.einsum("prq,nrqvw->npvw", w3, d) + br["b_rm"].view(1, p_, 1, 1)
            xm = al * pdb + _aeff(br).reshape(1, 1, k_, k_)
            xmu = xm.transpose(-1, -2) if adj_t else xm
            xab = torch.einsum("ncpv,npvw->ncpw", xaug, xmu)
            wcat = torch.cat((br["w_f"].reshape(cout, cin), br["b_f"].view(cout, 1)), 1)
            out = out + torch.einsum("oc,ncpw->nopw", wcat, xab)
            m[:, b], pd[:, b], xa[:, b] = mb, pdb, xab
        if skip_u is not None:
            out = out + skip_u
        res = _like_layout(xu, cout)
        res.copy_(out)
        return res, m, pd, xa

    def gc_backward(self, xu, gout_u, alpha, brs, m, pd, xa, adj_t, need_galpha=True, gx_add=None):
        n, cin, p_, k_ = xu.shape
        cout = brs[0]["w_f"].shape[0]
        al = 1.0 if alpha is None else alpha.reshape(())
        xaug = torch.cat((xu, xu.new_ones((n, 1, p_, k_))), dim=1)
        gx = torch.zeros_like(xu)
        if gx_add is not None:
            gx += gx_add
        galpha = xu.new_zeros((1,))
        grads = []
        for b, br in enumerate(brs):
            mb, pdb, xab = m[:, b], pd[:, b], xa[:, b]
            wcat = torch.cat((br["w_f"].reshape(cout, cin), br["b_f"].view(cout, 1)), 1)
            gwcat = torch.einsum("nopw,ncpw->oc", gout_u, xab)
            gxa = torch.einsum("oc,nopw->ncpw", wcat, gout_u)
            xm = al * pdb + _aeff(br).reshape(1, 1, k_, k_)
            xmu = xm.transpose(-1, -2) if adj_t else xm
            gx = gx + torch.einsum("ncpw,npvw->ncpv", gxa[:, :cin], xmu)
            gxmu = torch.einsum("ncpv,ncpw->npvw", xaug, gxa)
            gxm = gxmu.transpose(-1, -2) if adj_t else gxmu
            galpha = galpha + (gxm * pdb).sum().reshape(1)
            gaeff = gxm.sum(dim=(0, 1))
            gp = al * gxm
            d = torch.tanh(mb[:, 0:2, :, :, None] - mb[:, 2:4, :, None, :])
            w3 = br["w_rm"].reshape(p_, 2, p_)
            gwrm = torch.einsum("npvw,nrqvw->prq", gp, d).reshape(p_, 2 * p_)
            gbrm = gp.sum(dim=(0, 2, 3))
            gs = torch.einsum("prq,npvw->nrqvw", w3, gp) * (1 - d * d)
            gm = torch.cat((gs.sum(dim=4), -gs.sum(dim=3)), dim=1)                       # [n,4,p,k]
            wm = torch.cat((br["w_m1"].reshape(2, cin), br["w_m2"].reshape(2, cin)), 0)
            gx = gx + torch.einsum("jc,njpk->ncpk", wm, gm)
            gwm = torch.einsum("njpk,ncpk->jc", gm, xu)
            gbm = gm.sum(dim=(0, 2, 3))
            g = {
                "w_m1": gwm[0:2].clone().reshape(br["w_m1"].shape), "b_m1": gbm[0:2].clone(),
                "w_m2": gwm[2:4].clone().reshape(br["w_m2"].shape), "b_m2": gbm[2:4].clone(),
                "w_rm": gwrm.reshape(br["w_rm"].shape), "b_rm": gbrm,
                "w_f": gwcat[:, :cin].clone().reshape(br["w_f"].shape), "b_f": gwcat[:, cin].clone(),
                "adj_eff": gaeff,
                "adj_w": (br["adj"] * gaeff) if br.get("adj_w") is not None else None,
            }
            grads.append(g)
        gxo = torch.empty_like(xu)
        gxo.copy_(gx)
        return gxo, (galpha if alpha is not None else None), grads

    # ------------------------------------------------------------------ BN + residual + PReLU + mask
    @staticmethod
    def _bn_index(t, c, v, vc_order):
        """[C*V] parameter vector -> [1,C,1,V] broadcastable."""
        return (t.view(v, c).t() if vc_order else t.view(c, v)).reshape(1, c, 1, v)

    @staticmethod
    def _bn_unindex(t, vc_order):
        """[C,V] -> [C*V] in parameter order."""
        return (t.t() if vc_order else t).reshape(-1).clone()

    def bn_act_forward(self, y, r, gamma, beta, running_mean, running_var, nbt, prelu, mask, vc_order, training,
                       eps, momentum, out_like=None):
        n, c, t, v = y.shape
        if training:
            mean = y.mean(dim=(0, 2))
            var = y.var(dim=(0, 2), unbiased=False)
            if running_mean is not None:
                cnt = n * t
                running_mean.mul_(1 - momentum).add_(momentum * self._bn_unindex(mean, vc_order))
                running_var.mul_(1 - momentum).add_(momentum * self._bn_unindex(var, vc_order) * cnt / max(cnt - 1, 1))
            if nbt is not None:
                nbt += 1
        else:
            mean = self._bn_index(running_mean, c, v, vc_order).reshape(c, v)
            var = self._bn_index(running_var, c, v, vc_order).reshape(c, v)
        invstd = torch.rsqrt(var + eps)
        g_, b_ = self._bn_index(gamma, c, v, vc_order), self._bn_index(beta, c, v, vc_order)
        pre = (y - mean.view(1, c, 1, v)) * invstd.view(1, c, 1, v) * g_ + b_
        if r is not None:
            pre = pre + r
        act = torch.where(pre > 0, pre, prelu.reshape(()) * pre) if prelu is not None else pre
        if mask is not None:
            act = act * mask.view(n, c, t, v)
        out = torch.empty_like(out_like if out_like is not None else y, device=y.device)
        out.copy_(act)
        return out, self._bn_unindex(mean, vc_order), self._bn_unindex(invstd, vc_order)

    def bn_act_backward(self, y, r, gout, gamma, beta, prelu, mask, save_mean, save_invstd, vc_order, training,
                        need_gr=True, gr_add=None):
        n, c, t, v = y.shape
        mean = self._bn_index(save_mean, c, v, vc_order)
        invstd = self._bn_index(save_invstd, c, v, vc_order)
        g_, b_ = self._bn_index(gamma, c, v, vc_order), self._bn_index(beta, c, v, vc_order)
        xhat = (y - mean) * invstd
        ga = gout * mask.view(n, c, t, v) if mask is not None else gout
        gprelu = None
        if prelu is not None:
            pre = xhat * g_ + b_
            if r is not None:
                pre = pre + r
            gprelu = (ga * pre * (pre <= 0)).sum().reshape(1)
            gpre = ga * torch.where(pre > 0, torch.ones_like(pre), prelu.reshape(()).expand_as(pre))
        else:
            gpre = ga
        gbeta = gpre.sum(dim=(0, 2))
        ggamma = (gpre * xhat).sum(dim=(0, 2))
        if training:
            cnt = n * t
            gy = g_ * invstd * (gpre - gbeta.view(1, c, 1, v) / cnt - xhat * ggamma.view(1, c, 1, v) / cnt)
        else:
            gy = g_ * invstd * gpre
        gyo = torch.empty_like(y)
        gyo.copy_(gy)
        gro = None
        if r is not None and need_gr:
            gro = torch.empty_like(r)
            gro.copy_(gpre if gr_add is None else gpre + gr_add)
        return gyo, gro, self._bn_unindex(ggamma, vc_order), self._bn_unindex(gbeta, vc_order), gprelu

    # ------------------------------------------------------------------ 1x1 channel mix
    def chmix_forward(self, xu, w, b):
        cout = w.shape[0]
        o = torch.einsum("oc,ncpk->nopk", w.reshape(cout, -1), xu)
        if b is not None:
            o = o + b.view(1, -1, 1, 1)
        res = _like_layout(xu, cout)
        res.copy_(o)
        return res

    def chmix_backward(self, xu, gout, w, need_gx=True):
        cout = w.shape[0]
        gw = torch.einsum("nopk,ncpk->oc", gout, xu).reshape(w.shape)
        gb = gout.sum(dim=(0, 2, 3))
        gx = None
        if need_gx:
            gx = torch.empty_like(xu)
            gx.copy_(torch.einsum("oc,nopk->ncpk", w.reshape(cout, -1), gout))
        return gx, gw, gb

    # ------------------------------------------------------------------ head / tail
    def prep_forward(self, x):
        last = x[:, -1:]
        return torch.cat((x, x - last), dim=-1).permute(0, 3, 1, 2).contiguous()

    def prep_backward(self, gh):
        g = gh.permute(0, 2, 3, 1)            # [N,T,V,6]
        gx = (g[..., :3] + g[..., 3:]).clone()
        gx[:, -1] -= g[..., 3:].sum(dim=1)
        return gx.contiguous()

    def finish_forward(self, z, x):
        return (z.permute(0, 2, 3, 1) + x[:, -1:]).contiguous()

    def finish_backward(self, gy, need_gx=True):
        gz = gy.permute(0, 3, 1, 2).contiguous()
        gx = None
        if need_gx:
            gx = torch.zeros_like(gy)
            gx[:, -1] = gy.sum(dim=1)
        return gz, gx

    # ------------------------------------------------------------------ engine glue
    def mpjpe(self, pred, target, scale, loss_accum, accumulate):
        loss_accum = loss_accum.view(())
        d = pred.reshape(-1, 3) - target.reshape(-1, 3)
        nrm = torch.sqrt((d * d).sum(dim=1))
        j = d.shape[0]
        loss = nrm.mean() * scale
        if accumulate:
            loss_accum += loss
        else:
            loss_accum.copy_(loss.reshape(loss_accum.shape))
        g = torch.where(nrm[:, None] > 0, d / nrm.clamp_min(1e-30)[:, None], torch.zeros_like(d)) * (scale / j)
        return g.reshape(pred.shape)

    def adam_step(self, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, grad_scale, step,
                  lr_dev=None, step_dev=None):
        if lr_dev is not None:
            lr = float(lr_dev)
        if step_dev is not None:
            step = int(step_dev)
        g = grad * grad_scale
        if weight_decay != 0:
            g = g + weight_decay * param
        exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
        param.addcdiv_(exp_avg, denom, value=-lr / bc1)
