#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the unmodified reference modules read-only from /root/reference,
runs them on CPU in float64 on seeded inputs and writes small ``.npz``/``.json``
files.  Nothing here is imported at test time; tests only read the files.

What is pinned:
  adjacency.npz     Graph/Time matrices for every layout / a few T
  state_keys.json   state_dict key order, shapes, dtypes, requires_grad (std+fast, 3 layouts)
  op_*.npz          DSTDGC spatial/temporal, std+fast: inputs, weights, output, all gradients
  block_*.npz       DSTDGCB (Cin!=Cout and Cin==Cout), std+fast, train mode
  model_*.npz       small DSTDGCN std+fast: output, loss, every parameter gradient, eval output,
                    BN running stats after the train pass
  train_*.npz       3 Adam steps of the engine loop (inverse=True) on a small model
  anchors.json      the Appendix-D known answers recomputed with the recipe of SURVEY.md
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
import model.dstdgcn as ref_std  # noqa: E402
import model.dstdgcn_fast as ref_fast  # noqa: E402
from model.layers.graph import Graph  # noqa: E402
from model.layers.time import Time  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def dealias(m):
    for p in m.parameters():
        p.data = p.data.clone()
    return m


def perturb(m, seed):
    """Make every path numerically active: non-zero alpha/W_s/R_t/biases, non-trivial BN affine and PReLU."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in m.named_parameters():
            leaf = k.split(".")[-1]
            if leaf in ("alpha_sm", "alpha_tm"):
                p.fill_(0.3 if leaf == "alpha_sm" else -0.2)
            elif leaf == "W_s":
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            elif leaf == "R_t":
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif leaf == "R_s":
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif leaf == "A_s" and p.requires_grad:       # fast variant: trainable
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif leaf == "bias":
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif leaf == "weight" and p.dim() == 1 and p.numel() > 1:   # BN gamma
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif leaf == "weight" and p.dim() == 1:                      # PReLU slope
                p.fill_(0.1 + 0.3 * float(torch.rand(1, generator=g)))
    return m


def npd(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


# ----------------------------------------------------------------------------- adjacency
def gen_adjacency():
    out = {}
    for lay in ("h36m", "cmu", "3dpw"):
        g = Graph(lay)
        out[f"graph_{lay}_all"] = g.get_all_adjacency()
        out[f"graph_{lay}_full"] = g.get_adjacency()
    for t in (2, 3, 8, 12, 35, 40, 125):
        out[f"time_{t}_all"] = Time(t).get_all_adjacency()
    save("adjacency.npz", **out)


# ----------------------------------------------------------------------------- state_dict layout
def gen_state_keys():
    cfgs = {
        "std_h36m": (ref_std, (6, 10, 25, 0.1, 22, 64, 5, "h36m")),
        "fast_h36m": (ref_fast, (6, 10, 25, 0.1, 22, 64, 5, "h36m")),
        "std_cmu": (ref_std, (6, 10, 25, 0.1, 25, 64, 5, "cmu")),
        "std_3dpw": (ref_std, (6, 10, 30, 0.0, 23, 64, 5, "3dpw")),
        "std_default_layers": (ref_std, (6, 10, 25, 0.1, 22, 16)),
    }
    out = {}
    for name, (mod, args) in cfgs.items():
        m = mod.DSTDGCN(*args)
        req = {k: p.requires_grad for k, p in m.named_parameters()}
        out[name] = {
            "args": list(args),
            "n_trainable": int(sum(p.numel() for p in m.parameters() if p.requires_grad)),
            "n_total": int(sum(p.numel() for p in m.parameters())),
            "keys": [[k, list(v.shape), str(v.dtype).replace("torch.", ""), bool(req.get(k, False)), k in req]
                     for k, v in m.state_dict().items()],
        }
    with open(os.path.join(HERE, "state_keys.json"), "w") as f:
        json.dump(out, f)
    print("state_keys.json", {k: len(v["keys"]) for k, v in out.items()})


# ----------------------------------------------------------------------------- operator
def gen_ops():
    for variant, mod in (("std", ref_std), ("fast", ref_fast)):
        for mode in ("spatial", "temporal"):
            torch.manual_seed(11)
            n, cin, cout, t, v = 3, 5, 7, 6, 4
            ref_c, kpt = (t, v) if mode == "spatial" else (v, t)
            op = mod.DSTDGC(cin, cout, ref_c, kpt, mode=mode).double()
            perturb(op, 5)
            g = torch.Generator().manual_seed(3)
            shape = (n, cin, t, v) if variant == "std" else (n, t, v, cin)
            x = torch.randn(shape, generator=g, dtype=torch.float64, requires_grad=True)
            A = torch.randn((1, kpt, kpt), generator=g, dtype=torch.float64, requires_grad=True)
            alpha = torch.tensor([0.37], dtype=torch.float64, requires_grad=True)
            y = op(x, A, alpha)
            gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
            (y * gy).sum().backward()
            arrs = {"x": x, "A": A, "alpha": alpha, "y": y, "gy": gy, "g_x": x.grad, "g_A": A.grad,
                    "g_alpha": alpha.grad}
            for k, p in op.named_parameters():
                arrs["p." + k] = p
                arrs["g." + k] = p.grad
            save(f"op_{variant}_{mode}.npz", **npd(arrs))


# ----------------------------------------------------------------------------- block
def gen_blocks():
    for variant, mod in (("std", ref_std), ("fast", ref_fast)):
        for tag, cin, cout in (("in", 6, 8), ("mid", 8, 8), ("out", 8, 3)):
            torch.manual_seed(21)
            t, v, n = 12, 22, 3
            blk = dealias(mod.DSTDGCB(cin, cout, t, v, "h36m")).double()
            perturb(blk, 7)
            blk.train()
            g = torch.Generator().manual_seed(8)
            shape = (n, cin, t, v) if variant == "std" else (n, t, v, cin)
            x = torch.randn(shape, generator=g, dtype=torch.float64, requires_grad=True)
            sd0 = {k: v_.clone() for k, v_ in blk.state_dict().items()}
            y = blk(x)
            gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
            (y * gy).sum().backward()
            arrs = {"x": x, "y": y, "gy": gy, "g_x": x.grad}
            for k, p in sd0.items():
                arrs["p." + k] = p
            for k, p in blk.named_parameters():
                if p.grad is not None:
                    arrs["g." + k] = p.grad
            for k, b in blk.state_dict().items():
                if "running" in k or "num_batches" in k:
                    arrs["after." + k] = b
            blk.eval()
            with torch.no_grad():
                arrs["y_eval"] = blk(x.detach())
            save(f"block_{variant}_{tag}.npz", **npd(arrs))


# ----------------------------------------------------------------------------- model
def small_model(mod, layout="h36m", v=22, seed=31):
    torch.manual_seed(seed)
    m = dealias(mod.DSTDGCN(6, 4, 6, 0.0, v, 8, 2, layout))
    return perturb(m, 9)


def gen_models():
    for variant, mod in (("std", ref_std), ("fast", ref_fast)):
        for layout, v in (("h36m", 22), ("cmu", 25), ("3dpw", 23)):
            if variant == "fast" and layout != "h36m":
                continue
            m = small_model(mod, layout, v).double().train()
            g = torch.Generator().manual_seed(12)
            x = torch.randn((3, 10, v, 3), generator=g, dtype=torch.float64, requires_grad=True)
            sd0 = {k: t.clone() for k, t in m.state_dict().items()}
            y = m(x)
            loss = y.pow(2).mean()
            loss.backward()
            arrs = {"x": x, "y": y, "loss": loss, "g_x": x.grad}
            for k, p in sd0.items():
                arrs["p." + k] = p
            for k, p in m.named_parameters():
                if p.grad is not None:
                    arrs["g." + k] = p.grad
            for k, b in m.state_dict().items():
                if "running" in k or "num_batches" in k:
                    arrs["after." + k] = b
            m.eval()
            with torch.no_grad():
                arrs["y_eval"] = m(x.detach())
            save(f"model_{variant}_{layout}.npz", **npd(arrs))


# ----------------------------------------------------------------------------- engine steps
def gen_train():
    """Three steps of the engine loop (engine/prediction.py:215-294) driven through the reference's own
    ModelWrapper + mpjpe_error_3d + Adam(lr=3e-3), inverse=True, dropout 0."""
    from engine.prediction import ModelWrapper
    for variant, mod in (("std", ref_std), ("fast", ref_fast)):
        m = small_model(mod).double().train()
        wrap = ModelWrapper(m, {"joint": ["jl2", 1]}, 1)
        opt = torch.optim.Adam(wrap.parameters(), lr=3e-3, weight_decay=0)
        g = torch.Generator().manual_seed(44)
        n, t, v = 4, 10, 22
        arrs = {}
        for k, p in m.state_dict().items():
            arrs["p." + k] = p.clone()
        losses = []
        for step in range(3):
            inputs = torch.randn((n, t, v * 3), generator=g, dtype=torch.float64)
            inputs_inv = torch.flip(inputs, dims=[1]).contiguous()
            targets = inputs + 0.1 * torch.randn((n, t, v * 3), generator=g, dtype=torch.float64)
            arrs[f"inputs{step}"], arrs[f"inputs_inv{step}"], arrs[f"targets{step}"] = inputs, inputs_inv, targets
            out = wrap(inputs.view(n, t, v, 3), False).view(n, t, v * 3)
            loss = wrap.calc_loss(out, targets, "all", None)
            all_loss = sum(loss.values())
            out_i = wrap(inputs_inv.view(n, t, v, 3), True).view(n, t, v * 3)
            inv_idx = torch.arange(t - 1, -1, -1).long()
            loss_i = wrap.calc_loss(out_i, targets[:, inv_idx].contiguous(), "all", None)
            all_loss = (all_loss + sum(loss_i.values())) / 2
            opt.zero_grad()
            all_loss.backward()
            opt.step()
            losses.append(all_loss.detach())
        arrs["losses"] = torch.stack(losses)
        for k, p in m.state_dict().items():
            arrs["after." + k] = p.clone()
        save(f"train_{variant}.npz", **npd(arrs))


# ----------------------------------------------------------------------------- Appendix D anchors
def gen_anchors():
    out = {}
    for variant, mod in (("std", ref_std), ("fast", ref_fast)):
        torch.manual_seed(777)
        m = dealias(mod.DSTDGCN(6, 10, 25, 0.0, 22, 64, 5, "h36m"))
        with torch.no_grad():
            for k, p in m.named_parameters():
                leaf = k.split(".")[-1]
                if leaf in ("alpha_sm", "alpha_tm"):
                    p.fill_(0.1)
                elif leaf == "W_s":
                    p.fill_(0.05)
                elif leaf == "R_t":
                    p.fill_(0.01)
        m.double().train()
        x = torch.randn(4, 35, 22, 3, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)
        y = m(x)
        loss = y.pow(2).mean()
        loss.backward()
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        out[variant] = {"sum_y": float(y.sum()), "mean_abs_y": float(y.abs().mean()),
                        "y000": y[0, 0, 0].tolist(), "y_last": y[3, 34, 21].tolist(),
                        "loss": float(loss), "grad_l2": float(gn)}
    with open(os.path.join(HERE, "anchors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("anchors.json", out)


if __name__ == "__main__":
    gen_adjacency()
    gen_state_keys()
    gen_ops()
    gen_blocks()
    gen_models()
    gen_train()
    gen_anchors()
