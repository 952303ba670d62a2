"""Diagnostic (GPU box): per-parameter gradient error of the CUDA path vs the fp64 oracle at a BASELINE shape, next to
the error of the fp32 oracle (= what the reference's own fp32 arithmetic gives) as the yardstick."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dstd_gcn_b200.model import dstdgcn as std          # noqa: E402
from oracle import dstd_oracle as orc                  # noqa: E402
from bench import perturb                               # noqa: E402


def run(dtype, m, x):
    p = orc.state_from_module(m, dtype)
    xx = x.to(dtype).clone().requires_grad_(True)
    y = orc.dstdgcn(xx, p, True, False)
    y.pow(2).mean().backward()
    return y.detach(), xx.grad, {k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None}


def main():
    torch.manual_seed(777)
    layout = sys.argv[1] if len(sys.argv) > 1 else "h36m"
    v, tin, tout = {"h36m": (22, 10, 25), "cmu": (25, 10, 25), "3dpw": (23, 10, 30)}[layout]
    m = perturb(std.DSTDGCN(6, tin, tout, 0.0, v, 64, 5, layout))
    x = torch.randn(4, tin + tout, v, 3, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)
    y64, gx64, g64 = run(torch.float64, m, x)
    y32, gx32, g32 = run(torch.float32, m, x)
    md = m.to("cuda").train()
    xd = x.float().cuda().requires_grad_(True)
    y = md(xd)
    y.pow(2).mean().backward()
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
    print(f"y   : ours {rel(y, y64):.3e}   oracle-fp32 {rel(y32, y64):.3e}")
    print(f"gx  : ours {rel(xd.grad, gx64):.3e}   oracle-fp32 {rel(gx32, gx64):.3e}")
    gmax = max(float(v.abs().max()) for v in g64.values())
    rows = []
    for k, p in md.named_parameters():
        if p.grad is None:
            continue
        ref = g64[k]
        e_ours = float((p.grad.double().cpu() - ref).abs().max())
        e_32 = float((g32[k].double() - ref).abs().max())
        rows.append((e_ours / (float(ref.abs().max()) + 1e-30), k, e_ours, e_32, float(ref.abs().max())))
    rows.sort(reverse=True)
    print(f"global max |g| = {gmax:.3e}")
    for r in rows[:25]:
        print(f"{r[1]:60s} rel {r[0]:.3e}  abs ours {r[2]:.3e}  abs oracle32 {r[3]:.3e}  |g|max {r[4]:.3e}")


if __name__ == "__main__":
    main()
