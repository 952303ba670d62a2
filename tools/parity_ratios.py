"""Parity-ratio table (GPU box): for each BASELINE shape (C=64, L=5) the error of the sm_100a path against the fp64
oracle, next to the error of the fp32 oracle (= the reference's own fp32 arithmetic on the CPU) against the same truth.
SURVEY.md section 4 sets the yardstick "our error vs fp64 <= 2x the reference's fp32 error vs fp64"; this prints the
measured ratio per tensor and writes profiles/r02_parity_ratios.md.

    python tools/parity_ratios.py [--out profiles/r02_parity_ratios.md]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import perturb                               # noqa: E402
from oracle import dstd_oracle as orc                  # noqa: E402

CONFIGS = [("std", "h36m", 22, 10, 25), ("std", "cmu", 25, 10, 25), ("std", "3dpw", 23, 10, 30), ("fast", "h36m", 22, 10, 25)]


def measure(variant, layout, v, tin, tout, batch=4):
    from dstd_gcn_b200.model import dstdgcn as std
    from dstd_gcn_b200.model import dstdgcn_fast as fastm
    fast = variant == "fast"
    torch.manual_seed(777)
    m = perturb((fastm if fast else std).DSTDGCN(6, tin, tout, 0.0, v, 64, 5, layout))
    x = torch.randn(batch, tin + tout, v, 3, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)

    def run(dtype):
        p = orc.state_from_module(m, dtype)
        xx = x.to(dtype).clone().requires_grad_(True)
        y = orc.dstdgcn(xx, p, True, fast)
        y.pow(2).mean().backward()
        return y.detach(), xx.grad, {k: t.grad for k, t in p.items() if t.requires_grad and t.grad is not None}

    y64, gx64, g64 = run(torch.float64)
    y32, gx32, g32 = run(torch.float32)
    md = m.to("cuda").train()
    xd = x.float().cuda().requires_grad_(True)
    y = md(xd)
    y.pow(2).mean().backward()
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
    gmax = max(float(t.abs().max()) for t in g64.values())
    rows = []
    for k, p in md.named_parameters():
        if p.grad is None:
            continue
        ref = g64[k]
        e = float((p.grad.double().cpu() - ref).abs().max())
        e32 = float((g32[k].double() - ref).abs().max())
        rows.append(dict(name=k, e=e, e32=e32, scale=float(ref.abs().max()), n=ref.numel()))
    num = sum(float(((p.grad.double().cpu() - g64[k]) ** 2).sum()) for k, p in md.named_parameters() if p.grad is not None)
    num32 = sum(float(((g32[k].double() - g64[k]) ** 2).sum()) for k in g64)
    den = sum(float((g64[k] ** 2).sum()) for k in g64)
    return dict(y=(rel(y, y64), rel(y32, y64)), gx=(rel(xd.grad, gx64), rel(gx32, gx64)),
                gall=((num / den) ** 0.5, (num32 / den) ** 0.5), gmax=gmax, rows=rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity_ratios.md"))
    a = ap.parse_args()
    lines = ["# Parity ratios: sm_100a path vs fp64 oracle, next to the fp32 oracle (reference arithmetic) vs fp64",
             "",
             "`python tools/parity_ratios.py` on a B200; C=64, L=5, batch 4, training mode, loss = mean(y^2); weights: seed 777 "
             "+ the Appendix-D perturbation.  `ours` / `ref32` = error against the fp64 oracle; ratio = ours / ref32.",
             "Per-tensor errors are max-abs over the tensor; `floor` = 2e-6 x the largest gradient entry of the model "
             "(tensors whose fp32-oracle error is below it are compared with the floor instead).", ""]
    for cfg in CONFIGS:
        r = measure(*cfg)
        floor = 2e-6 * r["gmax"]
        lines += [f"## {cfg[0]} / {cfg[1]} (V={cfg[2]}, T={cfg[3] + cfg[4]})", "",
                  "| quantity | ours | ref32 | ratio |", "|---|---:|---:|---:|",
                  f"| y (rel L2) | {r['y'][0]:.3e} | {r['y'][1]:.3e} | {r['y'][0] / r['y'][1]:.2f} |",
                  f"| dx (rel L2) | {r['gx'][0]:.3e} | {r['gx'][1]:.3e} | {r['gx'][0] / r['gx'][1]:.2f} |",
                  f"| all parameter gradients (rel L2) | {r['gall'][0]:.3e} | {r['gall'][1]:.3e} | {r['gall'][0] / r['gall'][1]:.2f} |",
                  ""]
        rows = sorted(r["rows"], key=lambda t: -t["e"] / max(t["e32"], floor))
        worst = rows[0]["e"] / max(rows[0]["e32"], floor)
        over2 = sum(1 for t in rows if t["e"] > 2 * max(t["e32"], floor))
        lines += [f"{len(rows)} gradient tensors; worst ratio (vs max(ref32, floor = {floor:.2e})) = {worst:.2f}; "
                  f"{over2} tensors above 2x.  Twelve worst:", "",
                  "| tensor | ours | ref32 | scale | ratio |", "|---|---:|---:|---:|---:|"]
        for t in rows[:12]:
            lines.append(f"| `{t['name']}` | {t['e']:.3e} | {t['e32']:.3e} | {t['scale']:.3e} | {t['e'] / max(t['e32'], floor):.2f} |")
        lines.append("")
        print(f"{cfg}: y {r['y'][0] / r['y'][1]:.2f}  dx {r['gx'][0] / r['gx'][1]:.2f}  gall {r['gall'][0] / r['gall'][1]:.2f}  "
              f"worst tensor {worst:.2f} ({rows[0]['name']})  over2x {over2}/{len(rows)}", flush=True)
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
