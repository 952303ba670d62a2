"""Run one DSTD-GC unit (forward + backward through the C ABI) at a BASELINE shape: the smallest command that launches
every hot kernel once, used for `ncu --set full` captures and CUDA-event timing of single kernels on the GPU box.

    python tools/run_unit.py [--n 256] [--mode spatial|temporal|both] [--reps 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dstd_gcn_b200 import _lib  # noqa: E402


def branches(nb, cin, cout, p, k, dev, g):
    out = []
    for _ in range(nb):
        r = lambda *s, sc=0.2: (torch.randn(*s, generator=g) * sc).to(dev)
        out.append(dict(w_m1=r(2, cin, 1, 1), b_m1=r(2), w_m2=r(2, cin, 1, 1), b_m2=r(2), w_rm=r(p, 2 * p, 1, 1),
                        b_rm=r(p), w_f=r(cout, cin, 1, 1), b_f=r(cout),
                        adj=(torch.rand(k, k, generator=g) > 0.7).float().to(dev), adj_w=r(k, k) if nb == 2 else None,
                        adj_r=r(k, k)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--c", type=int, default=64)
    ap.add_argument("--t", type=int, default=35)
    ap.add_argument("--v", type=int, default=22)
    ap.add_argument("--mode", default="both")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    be = _lib.backend()
    g = torch.Generator().manual_seed(0)
    alpha = torch.tensor([0.3], device=dev)
    cases = []
    if a.mode in ("spatial", "both"):
        cases.append(("spatial", a.t, a.v, 2))
    if a.mode in ("temporal", "both"):
        cases.append(("temporal", a.v, a.t, 1))
    for name, p, k, nb in cases:
        x = torch.randn(a.n, a.c, p, k, generator=g).to(dev)
        go = torch.randn(a.n, a.c, p, k, generator=g).to(dev)
        brs = branches(nb, a.c, a.c, p, k, dev, g)
        for rep in range(a.reps):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out, m, pd, xa = be.gc_forward(x, alpha, brs, None, False)
            e[1].record()
            gx, ga, gr = be.gc_backward(x, go, alpha, brs, m, pd, xa if xa.numel() else None, False)
            e[2].record()
            torch.cuda.synchronize()
        print(f"{name:9s} N={a.n} C={a.c} P={p} K={k} nb={nb}: fwd {e[0].elapsed_time(e[1]):.3f} ms  "
              f"bwd {e[1].elapsed_time(e[2]):.3f} ms   (out {float(out.abs().mean()):.4f}, gx {float(gx.abs().mean()):.4f})")


if __name__ == "__main__":
    main()
