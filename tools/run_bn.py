"""Time the fused BN + residual + PReLU kernels (forward and backward) at the model's two layout cases.
    python tools/run_bn.py [--n 256]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dstd_gcn_b200 import _lib          # noqa: E402
from dstd_gcn_b200.ops import _out_like  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--gr-add", action="store_true", help="backward with a gr_add tensor (layer-skip gradient)")
    a = ap.parse_args()
    dev, be = torch.device("cuda"), _lib.backend()
    n, c, t, v = a.n, 64, 35, 22
    g = torch.Generator().manual_seed(0)
    gamma, beta = torch.ones(c * v, device=dev), torch.zeros(c * v, device=dev)
    rm, rv, nbt = torch.zeros(c * v, device=dev), torch.ones(c * v, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    prelu = torch.tensor([0.25], device=dev)
    cases = {
        "block BN  (y T-major, r T-major -> out V-major)": (torch.randn(n, c, t, v, generator=g).to(dev),
                                                            torch.randn(n, c, t, v, generator=g).to(dev), 2),
        "encoder BN (y V-major -> out T-major)": (torch.randn(n, c, v, t, generator=g).to(dev).permute(0, 1, 3, 2), None, 1),
    }
    for name, (y, r, order) in cases.items():
        ol = _out_like(y, order)
        out, mean, istd = be.bn_act_forward(y, r, gamma, beta, rm, rv, nbt, prelu, None, False, True, 1e-5, 0.1, ol)
        gout = out    # any tensor with the output's layout
        gadd = torch.randn_like(gout) if (a.gr_add and r is not None) else None
        be.bn_act_backward(y, r, gout, gamma, beta, prelu, None, mean, istd, False, True, True, gadd)
        torch.cuda.synchronize()
        # device time without host gaps: replay each direction as a CUDA graph
        gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(gf):
            be.bn_act_forward(y, r, gamma, beta, rm, rv, nbt, prelu, None, False, True, 1e-5, 0.1, ol)
        with torch.cuda.graph(gb):
            be.bn_act_backward(y, r, gout, gamma, beta, prelu, None, mean, istd, False, True, True, gadd)
        res = []
        for gr in (gf, gb):
            gr.replay()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            for _ in range(a.reps):
                gr.replay()
            e[1].record()
            torch.cuda.synchronize()
            res.append(e[0].elapsed_time(e[1]) * 1e3 / a.reps)
        print(f"{name}: fwd {res[0]:.1f} us  bwd {res[1]:.1f} us")


if __name__ == "__main__":
    main()
