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
        for rep in range(3):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out, mean, istd = be.bn_act_forward(y, r, gamma, beta, rm, rv, nbt, prelu, None, False, True, 1e-5, 0.1, ol)
            e[1].record()
            gout = out    # any tensor with the output's layout
            res = be.bn_act_backward(y, r, gout, gamma, beta, prelu, None, mean, istd, False, True, True)
            e[2].record()
            torch.cuda.synchronize()
        print(f"{name}: fwd {e[0].elapsed_time(e[1]) * 1e3:.1f} us  bwd {e[1].elapsed_time(e[2]) * 1e3:.1f} us")


if __name__ == "__main__":
    main()
