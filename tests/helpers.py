"""Shared helpers for the test-suite (fixtures -> torch, error metrics)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name, dtype=None, device="cpu"):
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        t = torch.from_numpy(z[k])
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        out[k] = t.to(device)
    return out


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def split(z, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in z.items() if k.startswith(prefix)}


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
