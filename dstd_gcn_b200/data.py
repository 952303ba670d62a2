"""Batch construction of the reference datasets, done with device-side gathers.

The reference builds three views of every window on the host (dataset/h36m.py:53-63, same in cmu_mocap.py / pw3d.py)
and ships all of them over PCIe.  Here only the raw window ``all_seqs [N, T, 3J]`` has to reach the device; the model
input (observed frames + the last observed frame repeated), the time-reversed input used by the inverse pass and the
target are index gathers on the device."""
from typing import Optional, Sequence, Tuple

import torch


def window_indices(input_n: int, output_n: int, padding: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Frame indices of the forward and the reversed input window (dataset/h36m.py:53-60).

    padding=True : forward  = [0 .. input_n-1] + output_n x [input_n-1]
                   reversed = [output_n+input_n-1 .. output_n] + output_n x [output_n]
    padding=False: forward  = [0 .. T-1], reversed = [T-1 .. 0]"""
    t = input_n + output_n
    if padding:
        fwd = torch.cat((torch.arange(0, input_n), torch.full((output_n,), input_n - 1)))
        inv = torch.cat((torch.arange(output_n + input_n - 1, output_n - 1, -1), torch.full((output_n,), output_n)))
    else:
        fwd = torch.arange(0, t)
        inv = torch.arange(t - 1, -1, -1)
    return fwd.long(), inv.long()


def window_batch(all_seqs: torch.Tensor, input_n: int, output_n: int, dim_used: Optional[Sequence[int]] = None,
                 padding: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(inputs, inputs_inv, targets)`` of ``__getitem__`` (dataset/h36m.py:121-127) for a whole batch, on
    ``all_seqs``'s device.  ``all_seqs``: raw windows ``[N, input_n + output_n, 3J]``; ``dim_used``: coordinate columns
    the model sees (``None`` = all)."""
    assert all_seqs.dim() == 3 and all_seqs.shape[1] == input_n + output_n
    dev = all_seqs.device
    used = all_seqs if dim_used is None else all_seqs.index_select(
        2, torch.as_tensor(dim_used, dtype=torch.long, device=dev))
    fwd, inv = window_indices(input_n, output_n, padding)
    return used.index_select(1, fwd.to(dev)), used.index_select(1, inv.to(dev)), used


# joint groups swapped by the mirror augmentation (dataset/h36m.py:105-112)
_H36M_RIGHT = [1, 2, 3, 4, 5] + [16, 17, 18, 19, 20, 21, 22, 23]
_H36M_LEFT = [6, 7, 8, 9, 10] + [24, 25, 26, 27, 28, 29, 30, 31]


def mirror_h36m(all_seqs: torch.Tensor) -> torch.Tensor:
    """Left/right mirrored copy of raw H3.6M windows ``[N, T, 96]`` (dataset/h36m.py:100-116): swap the limb joints
    and negate x."""
    n, t, vc = all_seqs.shape
    src = all_seqs.view(n, t, vc // 3, 3)
    perm = torch.arange(vc // 3, device=all_seqs.device)
    perm[_H36M_RIGHT] = torch.as_tensor(_H36M_LEFT, device=all_seqs.device)
    perm[_H36M_LEFT] = torch.as_tensor(_H36M_RIGHT, device=all_seqs.device)
    out = src.index_select(2, perm).clone()
    out[..., 0] = -out[..., 0]
    return out.view(n, t, vc)
