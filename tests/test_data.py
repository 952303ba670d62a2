"""`dstd_gcn_b200.data` against a numpy restatement of the reference dataset code (dataset/h36m.py:53-63, :100-116)."""
import numpy as np
import torch

from dstd_gcn_b200 import data


def _ref_windows(all_seqs, input_n, output_n, dim_used, padding):
    if padding:
        i_idx = np.append(np.arange(0, input_n), np.repeat([input_n - 1], output_n))
        i_inv = np.append(np.arange(output_n, output_n + input_n)[::-1], np.repeat([output_n], output_n))
    else:
        i_idx = np.arange(0, input_n + output_n)
        i_inv = i_idx[::-1]
    used = all_seqs[:, :, dim_used]
    return used[:, i_idx, :], used[:, i_inv, :], used


def test_window_batch_matches_reference_indexing():
    rng = np.random.default_rng(3)
    all_seqs = rng.standard_normal((5, 35, 96)).astype(np.float32)
    dim_used = np.setdiff1d(np.arange(96), np.concatenate([np.array([0, 1, 6, 11, 16, 20, 23, 24, 28, 31]) * 3 + c
                                                             for c in range(3)]))
    for padding in (True, False):
        got = data.window_batch(torch.from_numpy(all_seqs), 10, 25, dim_used, padding)
        ref = _ref_windows(all_seqs, 10, 25, dim_used, padding)
        for g, r in zip(got, ref):
            assert g.shape == r.shape and np.array_equal(g.numpy(), r)
    fwd, inv = data.window_indices(10, 25)
    assert fwd.tolist() == list(range(10)) + [9] * 25
    assert inv.tolist() == list(range(34, 24, -1)) + [25] * 25


def test_mirror_h36m_matches_reference():
    rng = np.random.default_rng(4)
    a = rng.standard_normal((3, 7, 96)).astype(np.float32)
    m = a.copy().reshape(3, 7, 32, 3)
    src = a.reshape(3, 7, 32, 3)
    right = [1, 2, 3, 4, 5, 16, 17, 18, 19, 20, 21, 22, 23]
    left = [6, 7, 8, 9, 10, 24, 25, 26, 27, 28, 29, 30, 31]
    m[:, :, right] = src[:, :, left]
    m[:, :, left] = src[:, :, right]
    m[..., 0] = -m[..., 0]
    got = data.mirror_h36m(torch.from_numpy(a))
    assert np.array_equal(got.numpy(), m.reshape(3, 7, 96))
    assert np.array_equal(data.mirror_h36m(got).numpy(), a)      # an involution
