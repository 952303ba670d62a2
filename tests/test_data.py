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


def _ref_mirror(a, right, left):
    n, t, vc = a.shape
    m = a.copy().reshape(n, t, vc // 3, 3)
    src = a.reshape(n, t, vc // 3, 3)
    m[:, :, right] = src[:, :, left]          # dataset/cmu.py:99-101, dataset/pw3d.py:123-125
    m[:, :, left] = src[:, :, right]
    m[..., 0] = -m[..., 0]
    return m.reshape(n, t, vc)


def test_mirror_cmu_and_3dpw_match_reference():
    rng = np.random.default_rng(7)
    cmu = rng.standard_normal((3, 6, 38 * 3)).astype(np.float32)
    got = data.mirror(torch.from_numpy(cmu), "cmu")
    ref = _ref_mirror(cmu, [2, 3, 4, 5, 6, 21, 22, 23, 24, 27, 25, 26, 28], [8, 9, 10, 11, 12, 30, 31, 32, 33, 36, 24, 35, 37])
    assert np.array_equal(got.numpy(), ref)
    pw = rng.standard_normal((2, 5, 24 * 3)).astype(np.float32)
    got = data.mirror(torch.from_numpy(pw), "3dpw")
    ref = _ref_mirror(pw, [1, 4, 7, 10, 13, 16, 18, 20, 22], [2, 5, 8, 11, 14, 17, 19, 21, 23])
    assert np.array_equal(got.numpy(), ref)
    assert np.array_equal(data.mirror(got, "3dpw").numpy(), pw)     # a clean permutation: an involution


def test_prefetcher_yields_reference_tuples_in_order_on_cpu():
    rng = np.random.default_rng(9)
    raws = [rng.standard_normal((4, 35, 96)).astype(np.float32) for _ in range(5)]
    dim_used = np.arange(0, 66)
    out = list(data.DevicePrefetcher(raws, 10, 25, dim_used, device="cpu", mirror_layout="h36m"))
    assert len(out) == 5
    for raw, (a, b, c, seqs) in zip(raws, out):
        full = np.concatenate((raw, _ref_mirror(raw, data._H36M_RIGHT, data._H36M_LEFT)), axis=0)
        ra, rb, rc = _ref_windows(full, 10, 25, dim_used, True)
        assert np.array_equal(seqs.numpy(), full)
        assert np.array_equal(a.numpy(), ra) and np.array_equal(b.numpy(), rb) and np.array_equal(c.numpy(), rc)
