"""`engine.evaluate` against a plain restatement of PredictionEngine.test (engine/prediction.py:366-406)."""
import numpy as np
import torch

from dstd_gcn_b200 import engine


class _Shift(torch.nn.Module):
    """Stand-in model: [N,T,V,3] -> same shape, a deterministic perturbation (the metric code is what is under test)."""

    def forward(self, x):
        return x * 0.9 + 0.05


def _reference_metric(batches, model, input_n, eval_frame, dim_used, ign, eq):
    t_metric, n_tot, acc_sum, acc_cnt = np.zeros(len(eval_frame)), 0, 0.0, 0
    for inputs, all_seqs in batches:
        n, t, vc = inputs.shape
        out = model(inputs.view(n, t, vc // 3, 3)).reshape(n, t, vc).numpy()
        seqs = all_seqs.numpy()
        pred = seqs.copy()
        pred[:, :, dim_used] = out
        i_ign = np.concatenate((ign * 3, ign * 3 + 1, ign * 3 + 2))
        i_eq = np.concatenate((eq * 3, eq * 3 + 1, eq * 3 + 2))
        pred[:, :, i_ign] = pred[:, :, i_eq]
        p3 = pred.reshape(n, seqs.shape[1], -1, 3)[:, input_n:]
        t3 = seqs.reshape(n, seqs.shape[1], -1, 3)[:, input_n:]
        for k, j in enumerate(eval_frame):
            m = np.mean(np.linalg.norm(t3[:, j].reshape(-1, 3) - p3[:, j].reshape(-1, 3), axis=1)) * n
            t_metric[k] += m
            acc_sum += m
            acc_cnt += n
        n_tot += n
    return acc_sum / acc_cnt, t_metric / n_tot


def test_evaluate_matches_reference_metric():
    g = torch.Generator().manual_seed(5)
    joints_full, t_all, input_n = 32, 35, 10
    used_j = np.array([j for j in range(joints_full) if j not in (0, 1, 6, 11, 16, 20, 23, 24, 28, 31)])   # 22 joints
    dim_used = np.concatenate([used_j * 3 + c for c in range(3)])
    dim_used.sort()
    ign, eq = np.array([16, 20, 23, 24, 28, 31]), np.array([13, 19, 22, 13, 27, 30])
    eval_frame = [1, 3, 7, 9, 13, 24]
    batches = []
    for n in (5, 3):
        all_seqs = torch.randn(n, t_all, joints_full * 3, generator=g)
        batches.append((all_seqs[:, :, dim_used].clone(), all_seqs))
    model = _Shift()
    avg, per_frame = engine.evaluate(model, batches, input_n, eval_frame, dim_used, ign, eq)
    ref_avg, ref_frames = _reference_metric(batches, model, input_n, eval_frame, dim_used, ign, eq)
    assert np.allclose(per_frame, ref_frames, rtol=1e-5, atol=1e-6)
    assert abs(avg - ref_avg) < 1e-5 * max(1.0, abs(ref_avg))


def test_evaluate_tail_prediction_without_dim_used():
    g = torch.Generator().manual_seed(6)
    all_seqs = torch.randn(4, 20, 9, generator=g)

    class Tail(torch.nn.Module):
        def forward(self, x):                 # predicts only the frames after input_n
            return x[:, 8:] + 1.0

    avg, per_frame = engine.evaluate(Tail(), [(all_seqs, all_seqs)], 8, [0, 5], None, None, None)
    assert np.allclose(per_frame, np.sqrt(3.0), atol=1e-5) and abs(avg - np.sqrt(3.0)) < 1e-5
