"""Checkpoint format: reference layout (`model.`-prefixed keys, torch-Adam optimizer state), round trips."""
import torch

from dstd_gcn_b200.engine import ModelWrapper, TrainStep, load_checkpoint, save_checkpoint
from dstd_gcn_b200.model import dstdgcn as std
from tests.helpers import load_json


def test_wrapper_prefix_matches_reference_checkpoint_keys():
    keys = load_json("state_keys.json")["std_h36m"]
    m = std.DSTDGCN(*keys["args"])
    got = list(ModelWrapper(m).state_dict().keys())
    assert got == ["model." + k[0] for k in keys["keys"]]


def test_save_load_round_trip(tmp_path):
    torch.manual_seed(0)
    m = std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m")
    step = TrainStep(m, lr=1e-3)
    step.exp_avg.normal_()
    step.exp_avg_sq.uniform_()
    step.step_dev.fill_(7)
    path = str(tmp_path / "last.pth")
    state = save_checkpoint(path, step, err=12.5, epoch=3)
    assert set(state) == {"lr", "err", "model", "optimizer", "scheduler", "epoch"}
    assert all(k.startswith("model.") for k in state["model"])
    # torch's own Adam accepts the optimizer state (what PredictionEngine.recover does, engine/prediction.py:165)
    opt = torch.optim.Adam(std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m").parameters(), lr=3e-3)
    opt.load_state_dict(state["optimizer"])

    m2 = std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m")
    step2 = TrainStep(m2, lr=3e-3)
    epoch, err = load_checkpoint(path, m2, step2)
    assert (epoch, err) == (3, 12.5)
    assert torch.equal(step2.flat.param, step.flat.param)
    assert torch.equal(step2.exp_avg, step.exp_avg) and torch.equal(step2.exp_avg_sq, step.exp_avg_sq)
    assert int(step2.step_dev) == 7 and step2.lr == 1e-3
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    # parameters are still views into the flat bucket after loading
    name, p = step2.flat.named[0]
    assert p.data_ptr() == step2.flat.param.data_ptr()


def test_load_rejects_mismatched_checkpoint():
    m = std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m")
    sd = {"model." + k: v for k, v in m.state_dict().items()}
    sd.pop(next(iter(sd)))
    try:
        load_checkpoint({"model": sd}, m)
    except RuntimeError as e:
        assert "does not match" in str(e)
    else:
        raise AssertionError("mismatch not detected")
