"""Data-parallel host logic on CPU: 2 ranks over gloo, the ABI emulation standing in for the CUDA library.

Parity definition for DP: K ranks, each running the reference step on its own batch shard with
per-rank BatchNorm statistics, gradients averaged -> compared with `TrainStep` under torch.distributed, where the only
collective is one all_reduce(sum) of the flat gradient bucket and the 1/world scale is folded into Adam."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_npz, split


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dstd_gcn_b200  # noqa: F401
        from dstd_gcn_b200 import _lib
        from dstd_gcn_b200.engine import TrainStep
        from dstd_gcn_b200.model import dstdgcn as std
        from oracle.abi_emul import EmulBackend
        _lib._set_backend_for_tests(EmulBackend())
        z = load_npz("train_std.npz")
        m = std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m").double()
        sd = m.state_dict()
        for k in sd:
            sd[k] = split(z, "p.")[k].clone()
            if rank != 0 and sd[k].is_floating_point():
                sd[k] = sd[k] + 0.37        # only rank 0 holds the real weights / running statistics ...
        m.load_state_dict(sd)
        m.train()
        step = TrainStep(m, lr=3e-3 if rank == 0 else 1.0, inverse=True)      # ... and the real learning rate:
        assert step.world == world                                            # TrainStep broadcasts rank 0's state
        assert abs(step.lr - 3e-3) < 1e-12
        losses = []
        for s in range(2):
            sl = slice(rank * 2, rank * 2 + 2)          # contiguous shard of the 4-sample golden batch
            losses.append(float(step(z[f"inputs{s}"][sl].contiguous(), z[f"inputs_inv{s}"][sl].contiguous(),
                                     z[f"targets{s}"][sl].contiguous())))
        flat = step.flat.param.detach().clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            q.put(("ok", losses, [g.numpy() for g in gathered], {k: v.numpy() for k, v in m.state_dict().items()}))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("err", repr(e), None, None))
        raise
    finally:
        dist.destroy_process_group()


def _emulate(world, steps):
    """Sequential emulation with the oracle: per-shard loss/backward, mean of grads, torch Adam."""
    from oracle import dstd_oracle as orc
    z = load_npz("train_std.npz")
    p = {k: v.clone() for k, v in split(z, "p.").items()}
    from dstd_gcn_b200.model import dstdgcn as std
    m = std.DSTDGCN(6, 4, 6, 0.0, 22, 8, 2, "h36m").double()
    req = {k: prm.requires_grad for k, prm in m.named_parameters()}
    params = {k: v.requires_grad_(True) for k, v in p.items() if req.get(k, False)}
    opt = torch.optim.Adam(list(params.values()), lr=3e-3)
    for s in range(steps):
        grads = {k: torch.zeros_like(v) for k, v in params.items()}
        bufs = None
        for r in range(world):
            pr = {k: (v if k in params else v.clone()) for k, v in p.items()}   # per-rank BN buffers
            sl = slice(r * 2, r * 2 + 2)
            loss = orc.train_loss(pr, z[f"inputs{s}"][sl], z[f"inputs_inv{s}"][sl], z[f"targets{s}"][sl])
            gs = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
            for (k, _), g in zip(params.items(), gs):
                if g is not None:
                    grads[k] += g / world
            if r == 0:
                bufs = {k: v for k, v in pr.items() if k not in params}
        for k, v in params.items():
            v.grad = grads[k]
        opt.step()
        for k, v in bufs.items():            # rank 0's running statistics are the ones that get checkpointed
            p[k] = v
    return {k: v.detach() for k, v in p.items()}


@pytest.mark.timeout(600)
def test_two_rank_gloo_step_matches_emulated_data_parallel_reference():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    status, losses, flats, sd = q.get(timeout=500)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert status == "ok", losses
    # replicas stay bit-identical: same all-reduced gradient, same Adam update
    assert (flats[0] == flats[1]).all()
    ref = _emulate(world, 2)
    for k, v in ref.items():
        if k.endswith("residual.0.bias") or not v.is_floating_point():
            continue                            # zero true gradient: Adam turns rounding noise into +-lr steps
        got = torch.from_numpy(sd[k])
        assert float((got - v).abs().max()) < 1e-8 * max(1.0, float(v.abs().max())), k
