"""Data-parallel correctness ON HARDWARE: 2 ranks over NCCL, each on its own GPU, through the sm_100a kernels.

Checks: the replicas are bit-identical after every step (same all-reduced gradient, same fused Adam
update), and they equal the K-rank emulation -- the same kernels run shard by shard on ONE device with per-shard
BatchNorm statistics, gradients summed in rank order and scaled by 1/world in Adam -- bit for bit (a two-term sum is
order independent).  Skipped with fewer than 2 GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

STEPS, PER_RANK, V, T_IN, T_OUT, C, L = 3, 8, 22, 10, 25, 16, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev, seed=777):
    from bench import perturb
    from dstd_gcn_b200.model import dstdgcn as std
    torch.manual_seed(seed)
    return perturb(std.DSTDGCN(6, T_IN, T_OUT, 0.0, V, C, L, "h36m")).to(dev).train()


def _batch(step, rank):
    from bench import synthetic_batch
    return synthetic_batch(PER_RANK, T_IN + T_OUT, V, T_IN, seed=1000 * step + rank)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dstd_gcn_b200.engine import TrainStep
        m = _model(dev, seed=777 + 13 * rank)          # different seeds: TrainStep must broadcast rank 0's replica
        step = TrainStep(m, lr=3e-3, inverse=True)
        losses = []
        for s in range(STEPS):
            losses.append(float(step(*(x.to(dev) for x in _batch(s, rank)))))
        flat = step.flat.param.detach().clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        torch.cuda.synchronize()
        if rank == 0:
            q.put(("ok", losses, [g.cpu() for g in gathered]))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("err", repr(e), None))
        raise
    finally:
        dist.destroy_process_group()


def _emulate(world):
    """K ranks one after the other on cuda:0: per-shard forward/backward (own BatchNorm statistics), gradients summed
    in rank order, one fused Adam step with grad_scale = 1/world."""
    from dstd_gcn_b200.engine import TrainStep
    dev = torch.device("cuda", 0)
    models = [_model(dev) for _ in range(world)]       # one replica per emulated rank (BN buffers are per rank)
    steps = [TrainStep(m, lr=3e-3, inverse=True) for m in models]
    for s in range(STEPS):
        total = torch.zeros_like(steps[0].flat.grad)
        for r in range(world):
            steps[r].loss_and_grads(*(x.to(dev) for x in _batch(s, r)))
            total += steps[r].flat.grad
        for r in range(world):
            st = steps[r]
            st.flat.grad.copy_(total)
            st.step_dev.add_(1)
            torch.ops.dstd_b200.adam_step(st.flat.param, st.flat.grad, st.exp_avg, st.exp_avg_sq, st.lr, st.betas[0],
                                          st.betas[1], st.eps, st.weight_decay, 1.0 / world, 0, st.lr_dev, st.step_dev)
    torch.cuda.synchronize()
    return steps[0].flat.param.detach().cpu()


@pytest.mark.timeout(900)
def test_two_rank_nccl_replicas_bit_identical_and_equal_to_emulation():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    status, losses, flats = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert status == "ok", losses
    assert torch.equal(flats[0], flats[1]), "replicas diverged"
    ref = _emulate(world)
    assert torch.equal(flats[0], ref), float((flats[0] - ref).abs().max())
