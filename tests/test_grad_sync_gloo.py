"""The one collective of the training path on CPU: world_size-2 gloo group, gradients of all parameters in one flat
buffer, averaged once per step (helen_b200/models/grad_sync.py; what DistributedDataParallel does for the reference,
train_distributed.py:128-131).  The CUDA training step itself is covered by tests/test_gpu_train.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helen_b200.models.grad_sync import DataParallelContext, FlatGradients


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _data(rank):
    gen = torch.Generator().manual_seed(100 + rank)
    return torch.randn(8, 6, generator=gen), torch.randint(0, 3, (8,), generator=gen)


def _fill_grads_like_the_cuda_step(model, x, y):
    """hb_train_step_chunk OVERWRITES every gradient tensor in place; emulate that with autograd + copy_."""
    loss = torch.nn.functional.cross_entropy(model(x), y)
    grads = torch.autograd.grad(loss, list(model.parameters()))
    for p, g in zip(model.parameters(), grads):
        p.grad.copy_(g)
    return float(loss.detach())


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = DataParallelContext(rank, world)
    model = _model(seed=rank)                           # ranks start from different values ...
    ctx.broadcast_parameters(model)                     # ... and continue from rank 0's
    shared = FlatGradients(model.parameters())
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2)
    x, y = _data(rank)
    for _ in range(3):
        _fill_grads_like_the_cuda_step(model, x, y)
        shared.all_reduce_mean(ctx.group)
        optimizer.step()
    value = ctx.broadcast_value(3.25 if rank == 0 else -1.0, torch.device("cpu"))
    ctx.barrier()
    torch.save({"params": [p.detach().clone() for p in model.parameters()], "value": value,
                "attached": shared.attached()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_flat_gradient_views():
    model = _model(0)
    shared = FlatGradients(model.parameters())
    assert shared.flat.numel() == sum(p.numel() for p in model.parameters()) and shared.attached()
    x, y = _data(0)
    _fill_grads_like_the_cuda_step(model, x, y)
    offset = 0
    for p in model.parameters():                        # every p.grad IS its slice of the flat buffer
        assert p.grad.is_contiguous() and torch.equal(shared.flat[offset:offset + p.numel()].view_as(p), p.grad)
        offset += p.numel()
    shared.all_reduce_mean()                            # no process group: a no-op
    torch.optim.Adam(model.parameters()).zero_grad(set_to_none=True)
    assert not shared.attached()
    shared.attach()
    assert shared.attached()
    with pytest.raises(ValueError):
        FlatGradients([])


def test_two_ranks_average_gradients_and_stay_in_step(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert got[0]["value"] == got[1]["value"] == 3.25 and got[0]["attached"] and got[1]["attached"]
    for a, b in zip(got[0]["params"], got[1]["params"]):
        assert torch.equal(a, b)                        # same start, same averaged gradients, same Adam: bit-identical ranks
    # one process applying the mean of the two ranks' gradients itself reaches the same parameters
    model = _model(seed=0)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(3):
        per_rank = []
        for rank in range(world):
            x, y = _data(rank)
            loss = torch.nn.functional.cross_entropy(model(x), y)
            per_rank.append(torch.autograd.grad(loss, list(model.parameters())))
        for i, p in enumerate(model.parameters()):
            p.grad = (per_rank[0][i] + per_rank[1][i]) / world
        optimizer.step()
    for a, p in zip(got[0]["params"], model.parameters()):
        assert torch.allclose(a, p.detach(), rtol=0, atol=1e-7)
