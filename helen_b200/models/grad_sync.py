"""Gradient exchange of data-parallel training: the one collective of the training path.

The reference wraps the model in DistributedDataParallel (helen/modules/python/models/train_distributed.py:128-131),
whose only effect on the arithmetic is that every rank applies the MEAN of the ranks' gradients.  Here the gradients
of all parameters live in ONE flat buffer (each ``p.grad`` is a view of it; hb_train_step_chunk writes them in place),
so a step costs a single all-reduce of ~1 MB over NCCL / NVLink instead of one per tensor or a bucketing pass.
"""
import torch
import torch.distributed as dist


class FlatGradients(object):
    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        first = self.params[0]
        for p in self.params:
            if p.device != first.device or p.dtype != torch.float32:
                raise ValueError("FlatGradients needs fp32 parameters on one device")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=first.device)
        self.attach()

    def attach(self):
        """(Re)point every p.grad at its slice of the flat buffer (an optimizer's zero_grad(set_to_none=True) drops them)."""
        offset = 0
        for p in self.params:
            p.grad = self.flat[offset:offset + p.numel()].view_as(p)
            offset += p.numel()

    def attached(self):
        offset = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * offset:
                return False
            offset += p.numel()
        return True

    def all_reduce_mean(self, group=None):
        """Every rank ends up with the mean of all ranks' gradients.  No-op without a process group / on one rank."""
        if not dist.is_available() or not dist.is_initialized():
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        if not self.attached():
            raise RuntimeError("FlatGradients: a parameter's .grad no longer points into the flat buffer "
                               "(use zero_grad(set_to_none=False) or call attach())")
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)


class DataParallelContext(object):
    """What the training loop needs to know about its peers (rank 0 logs, evaluates and saves, like
    train_distributed.py:121-124,241-262)."""

    def __init__(self, rank, world_size, group=None):
        self.rank, self.world_size, self.group = rank, world_size, group

    @property
    def is_main(self):
        return self.rank == 0

    def broadcast_parameters(self, model):
        """Start every rank from rank 0's values (what DistributedDataParallel's constructor does)."""
        for tensor in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(tensor.data, src=0, group=self.group)

    def broadcast_value(self, value, device):
        """rank 0's number on every rank (the evaluation loss that drives ReduceLROnPlateau on all of them)."""
        box = torch.tensor([float(value) if self.is_main else 0.0], dtype=torch.float64, device=device)
        dist.broadcast(box, src=0, group=self.group)
        return float(box.item())

    def barrier(self):
        dist.barrier(group=self.group)

    def assert_replicas_identical(self, model):
        """Every rank applied the same mean gradients to the same start values, so the replicas must stay bit-identical;
        a rank that drifted (a lost all-reduce, a different data order feeding the optimizer state) is caught here, once
        per epoch, instead of silently training a different model."""
        device = next(model.parameters()).device
        mine = torch.stack([p.detach().double().sum() for p in model.parameters()] +
                           [p.detach().double().abs().sum() for p in model.parameters()])
        gathered = [torch.empty_like(mine) for _ in range(self.world_size)]
        dist.all_gather(gathered, mine.to(device), group=self.group)
        for rank, other in enumerate(gathered):
            if not torch.equal(other, gathered[0]):
                raise RuntimeError("data-parallel replicas diverged: rank %d differs from rank 0" % rank)
