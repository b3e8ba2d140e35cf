"""One-process-per-GPU data parallelism for the DIR training step, replacing
the single-process torch.nn.DataParallel of agedb-dir/train.py:143.

`DataParallel(model)` keeps the reference's call shape (`model.module.FDS`,
`model(inputs, targets, epoch)`), but each rank owns the whole replica and its
shard of the mini-batch; after backward the flat fp32 gradient buffer is
all-reduced (sum) over NCCL / NVLink in one collective and the optimizer
applies 1/world_size (`grad_scale`).  BN statistics stay per rank (reference
behaviour under DataParallel); FDS per-bin statistics are all-reduced per
epoch inside fds.FDS.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class DataParallel(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def reduce_gradients(self, async_op=False):
        """Sum the flat gradient buffer across ranks (one NCCL all-reduce)."""
        if not is_distributed():
            return None
        return dist.all_reduce(self.module.flat_grads(), op=dist.ReduceOp.SUM, async_op=async_op)

    def broadcast_parameters(self, src=0):
        if is_distributed():
            dist.broadcast(self.module.flat_parameters(), src)
            dist.broadcast(self.module._flat["running"], src)

    @property
    def grad_scale(self):
        return 1.0 / dist.get_world_size() if is_distributed() else 1.0
