"""One-process-per-GPU data parallelism for the DIR training step, replacing
the single-process torch.nn.DataParallel of agedb-dir/train.py:143.

`DataParallel(model)` keeps the reference's call shape (`model.module.FDS`,
`model(inputs, targets, epoch)`), but each rank owns the whole replica and its
shard of the mini-batch.  The flat fp32 gradient buffer is all-reduced (sum)
over NCCL / NVLink in buckets that follow the backward pass: as soon as a layer
group (layer4, layer3, layer2, layer1) has produced its gradients, its slice is
all-reduced asynchronously while the earlier layer groups still compute; the
stem / regressor slices and the wait happen in `reduce_gradients()`.  The
optimizer applies 1/world_size (`grad_scale`).  BN statistics stay per rank (reference
behaviour under DataParallel); FDS per-bin statistics are all-reduced per
epoch inside fds.FDS.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.utils.data import Sampler


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class ShardSampler(Sampler):
    """This rank's slice of the index space 0..n-1: `rank::world` of one global order (a permutation seeded by
    `seed + epoch` when `shuffle`, identical on every rank).  `pad=True` wraps the order around to a multiple of
    `world` first, so every rank yields the same number of indices -- a training epoch then runs the same number of
    steps (= gradient all-reduces) everywhere; `pad=False` gives the exact partition (every index once across the
    ranks), which is what a statistics pass over the training set needs."""

    def __init__(self, n, rank=0, world=1, shuffle=True, pad=True, seed=0):
        assert 0 <= rank < world and n >= 0
        self.n, self.rank, self.world, self.shuffle, self.pad, self.seed = n, rank, world, shuffle, pad, seed
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def indices(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g)
        else:
            order = torch.arange(self.n)
        if self.pad and self.n % self.world and self.n > 0:
            extra = self.world - self.n % self.world
            order = torch.cat([order, order[:extra]]) if extra <= self.n else order.repeat(self.world)[:self.__len__() * self.world]
        return order[self.rank::self.world].numpy()

    def __iter__(self):
        return iter(self.indices().tolist())

    def __len__(self):
        if self.pad:
            return (self.n + self.world - 1) // self.world
        return len(range(self.rank, self.n, self.world))


class DataParallel(nn.Module):
    def __init__(self, module, overlap=None):
        super().__init__()
        self.module = module
        if overlap is None:                       # DIRB200_OVERLAP_ALLREDUCE=0: one all-reduce after the backward pass
            import os
            overlap = os.environ.get("DIRB200_OVERLAP_ALLREDUCE", "1") != "0"
        self._works, self._done = [], []          # pending async all-reduces, flat ranges they cover
        if overlap and is_distributed() and hasattr(module, "_grad_bucket_hook"):
            module._grad_bucket_hook = self._reduce_bucket

    def _reduce_bucket(self, lo, hi):
        """Backward has finished the flat-gradient range [lo, hi): start its all-reduce (NCCL's stream waits for the
        kernels enqueued so far on the compute stream, then runs beside the rest of the backward pass)."""
        flat = self.module._flat["grads"] if hasattr(self.module, "_flat") else self.module.flat_grads()
        self._works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        self._done.append((lo, hi))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def reduce_gradients(self, async_op=False):
        """Sum the flat gradient buffer across ranks (one NCCL all-reduce)."""
        grads = self.module.flat_grads()     # also (re)attaches every .grad to its view of the flat buffer
        if not is_distributed():
            return None
        # whatever the bucket hook has not covered yet (the stem and the regressor; everything without the hook)
        covered, self._done = sorted(self._done), []
        pos, rest = 0, []
        for lo, hi in covered:
            if lo > pos:
                rest.append((pos, lo))
            pos = max(pos, hi)
        if pos < grads.numel():
            rest.append((pos, grads.numel()))
        for lo, hi in rest:
            self._works.append(dist.all_reduce(grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        works, self._works = self._works, []
        if async_op:
            return works
        for w in works:
            w.wait()                           # the compute stream waits; the host does not block
        return None

    def broadcast_parameters(self, src=0):
        if is_distributed():
            dist.broadcast(self.module.flat_parameters(), src)
            dist.broadcast(self.module._flat["running"], src)

    @property
    def grad_scale(self):
        return 1.0 / dist.get_world_size() if is_distributed() else 1.0
