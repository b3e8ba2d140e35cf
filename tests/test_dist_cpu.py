"""world_size-2 gloo tests (CPU) of the N>1 host logic: the FDS accumulator merge and the gradient all-reduce +
1/world scaling.  The kernels themselves need a GPU; here the per-rank accumulators come from numpy."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dir_oracle as O


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "imbalanced-regression_b200")]
    from fds import FDS
    from parallel import DataParallel

    # ---- FDS: each rank accumulates its shard; merged statistics == statistics of the whole set
    rng = np.random.RandomState(0)
    n, d, bn, bs = 400, 12, 30, 2
    feats = np.maximum(rng.randn(n, d) + 0.5, 0).astype(np.float32)
    labels = rng.randint(0, 36, size=n).astype(np.float32)
    flags = torch.tensor([int((labels[rank::world] == bs).any()), int((labels[rank::world] == bn - 1).any())],
                         dtype=torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)                 # as FDS.begin_epoch_stats does
    lo, hi = float(bs), float(bn - 1)
    lab = labels[rank::world]
    bins = np.where((lab >= lo) & (lab <= hi), (lab - lo).astype(np.int64), -1)
    if flags[0]:
        bins[lab < lo] = 0
    if flags[1]:
        bins[lab > hi] = bn - bs - 1
    nb = bn - bs
    acc = dict(sums=torch.zeros(nb, d, dtype=torch.float64), sumsq=torch.zeros(nb, d, dtype=torch.float64),
               counts=torch.zeros(nb, dtype=torch.int64))
    x = feats[rank::world].astype(np.float64)
    for b in range(nb):
        rows = x[bins == b]
        acc["sums"][b] = torch.from_numpy(rows.sum(0))
        acc["sumsq"][b] = torch.from_numpy((rows ** 2).sum(0))
        acc["counts"][b] = rows.shape[0]
    FDS.reduce_accumulators(acc)
    cnt, mean, var = O.fds_batch_stats(feats, labels, bn, bs)
    assert np.array_equal(acc["counts"].numpy(), cnt)
    has = cnt > 1
    nn_ = acc["counts"].numpy().astype(np.float64)[:, None]
    m = acc["sums"].numpy() / np.maximum(nn_, 1)
    v = (acc["sumsq"].numpy() - acc["sums"].numpy() * m) / np.maximum(nn_ - 1, 1)
    assert np.allclose(m[has], mean[has], rtol=1e-5, atol=1e-7)
    assert np.allclose(v[has], var[has], rtol=1e-5, atol=1e-7)

    # ---- LDS: the int64 label histograms of the rank::world shards all-reduce to the histogram of the whole
    # column, so the per-bin table / normaliser every rank derives equals the unsharded one (datasets.py:55-83)
    ages = rng.randint(0, 140, size=1001).astype(np.float32)
    hist = torch.from_numpy(O.lds_histogram(ages[rank::world], 121).astype(np.int64))
    cnt = torch.tensor([len(ages[rank::world])], dtype=torch.int64)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    assert int(cnt) == 1001 and np.array_equal(hist.numpy(), O.lds_histogram(ages, 121))

    # ---- gradient all-reduce + 1/world
    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._g = torch.full((10,), float(rank + 1))

        def flat_grads(self):
            return self._g

        def flat_parameters(self):
            return torch.zeros(10)
    dp = DataParallel(Dummy())
    dp.reduce_gradients()
    assert torch.allclose(dp.module.flat_grads() * dp.grad_scale, torch.full((10,), (1 + 2) / 2.0))
    # ---- bucketed / overlapped reduction: ranges handed over during backward + the remainder in reduce_gradients
    # cover every element exactly once (each ends up as the sum over ranks, never twice)
    class Staged(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._g = torch.arange(100, dtype=torch.float32) * (rank + 1)
            self._grad_bucket_hook = None

        def flat_grads(self):
            return self._g

        def flat_parameters(self):
            return torch.zeros(100)
    dp2 = DataParallel(Staged())
    assert dp2.module._grad_bucket_hook is not None
    dp2.module._grad_bucket_hook(60, 100)          # "layer4" finished first
    dp2.module._grad_bucket_hook(30, 60)           # then "layer3"
    dp2.reduce_gradients()                         # stem part [0, 30) + wait
    assert torch.equal(dp2.module.flat_grads(), torch.arange(100, dtype=torch.float32) * 3)
    assert not dp2._works and not dp2._done
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world2_gloo_fds_merge_and_grad_allreduce(tmp_path):
    port = 29000 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_sampler_partitions_and_pads():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "imbalanced-regression_b200")]
    from parallel import ShardSampler
    for n, world in ((191509, 8), (12208, 8), (10, 4), (3, 8), (16, 2), (7, 1)):
        for shuffle in (False, True):
            exact = [ShardSampler(n, r, world, shuffle=shuffle, pad=False, seed=5) for r in range(world)]
            padded = [ShardSampler(n, r, world, shuffle=shuffle, pad=True, seed=5) for r in range(world)]
            for s in exact + padded:
                s.set_epoch(3)
            # exact shards: every index exactly once across the ranks (the FDS collection pass)
            allidx = np.concatenate([s.indices() for s in exact])
            assert sorted(allidx.tolist()) == list(range(n))
            assert all(len(s) == len(s.indices()) for s in exact + padded)
            # padded shards: the same length on every rank (same number of all-reduces), covering every index
            lens = {len(s) for s in padded}
            assert lens == {(n + world - 1) // world}
            assert set(np.concatenate([s.indices() for s in padded]).tolist()) == set(range(n))
    # a new epoch reshuffles, identically on every rank
    a, b = ShardSampler(100, 0, 2, seed=1), ShardSampler(100, 1, 2, seed=1)
    a.set_epoch(0); b.set_epoch(0)
    e0 = np.concatenate([a.indices(), b.indices()])
    a.set_epoch(1); b.set_epoch(1)
    e1 = np.concatenate([a.indices(), b.indices()])
    assert sorted(e0.tolist()) == sorted(e1.tolist()) == list(range(100)) and not np.array_equal(e0, e1)
