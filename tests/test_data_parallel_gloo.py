"""world_size-2 gloo tests (CPU) of the N > 1 host logic: batch sharding is a partition, the flat gradient bucket's single
all-reduce equals the average of per-rank gradients and keeps .grad as views, and parameter broadcast makes replicas identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pasta_gan_b200 import data_parallel as dp


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        for gb in (16, 32, 7):
            seen = sorted(i for r in range(world) for i in dp.shard_indices(gb, world, r))
            assert seen == list(range(gb))
    b = dict(a=torch.arange(16).float().reshape(16, 1), c=torch.arange(32).float().reshape(16, 2))
    s = dp.shard_batch(b, 4, 1)
    assert s['a'].flatten().tolist() == [1, 5, 9, 13] and s['c'].shape == (4, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                               # different initial weights per rank on purpose
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        dp.broadcast_parameters(net, src=0)
        bucket = dp.FlatGradBucket(net.parameters())
        torch.manual_seed(100)
        x = torch.randn(8, 6)
        y = torch.randn(8, 3)
        xs = dp.shard_batch(dict(x=x, y=y), world, rank)
        bucket.zero()
        torch.nn.functional.mse_loss(net(xs['x']), xs['y'], reduction='sum').backward()
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in bucket.params)      # grads are views into the flat buffer
        bucket.allreduce()
        bucket.sanitize()
        out[rank] = (bucket.flat.clone(), torch.cat([p.detach().flatten() for p in net.parameters()]))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (g0, p0), (g1, p1) = out[0], out[1]
    assert torch.equal(p0, p1)                                 # broadcast made the replicas identical
    assert torch.allclose(g0, g1)                              # every rank holds the same averaged gradient
    # single-process reference: sum-loss over the full batch, divided by world (average of the two per-rank sums)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    torch.manual_seed(100)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    torch.nn.functional.mse_loss(net(x), y, reduction='sum').backward()
    ref = torch.cat([p.grad.flatten() for p in net.parameters()]) / world
    assert torch.allclose(g0, ref, atol=1e-5)
