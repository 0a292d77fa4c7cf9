"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed all-reduce of the flat gradient buffer,
parameter broadcast, gradient scale, batch sharding.  The CUDA kernels are not involved (they have no CPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_segmentation_b200.parallel import DataParallelUNet, shard_batch


class _FakePlan:
    def __init__(self, total, rank):
        self.grad_total = total
        self.flat_grad = torch.arange(total, dtype=torch.float32) * (rank + 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                      # different init per rank: broadcast must equalise
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
        dp = DataParallelUNet(model, bucket_mb=4096 * 4 / (1 << 20))    # 4096-element buckets
        ref = [torch.zeros_like(t) for t in list(model.parameters()) + list(model.buffers())]
        for t, r in zip(list(model.parameters()) + list(model.buffers()), ref):
            r.copy_(t)
            dist.broadcast(r, src=0)
            assert torch.equal(t, r)
        assert model._grad_scale == 1.0 / world
        total = 10000
        plan = _FakePlan(total, rank)
        launches = []
        orig = dp._launch

        def spy(p, end):
            launches.append((dp._sent, end))
            orig(p, end)
        dp._launch = spy
        for end in (1000, 3000, 4500, 7000, 9999):     # segment boundaries reported by the engine
            dp._on_bucket(plan, end)
        dp._on_done(plan)
        expect = torch.arange(total, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(plan.flat_grad, expect)
        # segments are coalesced until a bucket is full; the tail is flushed at the end
        assert launches == [(0, 4500), (4500, 9999), (9999, 10000)], launches
        assert dp._sent == 0 and dp._works == []
        # a second pass works on a fresh buffer
        plan2 = _FakePlan(total, rank)
        dp._on_bucket(plan2, 5000)
        dp._on_done(plan2)
        assert torch.equal(plan2.flat_grad, expect)
        x = torch.arange(8).view(8, 1)
        assert torch.equal(shard_batch(x, rank, world), x[rank * 4:(rank + 1) * 4])
        dp.detach()
        assert not hasattr(model, "_bucket_hook")
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_requires_initialised_process_group():
    with pytest.raises(RuntimeError):
        DataParallelUNet(torch.nn.Linear(2, 2))
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(5, 1), 0, 2)
