"""CPU, world_size 2, gloo: the host-side data-parallel logic (ray sharding, bucketed gradient all-reduce,
visibility reduction).  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from occnerf_b200 import distributed as D


def test_shard_range_partitions_everything():
    for n in (0, 1, 6144, 6145, 262144, 1000003):
        for world in (1, 2, 4, 8):
            for granule in (1, 1024):
                spans = [D.shard_range(n, r, world, granule) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                for (b0, e0), (b1, e1) in zip(spans[:-1], spans[1:]):
                    assert e0 == b1 and b0 <= e0
                assert all(b % granule == 0 or b == n for b, _e in spans)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # two parameter tensors: one "hash table" (large -> own message), a few small ones (one bucket)
        params = [torch.nn.Parameter(torch.zeros(300000, 2)), torch.nn.Parameter(torch.zeros(256, 68)), torch.nn.Parameter(torch.zeros(256))]
        full = [torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape) / p.numel() for p in params]
        for p, f in zip(params, full):
            p.grad = f * (rank + 1)                       # rank r contributes (r+1) * f
        extra = torch.full((25, 4), float(rank))
        D.allreduce_gradients(params, extra=[extra], average=True, bucket_bytes=4 << 20)
        scale = sum(r + 1 for r in range(world)) / world
        for p, f in zip(params, full):
            assert torch.allclose(p.grad, f * scale, rtol=1e-6)
        assert torch.allclose(extra, torch.full((25, 4), sum(range(world)) / world))
        # GradReducer: the persistent flat bucket for the small tensors, in-place reduction of the large one, twice (buffer reuse)
        red = D.GradReducer(bucket_bytes=1 << 20)
        for rep in range(2):
            grads = [f * (rank + 1 + rep) for f in full] + [None]
            red(grads)
            tot = sum(r + 1 + rep for r in range(world))
            for gr, f in zip(grads, full):
                assert torch.allclose(gr, f * tot, rtol=1e-6)
        assert red.flat is not None and red.flat.numel() == 256 * 68 + 256
        # structural zeros: only the active block of a 5-D "first deconvolution" gradient is exchanged; votes ride along as a sum
        conv = torch.zeros(1024, 4, 4, 4, 4)
        conv[:, :, 1:3, 1:3, 1:3] = float(rank + 1)
        params5 = [torch.nn.Parameter(torch.zeros(1024, 4, 4, 4, 4)), torch.nn.Parameter(torch.zeros(3))]
        act = D.structural_zero_slices(params5)
        assert list(act) == [0]
        red2 = D.GradReducer(bucket_bytes=1 << 10, active=act)
        votes = torch.zeros(50)
        votes[rank * 5:rank * 5 + 10] = 1.0
        small = torch.full((3,), float(rank))
        red2([conv, small], hits=votes)
        assert float(conv[:, :, 1:3, 1:3, 1:3].min()) == float(sum(r + 1 for r in range(world))) and float(conv[:, :, 0].abs().max()) == 0.0
        assert torch.equal(small, torch.full((3,), float(sum(range(world)))))
        assert float(votes.max()) == 1.0 and int(votes.sum()) == 15
        hits = torch.zeros(100)
        hits[rank * 10:rank * 10 + 5] = 1.0
        D.allreduce_visibility(hits)
        assert int(hits.sum()) == 5 * world
        rays = torch.arange(6144 * 8, dtype=torch.float32).reshape(6144, 8)
        mine = D.shard_rays(rays, rank, world, granule=1024)
        back = D.gather_rays(mine * 2, 6144, rank, world, granule=1024)
        assert torch.equal(back, rays * 2)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_and_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
