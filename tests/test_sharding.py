"""CPU: the multi-GPU sampling partition (sharding.py) -- bounds, uneven / empty shards, and a world_size-2 gloo run of
the gather path with a stand-in per-shard sampler (the CUDA sampler itself is covered by the -m gpu tests)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_models_collection_b200.sharding import gather_shards, shard_bounds, sharded_call


def test_shard_bounds_cover_and_are_contiguous():
    for total in (0, 1, 7, 16, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_sampler(local_shape, y, noise):
    # per-sample function of (noise, label): what a denoiser + scheduler is, as far as sharding is concerned
    return noise * 2.0 + y.float().view(-1, 1, 1, 1)


def _worker(rank, world, total, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    noise = torch.randn(total, 3, 4, 4, generator=g)
    y = torch.randint(0, 11, (total,), generator=g)
    out = sharded_call(_fake_sampler, (total, 3, 4, 4), y, noise, rank, world)
    lo, hi = shard_bounds(total, rank, world)
    out2 = sharded_call(_fake_sampler, (total, 3, 4, 4), y[lo:hi], noise[lo:hi], rank, world, sliced=True)
    want = _fake_sampler(None, y, noise)
    q.put((rank, bool(torch.equal(out, want)), bool(torch.equal(out2, want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 1])
def test_world2_gloo_gather_equals_single_rank(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + total
    procs = [ctx.Process(target=_worker, args=(r, 2, total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_single_rank_is_identity():
    x = torch.arange(12.0).view(3, 4)
    assert gather_shards(x, 3, 0, 1) is x
