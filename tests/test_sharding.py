"""N > 1 host logic on CPU: world_size-2 gloo processes (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ir_ads_b200 import sharding


def test_shard_batch_covers_batch_exactly_once():
    for total, world in ((8, 1), (8, 2), (16, 8), (64, 8)):
        seen = []
        for r in range(world):
            start, n = sharding.shard_batch(total, world, r)
            seen.extend(range(start, start + n))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        sharding.shard_batch(9, 2, 0)
    with pytest.raises(ValueError):
        sharding.shard_batch(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ir_ads_b200 import MultiScaleDeformableAttention

    torch.manual_seed(0)                      # identical weights on every rank, as DDP guarantees
    mods = [MultiScaleDeformableAttention(embed_dim=64, num_heads=4, num_levels=2, num_points=2) for _ in range(2)]
    params = sharding.projection_parameters(mods)
    assert len(params) == 16
    for i, p in enumerate(params):            # rank-dependent gradients
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params[3].grad = None                     # a parameter that received no gradient on this rank
    bucket = sharding.GradBucket(params)
    assert all(p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr() for p in params)
    bucket.all_reduce_mean()
    mean = sum(range(1, world + 1)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, mean * (i + 1))) for i, p in enumerate(params) if i != 3)
    ok = ok and float(params[3].grad.abs().max()) == 0.0
    # autograd accumulates INTO the views, so a later backward lands in the bucket
    (params[0] * 2.0).sum().backward()
    ok = ok and torch.allclose(bucket.flat[:params[0].numel()], torch.full((params[0].numel(),), mean + 2.0))
    slowest = sharding.max_over_ranks(10.0 + rank, "cpu")
    start, n = sharding.shard_batch(8, world, rank)
    out[rank] = (ok, slowest, start, n)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == res[1][1] == 11.0
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 4, 4, 4)


def _hook_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                      # identical weights on every rank
    lins = [torch.nn.Linear(6, 6) for _ in range(3)]       # three "modules" chained, one bucket each
    sync = sharding.OverlappedGradSync([[l.weight, l.bias] for l in lins])
    ok = True
    for step in range(2):                     # two steps: the hooks re-arm
        sync.zero_()
        x = torch.full((4, 6), float(rank + 1 + step))
        h = x
        for l in lins:
            h = l(h)
        h.sum().backward()                    # hooks fire last module first, launching its all-reduce
        launched = len(sync._works)
        sync.finish()
        # reference: the same computation for every rank's input, averaged
        want = [torch.zeros_like(p) for l in lins for p in (l.weight, l.bias)]
        for r in range(world):
            ps = [p.detach().clone().requires_grad_(True) for l in lins for p in (l.weight, l.bias)]
            h = torch.full((4, 6), float(r + 1 + step))
            for i in range(3):
                h = h @ ps[2 * i].T + ps[2 * i + 1]
            h.sum().backward()
            for wv, p in zip(want, ps):
                wv += p.grad / world
        got = [p.grad for l in lins for p in (l.weight, l.bias)]
        ok = ok and launched == 3 and all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(got, want))
        ok = ok and all(p.grad.untyped_storage().data_ptr() == sync.all.flat.untyped_storage().data_ptr()
                        for l in lins for p in (l.weight, l.bias))
    out[rank] = ok
    dist.destroy_process_group()


def test_overlapped_grad_sync_hooks_world2_gloo():
    """The training config's exchange (bench.py --mode train): per-module buckets all-reduced from gradient hooks
    while backward is still running give the mean gradient over ranks, step after step."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_hook_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0] and res[1]


def test_overlapped_grad_sync_without_process_group():
    lin = torch.nn.Linear(3, 3)
    sync = sharding.OverlappedGradSync([[lin.weight, lin.bias]])
    sync.zero_()
    lin(torch.ones(2, 3)).sum().backward()
    sync.finish()
    assert torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
    sync.remove()


def test_bucket_without_process_group_is_identity():
    from ir_ads_b200 import MultiScaleDeformableAttention

    m = MultiScaleDeformableAttention(embed_dim=32, num_heads=2, num_levels=1, num_points=1)
    params = sharding.projection_parameters([m])
    for p in params:
        p.grad = torch.ones_like(p)
    sharding.GradBucket(params).all_reduce_mean()
    assert all(torch.equal(p.grad, torch.ones_like(p)) for p in params)
    assert sharding.max_over_ranks(3.5, "cpu") == 3.5
