"""N>1 host logic on CPU: world_size-2 gloo.  A sharded step (each rank its slice, flat gradient
all-reduce, 1/world scaling) must reproduce the single-process global-batch gradient."""
import os
import socket
import sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.ReLU(), torch.nn.Linear(9, 4))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meshvae_b200 import dp
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 12, generator=g)
    y = torch.randn(8, 4, generator=g)
    lo, hi = dp.shard_bounds(8, rank, world)
    net = _toy()
    loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).sum(-1).mean()          # mean over the LOCAL slice
    loss.backward()
    live = dp.live_parameters(list(net.parameters()))
    offsets, n = dp.flat_layout(live)
    assert all(o % 32 == 0 for o in offsets)
    flat = torch.zeros(n)
    views = dp.flat_views(flat, live, offsets)
    dp.pack_grads(live, views)
    dp.allreduce_sum_(flat)
    flat.mul_(1.0 / world)
    dp.unpack_grads(views, live)
    if rank == 0:
        q.put([p.grad.clone() for p in net.parameters()])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_equals_global_batch_step():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 12, generator=g)
    y = torch.randn(8, 4, generator=g)
    net = _toy()
    ((net(x) - y) ** 2).sum(-1).mean().backward()
    for a, p in zip(grads, net.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6)


def test_shard_bounds():
    from meshvae_b200 import dp
    assert [dp.shard_bounds(512, r, 8) for r in range(8)] == [(64 * r, 64 * (r + 1)) for r in range(8)]
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)


def _ragged_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meshvae_b200 import dp, loader
    g = torch.Generator().manual_seed(12)
    n_total, bs = 11, 4                                  # global batches of 8 and 3: the last one is ragged (1 + 2 items)
    x = torch.randn(n_total, 12, generator=g)
    y = torch.randn(n_total, 4, generator=g)
    ld = loader.ShardedMeshLoader(list(range(n_total)), bs, rank=rank, world=world)
    sizes = ld.global_chunk_sizes()
    out = []
    for i, idx in enumerate(ld.index_batches()):
        net = _toy()
        full = sizes[i] == bs * world                    # the SAME decision on every rank (engine graph path vs ragged path)
        w = 1.0 if full else dp.ragged_weight(sizes[i], rank, world)
        sel = torch.as_tensor(idx)
        ((net(x[sel]) - y[sel]) ** 2).sum(-1).mean().backward()
        live = dp.live_parameters(list(net.parameters()))
        offsets, n = dp.flat_layout(live)
        flat = torch.zeros(n)
        views = dp.flat_views(flat, live, offsets)
        dp.pack_grads(live, views)
        flat.mul_(w)
        dp.allreduce_sum_(flat)
        flat.mul_(1.0 / world)
        out.append((full, flat.clone()))
    if rank == 0:
        q.put((sizes, out))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_last_global_batch_is_weighted_and_decided_collectively():
    """with drop_last=False the last global batch splits unevenly; every rank must take the same step path
    (decided from the GLOBAL batch size) and weight its local-mean gradient by local/global so that the average is the
    global-batch gradient."""
    from meshvae_b200 import dp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ragged_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sizes, out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sizes == [8, 3] and [f for f, _ in out] == [True, False]
    assert dp.ragged_weight(3, 0, 2) + dp.ragged_weight(3, 1, 2) == pytest.approx(2.0)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(11, 12, generator=g)
    y = torch.randn(11, 4, generator=g)
    for (lo, hi), (_, flat) in zip(((0, 8), (8, 11)), out):
        net = _toy()
        ((net(x[lo:hi]) - y[lo:hi]) ** 2).sum(-1).mean().backward()
        ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        live = list(net.parameters())
        offsets, n = dp.flat_layout(live)
        got = torch.cat([flat[o:o + p.numel()] for p, o in zip(live, offsets)])
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
