"""CPU, world_size 2 over gloo: the data-parallel gradient buckets (the only collective on the path)
and the no-collective scene sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from sparse_rcnn_b200.parallel import GradientBuckets, shard_scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3))
    buckets = GradientBuckets(list(net.parameters()), n_buckets=2)
    assert len(buckets.buckets) == 2 and sum(len(b) for b in buckets.buckets) == 6
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)          # every rank sees different data
    for step in range(3):
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        buckets.zero()
        ((net(x) - y) ** 2).mean().backward()
        buckets.finish()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    q.put((rank, float((gathered[0] - gathered[1]).abs().max()), flat.tolist() if rank == 0 else None))
    dist.destroy_process_group()


def test_gradient_buckets_match_single_process_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] == 0.0 for r in res)                   # replicas stay bit-identical
    got = torch.tensor(next(r[2] for r in res if r[2] is not None))
    # single-process reference: average the two ranks' gradients by hand
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    for step in range(3):
        grads = None
        for g in gens:
            x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
            net.zero_grad()
            ((net(x) - y) ** 2).mean().backward()
            cur = [p.grad.clone() for p in net.parameters()]
            grads = cur if grads is None else [a + b for a, b in zip(grads, cur)]
        for p, gr in zip(net.parameters(), grads):
            p.grad = gr / 2
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.allclose(got, ref, atol=1e-6)


def test_scene_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        shards = [shard_scenes(19, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(19))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_flat_parameter_adam_equals_per_parameter_adam():
    """GradientBuckets.flatten_parameters: Adam over the flat buffers is the same element-wise update as Adam over the
    individual parameters (run.py:1441-1449), the modules' parameters stay live views of the flat storage, every view
    starts on a 16-byte boundary and the pad elements never move."""
    def make():
        torch.manual_seed(3)
        return nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 9), nn.ReLU(), nn.Linear(9, 3))     # odd sizes: padding
    a, b = make(), make()
    buckets = GradientBuckets(list(b.parameters()), n_buckets=2)
    flat = buckets.flatten_parameters()
    assert len(flat) == len(buckets.buckets) == 2
    for bucket, offs, fp in zip(buckets.buckets, buckets.offsets, flat):
        assert all(o % 4 == 0 for o in offs)
        for p, o in zip(bucket, offs):
            assert p.data_ptr() == fp.data_ptr() + 4 * o and p.grad.data_ptr() == fp.grad.data_ptr() + 4 * o
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-2)
    opt_b = torch.optim.Adam(flat, lr=1e-2)
    g = torch.Generator().manual_seed(5)
    for step in range(4):
        x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
        opt_a.zero_grad()
        ((a(x) - y) ** 2).mean().backward()
        opt_a.step()
        buckets.zero()
        ((b(x) - y) ** 2).mean().backward()
        buckets.finish()
        opt_b.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-7, rtol=1e-6)
    used = torch.zeros(flat[0].numel(), dtype=torch.bool)
    for p, o in zip(buckets.buckets[0], buckets.offsets[0]):
        used[o:o + p.numel()] = True
    assert (flat[0].detach()[~used] == 0).all()


def _worker_unused(rank, world, port, q):
    """Rank 1 never touches the parameters of the first-registered head, so its buckets COMPLETE in a different order than
    rank 0's; collectives must still be issued in the same (bucket) order on both ranks."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    trunk, head_a, head_b = nn.Linear(4, 8), nn.Linear(8, 8), nn.Linear(8, 2)
    params = list(trunk.parameters()) + list(head_a.parameters()) + list(head_b.parameters())
    buckets = GradientBuckets(params, n_buckets=3)
    order = []
    orig = buckets._launch
    buckets._launch = lambda bi: (order.append(bi), orig(bi))[1]
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(5, 4, generator=g)
    buckets.zero()
    h = trunk(x)
    if rank == 0:
        loss = head_b(head_a(h)).sum()
    else:
        loss = head_a(h).sum()                      # head_b (the bucket that finishes first on rank 0) gets no gradient here
    loss.backward()
    buckets.finish()
    flat = torch.cat([f.reshape(-1) for f in buckets.flats])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    q.put((rank, order, float((gathered[0] - gathered[1]).abs().max()), len(buckets.buckets)))
    dist.destroy_process_group()


def test_bucket_collectives_keep_a_fixed_order_with_unused_parameters():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_unused, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, order, diff, nb in res:
        assert nb >= 2 and order == list(range(nb)), (rank, order)
        assert diff == 0.0


def test_flat_parameters_expose_the_optimizers_version_counter():
    """Adam over the flat buffers never bumps the module Parameters' version counters; the packed
    tensor-core weight images key their staleness on `_wver`, which must change with every optimizer step."""
    from sparse_rcnn_b200.scn.functions import _wver
    net = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 2))
    buckets = GradientBuckets(list(net.parameters()), n_buckets=2)
    opt = torch.optim.Adam(buckets.flatten_parameters(), lr=1e-2)
    w = net[0].weight
    before = _wver(w)
    buckets.zero()
    net(torch.randn(3, 4)).sum().backward()
    buckets.finish()
    opt.step()
    assert _wver(w) != before
    assert w._version == before[0]                  # the module Parameter's own counter never moves
    # fused optimizers do not even move the flat buffer's counter (measured on the B200 box): the global optimizer-step
    # hook must have bumped the generation
    assert _wver(w)[3] > before[3]


def _worker_groups(rank, world, port, q):
    """Explicit buckets (the executor's decoder / coarse / fine phases): mean over the ranks, parameters that were not listed
    land in a last bucket, duplicates are rejected."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a, b, c = nn.Linear(4, 8), nn.Linear(8, 8), nn.Linear(8, 2)
    params = list(a.parameters()) + list(b.parameters()) + list(c.parameters())
    try:
        GradientBuckets(params, groups=[list(c.parameters()), list(c.parameters())])
        dup = False
    except ValueError:
        dup = True
    buckets = GradientBuckets(params, groups=[list(c.parameters()), list(b.parameters())])      # `a` is not listed
    assert not buckets.average_in_collective                                                   # gloo: sum, then divide
    g = torch.Generator().manual_seed(11 + rank)
    x = torch.randn(6, 4, generator=g)
    # this rank's own gradients from an untouched copy (the buckets' allreduce may already run during the backward below)
    import copy
    a2, b2, c2 = copy.deepcopy(a), copy.deepcopy(b), copy.deepcopy(c)
    for m in (a2, b2, c2):
        for p in m.parameters():
            p.grad = None
    c2(b2(a2(x))).sum().backward()
    local = [p.grad.clone() for m in (a2, b2, c2) for p in m.parameters()]
    buckets.zero()
    c(b(a(x))).sum().backward()
    buckets.finish()
    gathered = [[torch.zeros_like(t) for _ in range(world)] for t in local]
    for t, out in zip(local, gathered):
        dist.all_gather(out, t)
    err = max(float((p.grad - sum(out) / world).abs().max()) for p, out in zip(params, gathered))
    q.put((rank, dup, [len(bk) for bk in buckets.buckets], err))
    dist.destroy_process_group()


def test_explicit_bucket_groups():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_groups, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dup, sizes, err in res:
        assert dup and sizes == [2, 2, 2] and err < 1e-6, (rank, dup, sizes, err)
