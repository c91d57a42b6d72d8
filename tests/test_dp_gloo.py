"""CPU, world_size 2 over gloo: the data-parallel host logic - bucketed, overlapped gradient all-reduce on the flat
buffer (FlatGradReducer) and the width-bucket sharder."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vistaocr_b200.optim import FlatGradReducer
    torch.manual_seed(0)  # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.Tanh(), torch.nn.Linear(33, 17), torch.nn.Tanh(),
                              torch.nn.Linear(17, 5))
    red = FlatGradReducer(list(net.parameters()), bucket_elems=100, overlap=True)
    assert len(red.buckets) >= 3
    res = []
    for step in range(2):
        g = torch.Generator().manual_seed(100 * step + rank)  # different data per rank
        x = torch.randn(8, 20, generator=g)
        red.zero()
        net(x).pow(2).sum().backward()
        local = red.flat_g.clone()  # may already contain reduced buckets: recompute the local gradient separately
        red.finish()
        res.append(red.flat_g.clone())
    # reference: both ranks' gradients computed locally, summed
    want = []
    for step in range(2):
        tot = None
        for r in range(world):
            torch.manual_seed(0)
            ref = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.Tanh(), torch.nn.Linear(33, 17),
                                      torch.nn.Tanh(), torch.nn.Linear(17, 5))
            g = torch.Generator().manual_seed(100 * step + r)
            ref(torch.randn(8, 20, generator=g)).pow(2).sum().backward()
            flat = torch.zeros_like(red.flat_g)
            for p, o in zip(ref.parameters(), red.offsets):
                flat[o:o + p.numel()] = p.grad.flatten()
            tot = flat if tot is None else tot + flat
        want.append(tot)
    ok = all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(res, want))
    views_ok = all(p.grad is gv for p, gv in zip(red.params, red.gviews))
    q.put((rank, ok, views_ok, red.collectives, len(red.buckets)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, views_ok, collectives, nb in out:
        assert ok and views_ok, (rank, ok, views_ok)
        assert collectives == 2 * nb  # every bucket reduced exactly once per step


def test_width_bucket_sharding_is_a_balanced_partition():
    from vistaocr_b200.sharding import bucket_of, shard_batches
    rng = np.random.default_rng(0)
    widths = rng.integers(60, 2400, size=5000)
    world, bs, h = 4, 16, 60
    plans = [shard_batches(widths, h, bs, world, r) for r in range(world)]
    assert len({len(p) for p in plans}) == 1  # same number of steps on every rank
    seen = set()
    for p in plans:
        for batch in p:
            assert len(batch) == bs
            ws = [int(widths[i]) for i in batch]
            assert ws == sorted(ws, reverse=True)  # SortByWidthCollater contract
            assert len({bucket_of(w, h) for w in ws}) == 1
            assert not (seen & set(batch))
            seen |= set(batch)
    for step in range(len(plans[0])):  # all ranks draw from the same bucket at each step
        assert len({bucket_of(int(widths[p[step][0]]), h) for p in plans}) == 1
    # keeping the tails loses nothing
    full = [shard_batches(widths, h, bs, world, r, drop_last=False) for r in range(world)]
    assert sorted(i for p in full for b in p for i in b) == list(range(5000))
