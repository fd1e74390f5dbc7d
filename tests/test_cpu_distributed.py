"""CPU, world_size 2 over gloo: the flat-gradient reducer (univl_b200/ddp.py) — parameter broadcast at construction and
the in-place bucketed sum all-reduce whose mean is folded into the optimizer's grad_scale."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
        from univl_b200.ddp import FlatGradReducer
        n = 10_000
        p = torch.full((n,), float(rank + 1))
        g = torch.arange(n, dtype=torch.float32) * (rank + 1)
        red = FlatGradReducer(p, g, n_buckets=3)
        ok = bool((p == 1.0).all())                       # broadcast from rank 0
        ok &= len(red.slices) == 3 and red.slices[0][0] == 0 and red.slices[-1][1] == n
        ok &= all(a % 1024 == 0 for a, _ in red.slices)
        works = red.all_reduce(async_op=True)
        for w in works:
            w.wait()
        want = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
        ok &= bool(torch.equal(g, want))
        ok &= abs(red.grad_scale - 1.0 / world) < 1e-12
        # bf16 payload: same sums up to bf16 rounding of the addends and of the result (2^-8 relative in total)
        g2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
        red2 = FlatGradReducer(p, g2, n_buckets=2, broadcast=False, compress="bf16")
        red2.all_reduce()
        ok &= bool(((g2 - want).abs() <= 2.0 ** -7 * want.abs() + 1e-6).all()) and red2.payload.dtype == torch.bfloat16
        g3 = torch.arange(n, dtype=torch.float32) * (rank + 1)
        red3 = FlatGradReducer(p, g3, n_buckets=2, broadcast=False, compress="bf16")
        red3.pack()                                        # the split form bench.py captures into its two graphs
        red3.all_reduce(packed=True)
        red3.unpack()
        ok &= bool(torch.equal(g3, g2))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_reducer_single_process_is_noop():
    from univl_b200.ddp import FlatGradReducer
    g = torch.ones(100)
    red = FlatGradReducer(torch.zeros(100), g, n_buckets=4)
    assert red.all_reduce() == [] and red.grad_scale == 1.0 and bool((g == 1).all())
