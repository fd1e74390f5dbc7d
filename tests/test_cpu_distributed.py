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


class _Obj:
    pass


def _mock_phased(cuts, n_text=4):
    """PhasedBackward's range bookkeeping on a CPU stand-in for the flat layout (2-D weights in the arena, then the
    vectors), with the parameter names of the real model."""
    from torch import nn
    from univl_b200.ddp import PhasedBackward

    def layer():
        m = nn.Module()
        m.w = nn.Linear(8, 8)
        return m
    model = nn.Module()
    model.bert = nn.Module()
    model.bert.embeddings = nn.Module()
    model.bert.embeddings.word_embeddings = nn.Embedding(50, 8)
    model.bert.embeddings.LayerNorm = nn.LayerNorm(8)
    model.bert.encoder = nn.Module()
    model.bert.encoder.layer = nn.ModuleList([layer() for _ in range(n_text)])
    model.visual = nn.ModuleList([layer() for _ in range(2)])
    model.cls = nn.Linear(8, 50, bias=False)
    model.cls.weight = model.bert.embeddings.word_embeddings.weight     # tied, as the MLM / decoder heads are
    flat = _Obj()
    params = list(model.parameters())
    two_d = [p for p in params if p.dim() == 2]
    one_d = [p for p in params if p.dim() != 2]
    flat.params, offs, off = two_d + one_d, [], 0
    for p in two_d:
        offs.append(off)
        off += (p.numel() + 63) // 64 * 64
    flat.arena = _Obj()
    flat.arena.entries = [(p, o, p.numel()) for p, o in zip(two_d, offs)]
    flat.arena.buf = torch.empty(off)
    for p in one_d:
        offs.append(off)
        off += (p.numel() + 63) // 64 * 64
    flat.total = off
    flat.by_id = {id(p): (o, p.numel(), tuple(p.shape)) for p, o in zip(flat.params, offs)}
    return model, flat, PhasedBackward(model, flat, cuts)


def test_phased_backward_ranges_tile_the_buffer():
    model, flat, ph = _mock_phased((3, 1))
    assert ph.n_phases == 3 and ph.cuts == [3, 1]
    seen = torch.zeros(flat.total, dtype=torch.int32)
    for runs in ph.ranges:
        for a, b in runs:
            assert a % 64 == 0 and a < b <= flat.total
            seen[a:b] += 1
    assert bool((seen == 1).all()) and ph.covered() == flat.total

    def phase_of(p):
        off = flat.by_id[id(p)][0]
        return [i for i, runs in enumerate(ph.ranges) if any(a <= off < b for a, b in runs)][0]
    enc = model.bert.encoder.layer
    assert phase_of(enc[3].w.weight) == 0 and phase_of(model.visual[0].w.weight) == 0
    assert phase_of(enc[2].w.weight) == 1 and phase_of(enc[1].w.weight) == 1
    assert phase_of(enc[0].w.weight) == 2
    assert phase_of(model.bert.embeddings.word_embeddings.weight) == 2      # tied table: complete only at the end
    assert phase_of(enc[3].w.bias) == 2 and phase_of(model.visual[1].w.bias) == 2   # vectors ride with the last phase
    assert model.bert.encoder.__dict__["_cut_layers"] == frozenset((3, 1))
    ph.remove()
    assert "_cut_layers" not in model.bert.encoder.__dict__
    with pytest.raises(ValueError):
        _mock_phased((0,))
    with pytest.raises(ValueError):
        _mock_phased((4,))


def _worker_ranges(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from univl_b200.ddp import FlatGradReducer
        _, flat, ph = _mock_phased((2,))
        n = flat.total
        ok = True
        for compress in (None, "bf16"):
            g = torch.arange(n, dtype=torch.float32) % 61 * (rank + 1)
            red = FlatGradReducer(torch.zeros(n), g, broadcast=False, compress=compress)
            works = []
            for runs in ph.ranges:                         # bench.py: pack + async all-reduce per phase, one wait
                red.pack(runs)
                works += red.all_reduce_ranges(runs)
            for w in works:
                w.wait()
            red.unpack()
            want = torch.arange(n, dtype=torch.float32) % 61 * sum(r + 1 for r in range(world))
            ok &= bool(torch.equal(g, want))               # integers < 2^8: exact in bf16 too
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_phase_ranges_all_reduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_ranges, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
