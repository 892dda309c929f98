"""CPU, world_size 2 over gloo: the batch-sharding host logic of the N>1 path (no GPU needed)."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("stylegan-for-facerec_b200.dist")
    torch.manual_seed(0)
    z = torch.randn(total, 8)                                   # same on every rank
    mine = d.shard_batch(z)

    class FakeG:                                                # stands in for Generator on CPU
        def __call__(self, styles, **kw):
            return styles[0] * 2 + 1, None

    img, _ = d.synthesize_sharded(FakeG(), [z], total=total, gather=True)
    ok = torch.equal(img, z * 2 + 1)
    lo, hi = d.shard_bounds(total, world, rank)
    q.put((rank, lo, hi, mine.shape[0], bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 1])                    # even, ragged, fewer samples than ranks
def test_two_rank_sharding_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total      # contiguous cover
    assert res[0][3] + res[1][3] == total and abs(res[0][3] - res[1][3]) <= 1
    assert all(r[4] for r in res)                                                 # gathered == unsharded result


def test_shard_bounds_properties():
    d = importlib.import_module("stylegan-for-facerec_b200.dist")
    for n in (0, 1, 5, 64, 1000):
        for w in (1, 2, 3, 8):
            spans = [d.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_bounds(4, 2, 2)


def _np(ts):
    return [None if t is None else t.detach().numpy().copy() for t in ts]


def _pt(ts):
    return [None if t is None else torch.from_numpy(t) for t in ts]


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("stylegan-for-facerec_b200.dist")
    torch.manual_seed(100 + rank)                               # different initial weights per rank ...
    enc = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    d.broadcast_parameters(enc, src=0)                          # ... until they are broadcast
    for p in enc[3].parameters():
        p.requires_grad_(False)                                 # a frozen tail (the decoder's role): no gradient, skipped
    torch.manual_seed(0)
    x = torch.randn(8, 6)                                       # the global batch, same on every rank
    y = torch.randn(8, 4)
    mine, tgt = d.shard_batch(x), d.shard_batch(y)
    loss = ((enc(mine) - tgt) ** 2).sum() / x.shape[0]          # per-rank share of the global mean loss
    loss.backward()
    if rank == 1:
        enc[2].bias.grad = None                                 # a rank without a gradient for one parameter: counts as zeros
    n_buckets = d.average_gradients(enc.parameters(), bucket_bytes=256)   # tiny buckets: several collectives
    grads = [None if p.grad is None else (p.grad * world).clone() for p in enc.parameters()]   # mean * world == sum over ranks
    w0 = [p.detach().clone() for p in enc.parameters()]
    q.put((rank, n_buckets, _np(grads), _np(w0)))        # by value: shared-memory tensor handles die with the worker
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_gloo():
    """the one collective of the fine-tuning path: bucketed all-reduce (mean) of the trainable gradients"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, nb0, g0, w0), (_, nb1, g1, w1) = res
    g0, w0, g1, w1 = _pt(g0), _pt(w0), _pt(g1), _pt(w1)
    assert nb0 == nb1 and nb0 >= 2                              # same bucket layout on both ranks, more than one bucket
    assert all(torch.equal(a, b) for a, b in zip(w0, w1))       # broadcast made the replicas identical
    assert all((a is None) == (b is None) and (a is None or torch.equal(a, b)) for a, b in zip(g0, g1))
    assert g0[-1] is None and g0[-2] is None                    # frozen parameters are left alone
    # single-process reference: the whole batch through the same weights
    enc = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    with torch.no_grad():
        for p, w in zip(enc.parameters(), w0):
            p.copy_(w)
    torch.manual_seed(0)
    x, y = torch.randn(8, 6), torch.randn(8, 4)
    (((enc(x) - y) ** 2).sum() / 8).backward()
    ref = [p.grad for p in enc.parameters()]
    # rank 1 dropped its share of enc[2].bias (index 3): the averaged value there is rank 0's share alone
    for i, (a, r) in enumerate(zip(g0[:4], ref[:4])):
        if i == 3:
            lo, hi = 0, 4
            enc.zero_grad()
            (((enc(x[lo:hi]) - y[lo:hi]) ** 2).sum() / 8).backward()
            r = list(enc.parameters())[3].grad
        assert torch.allclose(a, r, rtol=1e-5, atol=1e-6), i


def _averager_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("stylegan-for-facerec_b200.dist")
    torch.manual_seed(100 + rank)
    enc = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.randn(3))                 # receives a gradient on NO rank
    d.broadcast_parameters(enc, src=0)
    params = list(enc.parameters()) + [unused]
    avg = d.GradientAverager(params, bucket_bytes=256)          # several buckets
    torch.manual_seed(0)
    x, y = torch.randn(8, 6), torch.randn(8, 4)
    mine, tgt = d.shard_batch(x), d.shard_batch(y)
    out = []
    for step in range(2):                                       # second step: the buckets are re-used
        avg.zero_grad()
        for it in range(3):                                     # gradient accumulation over refinement iterations
            if it == 2:
                avg.arm()                                       # only the last backward triggers the exchange
            loss = ((enc(mine * (it + 1)) - tgt) ** 2).sum() / x.shape[0]
            loss.backward()
        n = avg.finish()
        out.append((n, [None if p.grad is None else (p.grad * world).clone() for p in params]))
    views_ok = all(p.grad is None or p.grad.data_ptr() == v.data_ptr()
                   for b in avg.buckets for p, v in zip(b["params"], b["views"]))
    q.put((rank, [(n, _np(g)) for n, g in out], views_ok, len(avg.buckets)))
    avg.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_overlapped_gradient_averager_gloo():
    """GradientAverager: gradients live in bucket views, the exchange is issued from hooks during the armed backward,
    accumulation over several backward passes is preserved, a parameter without a gradient anywhere keeps grad = None"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_averager_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, out0, ok0, nb0), (_, out1, ok1, nb1) = res
    assert ok0 and ok1 and nb0 == nb1 and nb0 >= 2
    # single-process reference: whole batch, same three accumulated passes
    torch.manual_seed(100)
    enc = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    torch.manual_seed(0)
    x, y = torch.randn(8, 6), torch.randn(8, 4)
    for it in range(3):
        (((enc(x * (it + 1)) - y) ** 2).sum() / 8).backward()
    ref = [p.grad for p in enc.parameters()]
    for (n0, g0), (n1, g1) in zip(out0, out1):
        g0, g1 = _pt(g0), _pt(g1)
        assert n0 == n1 == nb0 + 1                               # one collective per bucket + the has-gradient mask
        assert g0[-1] is None and g1[-1] is None                 # unused parameter: no gradient invented
        for a, b, r in zip(g0[:-1], g1[:-1], ref):
            assert torch.equal(a, b) and torch.allclose(a, r, rtol=1e-5, atol=1e-6)
