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
