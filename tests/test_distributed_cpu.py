"""world_size-2 gloo test of the multi-GPU host logic (sharding + the single
all-gather) with a stand-in per-rank solver."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from upright_b200 import distributed as D


def test_shard_bounds_cover_batch():
    for total in (1, 7, 16, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(total, 5, generator=g)
    target = torch.randn(total, 3, 3, generator=g)

    def fake_solve(x, t, body):
        X = x[:, None, :] * torch.arange(1, 4)[None, :, None]
        U = t.sum(dim=2)
        return dict(X=X, U=U, status=torch.full((x.shape[0],), rank, dtype=torch.int32))

    out = D.sharded_solve(fake_solve, x0, target)
    ok = torch.allclose(out["X"], x0[:, None, :] * torch.arange(1, 4)[None, :, None]) and torch.allclose(out["U"], target.sum(dim=2))
    lo, hi = D.shard_bounds(total, world, 1)
    ok = ok and bool((out["status"][lo:hi] == 1).all()) and bool((out["status"][:lo] == 0).all())
    results[rank] = bool(ok)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(total):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), total, results), nprocs=2, join=True)
    assert results[0] and results[1]


def test_sharded_solve_even():
    _run(8)


def test_sharded_solve_uneven():
    _run(7)


def test_pipelined_gather_views_are_consistent():
    """Layout of the packed [X | U] buffers of PipelinedSolveGather (pure indexing, no device needed)."""
    import torch
    B, N, nx, nu, world = 3, 4, 5, 2, 2
    nX, nU = (N + 1) * nx, N * nu
    per = B * (nX + nU)
    full = torch.arange(world * per, dtype=torch.float32)
    v = full.view(world, per)
    X = v[:, : B * nX].reshape(world, B, N + 1, nx)
    U = v[:, B * nX:].reshape(world, B, N, nu)
    # rank r, instance b: X rows start at r*per + b*nX, U rows at r*per + B*nX + b*nU
    assert X[1, 2, 0, 0] == 1 * per + 2 * nX and U[1, 2, 0, 0] == 1 * per + B * nX + 2 * nU
    assert X[0, 1, 3, 4] == 1 * nX + 3 * nx + 4 and U[0, 0, 3, 1] == B * nX + 3 * nu + 1
