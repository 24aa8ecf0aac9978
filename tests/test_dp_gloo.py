"""CPU, world_size 2, gloo: the data-parallel host logic (sharding, gradient-arena all-reduce plan, centre exchange)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_ssl_avmnist_b200 import dp


def test_shard_ranges_cover_the_batch():
    for G in (0, 1, 7, 16, 16384):
        for W in (1, 2, 3, 8):
            got = [dp.shard_range(G, r, W) for r in range(W)]
            assert got[0][0] == 0 and got[-1][1] == G
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # gradient plan: only the planned slices are summed, the rest (unused fc1/fc2 region) is untouched
        g = torch.Generator().manual_seed(10 + rank)
        flat = torch.randn(1000, generator=g)
        keep = flat.clone()
        plan = dp.GradientPlan([(0, 400), (900, 1000), (500, 500)])
        assert plan.bytes() == 4 * 500
        scale = plan.allreduce_(flat)
        assert scale == 1.0 / world
        expect = sum(torch.randn(1000, generator=torch.Generator().manual_seed(10 + r)) for r in range(world))
        assert torch.allclose(flat[:400], expect[:400]) and torch.allclose(flat[900:], expect[900:])
        assert torch.equal(flat[400:900], keep[400:900])
        # centre: per-rank column sums -> all-reduce -> EMA equals the EMA over the concatenated rows
        D, rows = 128, 6
        all_rows = torch.randn(world * rows, D, generator=torch.Generator().manual_seed(3))
        lo, hi = dp.shard_range(world * rows, rank, world)
        colsum = all_rows[lo:hi].sum(0)
        n = dp.allreduce_colsum_(colsum, hi - lo)
        assert n == world * rows
        center = torch.full((D,), 0.25)
        got = dp.center_ema_reference(center, colsum, n, 0.9)
        want = center * 0.9 + all_rows.mean(0) * (1 - 0.9)
        assert torch.allclose(got, want, atol=1e-6)
        assert dp.world_size() == world and dp.rank() == rank
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
