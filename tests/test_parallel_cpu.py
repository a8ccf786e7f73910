"""world_size-2 (gloo, CPU) tests of the frame-shard <-> pixel-shard transposition used by the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, F, HW, C, ret, exchange="gather"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VMV_SHARD_EXCHANGE=exchange)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videomv_b200 import parallel
        ctx = parallel.ShardCtx()
        ctx.check(F, HW)
        Fl, HWl = F // world, HW // world
        g = torch.Generator().manual_seed(0)
        full = torch.randn(B, F, HW, C, generator=g)                       # the unsharded activation
        xa = full[:, rank * Fl:(rank + 1) * Fl].reshape(B * Fl * HW, C)     # layout A on this rank
        xb = parallel.frames_to_pixels(xa, B, Fl, HW, ctx)
        want_b = full[:, :, rank * HWl:(rank + 1) * HWl].reshape(B * F * HWl, C)
        ok = torch.equal(xb, want_b)
        back = parallel.pixels_to_frames(xb, B, Fl, HW, ctx)
        ok &= torch.equal(back, xa)
        stats = torch.full((B * 64,), float(rank + 1), dtype=torch.float64)
        parallel.allreduce_stats(stats, ctx)
        ok &= bool((stats == sum(range(1, world + 1))).all())
        out = torch.full((B, 4, Fl, 2, 2), float(rank))
        gathered = parallel.gather_frames(out, ctx)
        ok &= gathered.shape == (B, 4, F, 2, 2) and all(
            bool((gathered[:, :, r * Fl:(r + 1) * Fl] == r).all()) for r in range(world))
        ok &= ctx.collectives == 4
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["gather", "a2a"])
@pytest.mark.parametrize("B,F,HW,C", [(1, 24, 16, 8), (2, 4, 64, 16)])
def test_layout_transposition_world2(B, F, HW, C, exchange):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), B, F, HW, C, ret, exchange), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _cfg_worker(rank, world, port, B, F, HW, C, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VMV_SHARD_EXCHANGE="gather")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videomv_b200 import parallel
        ctx = parallel.ShardCtx(cfg_split=True)
        P = world // 2
        ok = ctx.cfg_ways == 2 and ctx.world == P and ctx.cfg_index == rank // P and ctx.rank == rank % P
        ok &= ctx.members == [ctx.cfg_index * P + i for i in range(P)]
        Fl, HWl = F // P, HW // P
        g = torch.Generator().manual_seed(0)
        full = torch.randn(2, B, F, HW, C, generator=g)                    # [cfg half, ...]: the two groups work on different data
        mine = full[ctx.cfg_index]
        xa = mine[:, ctx.rank * Fl:(ctx.rank + 1) * Fl].reshape(B * Fl * HW, C)
        xb = parallel.frames_to_pixels(xa, B, Fl, HW, ctx)                 # exchange inside my frame group only
        ok &= torch.equal(xb, mine[:, :, ctx.rank * HWl:(ctx.rank + 1) * HWl].reshape(B * F * HWl, C))
        ok &= torch.equal(parallel.pixels_to_frames(xb, B, Fl, HW, ctx), xa)
        stats = torch.full((B * 64,), float(rank + 1), dtype=torch.float64)
        parallel.allreduce_stats(stats, ctx)
        ok &= bool((stats == sum(r + 1 for r in ctx.members)).all())
        # output all-gather over ALL ranks: [cfg half, B, C, F, h, w] on every rank
        outs = torch.randn(2, B, 4, F, 2, 2, generator=g)
        local = outs[ctx.cfg_index][:, :, ctx.rank * Fl:(ctx.rank + 1) * Fl].contiguous()
        ok &= torch.equal(parallel.gather_output(local, ctx), outs)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cfg_split_times_frame_sharding(world):
    """P ranks = 2 CFG groups x P/2 frame shards: exchanges stay inside a group, the output all-gather spans all ranks."""
    ret = mp.Manager().dict()
    mp.spawn(_cfg_worker, args=(world, _free_port(), 1, 24, 16, 8, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
