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
