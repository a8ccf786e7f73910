"""Frame-sharded execution of ONE sample over P GPUs (BASELINE config 5).

The reference has no such mode (its multi-GPU inference is independent replicas); this is new design.  Each rank owns
F/P consecutive frames of every sample.  Every *spatial* op (3x3 convs, per-frame GroupNorm, spatial/cross attention,
all Linear layers) is frame-local and runs unchanged on the local rows.  The three temporal couplings

    (1) temporal self-attention over the F frames of a pixel        (util.py:1061-1065, 17 blocks)
    (2) the four (3,1,1) temporal convolutions of every ResBlock tail (util.py:1381-1392, 22 blocks)
    (3) 5-D GroupNorm statistics over (C/32, F, H, W)                (util.py:1014,1358-1372)

are handled by switching layouts around each temporal segment (Ulysses-style):

    layout A  "frame shard":  rows = (b, f_local, pixel)        x_A [B * F/P * HW, C]
    layout B  "pixel shard":  rows = (b, f, pixel_local)        x_B [B * F * HW/P, C]

`frames_to_pixels` / `pixels_to_frames` are one all-to-all each; inside layout B (1) and (2) are local, and (3) needs
one all-reduce of 2*32*B doubles between the statistics and apply kernels.  Collectives go through torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) and are CUDA-graph capturable.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

# How layouts are exchanged.  "a2a": one all_to_all_single (minimal traffic, NCCL send/recv channels).  "gather": one
# all_gather_into_tensor + local slicing (P x the traffic, but only the collective transports that all_reduce uses; the
# payloads here are <= 31 MB so the exchange is latency- not bandwidth-bound either way).  Default "gather": on the
# sandboxed 2-GPU box of round 1 the first NCCL send/recv never completed (profiles/r1_multi_gpu.md).
EXCHANGE = os.environ.get("VMV_SHARD_EXCHANGE", "gather")


class ShardCtx:
    """Process-group view for frame sharding: `world` ranks, this rank owns frames [rank*Fl, (rank+1)*Fl)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.collectives = 0          # issued per forward (reported by bench.py)

    def check(self, frames: int, hw_min: int):
        if frames % self.world != 0:
            raise ValueError(f"frame sharding needs F={frames} divisible by the {self.world} ranks")
        if hw_min % self.world != 0:
            raise ValueError(f"frame sharding needs every level's H*W (min {hw_min}) divisible by {self.world}")


def _gather(x: torch.Tensor, ctx: ShardCtx) -> torch.Tensor:
    x = x.contiguous()
    flat = torch.empty((ctx.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=ctx.group)      # concatenation along dim 0 (accepted by NCCL and gloo)
    ctx.collectives += 1
    return flat.view((ctx.world,) + tuple(x.shape))


def frames_to_pixels(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx) -> torch.Tensor:
    """layout A [B*Fl*HW, C] -> layout B [B*F*HWl, C]   (F = Fl*P, HWl = HW/P)."""
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
    if EXCHANGE == "gather":
        g = _gather(x.reshape(B, Fl, HW, C), ctx)                              # [P(src frames), B, Fl, HW, C]
        mine = g[:, :, :, ctx.rank * HWl:(ctx.rank + 1) * HWl]                 # my pixel chunk of every frame
        return mine.permute(1, 0, 2, 3, 4).reshape(B * P * Fl * HWl, C)
    # [B, Fl, P, HWl, C] -> [P(dest), B, Fl, HWl, C]
    send = x.reshape(B, Fl, P, HWl, C).permute(2, 0, 1, 3, 4).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=ctx.group)
    ctx.collectives += 1
    # [P(src), B, Fl, HWl, C] -> [B, P(src)*Fl = F, HWl, C]
    return recv.permute(1, 0, 2, 3, 4).reshape(B * P * Fl * HWl, C)


def pixels_to_frames(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx) -> torch.Tensor:
    """layout B [B*F*HWl, C] -> layout A [B*Fl*HW, C]."""
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
    if EXCHANGE == "gather":
        g = _gather(x.reshape(B, P * Fl, HWl, C), ctx)                         # [P(src pixels), B, F, HWl, C]
        mine = g[:, :, ctx.rank * Fl:(ctx.rank + 1) * Fl]                      # my frames of every pixel chunk
        return mine.permute(1, 2, 0, 3, 4).reshape(B * Fl * P * HWl, C)
    # [B, P(dest frames), Fl, HWl, C] -> [P(dest), B, Fl, HWl, C]
    send = x.reshape(B, P, Fl, HWl, C).permute(1, 0, 2, 3, 4).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=ctx.group)
    ctx.collectives += 1
    # [P(src pixels), B, Fl, HWl, C] -> [B, Fl, P*HWl = HW, C]
    return recv.permute(1, 2, 0, 3, 4).reshape(B * Fl * P * HWl, C)


def allreduce_stats(stats: torch.Tensor, ctx: ShardCtx) -> None:
    """Sum the (fp64) GroupNorm partial statistics of the pixel shards: every rank then holds the 5-D statistics."""
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=ctx.group)
    ctx.collectives += 1


def gather_frames(out: torch.Tensor, ctx: ShardCtx) -> torch.Tensor:
    """[B, C, Fl, h, w] per rank -> [B, C, F, h, w] on every rank (final output only)."""
    g = _gather(out, ctx)                                                      # [P, B, C, Fl, h, w]
    return g.permute(1, 2, 0, 3, 4, 5).reshape(out.shape[0], out.shape[1], -1, out.shape[3], out.shape[4]).contiguous()
