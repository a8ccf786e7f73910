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

from typing import Optional

import torch
import torch.distributed as dist


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


def frames_to_pixels(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx) -> torch.Tensor:
    """layout A [B*Fl*HW, C] -> layout B [B*F*HWl, C]   (F = Fl*P, HWl = HW/P)."""
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
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
    parts = [torch.empty_like(out) for _ in range(ctx.world)]
    dist.all_gather(parts, out.contiguous(), group=ctx.group)
    ctx.collectives += 1
    return torch.cat(parts, dim=2)
