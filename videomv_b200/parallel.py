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

`frames_to_pixels` / `pixels_to_frames` are one exchange each; inside layout B (1) and (2) are local, and (3) needs
one all-reduce of 2*32*B doubles between the statistics and apply kernels.

Transport.  On GPUs the default is PEER MEMORY (`VMV_SHARD_EXCHANGE=peer`): every rank owns an arena its peers map through
CUDA IPC, and an exchange is ONE kernel per rank (csrc/peer.cu: remote 16 B stores of the slices the other ranks need into
their output tensors + an epoch flag barrier) -- no NCCL call, no staging, no separate permute pass; the GroupNorm
statistics use the same flags.  The payloads are <= 8 MB per rank (11 us of NVLink time), so what matters is the
per-exchange latency: ~46 us through NCCL all_gather + strided copy vs a few us here.  `gather` / `a2a` keep the
torch.distributed baseline (NCCL on GPUs, gloo in the CPU tests).  Everything is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
import torch.distributed as dist

# How layouts are exchanged.  "a2a": one all_to_all_single (minimal traffic, NCCL send/recv channels).  "gather": one
# all_gather_into_tensor + local slicing (P x the traffic, but only the collective transports that all_reduce uses; the
# payloads here are <= 31 MB so the exchange is latency- not bandwidth-bound either way).  Default "gather": on the
# sandboxed 2-GPU box of round 1 the first NCCL send/recv never completed (profiles/r1_multi_gpu.md).
EXCHANGE = os.environ.get("VMV_SHARD_EXCHANGE", "")          # "" = peer on CUDA, gather otherwise


class PeerArena:
    """This rank's IPC-shared arena + the peers' mappings of theirs.  Bump allocation in the same order on every rank, so
    an offset names the same object everywhere.  Layout: [control region: flags / epoch / done words, zero at start and
    NEVER reset (epochs are monotonic)] [data region: exchange outputs, statistics slots; rewound every forward]."""
    CTRL_BYTES = 4 << 20

    def __init__(self, group, world: int, rank: int, device, data_bytes: int):
        from . import _lib
        self.world, self.rank = world, rank
        self.buf = torch.zeros(self.CTRL_BYTES + data_bytes, dtype=torch.uint8, device=device)
        L = _lib.lib()
        handle = (ctypes.c_uint8 * 64)()
        off = ctypes.c_int64(0)
        _lib.check(L.vmv_ipc_export(self.buf.data_ptr(), handle, ctypes.byref(off)), "vmv_ipc_export")
        mine = (bytes(handle), int(off.value), int(self.buf.numel()))
        everyone: List = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        self.base: List[int] = []
        for q, (h, o, n) in enumerate(everyone):
            if n != self.buf.numel():
                raise RuntimeError("videomv_b200: peer arenas differ in size across ranks")
            if q == rank:
                self.base.append(self.buf.data_ptr())
                continue
            out = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64).from_buffer_copy(h)
            _lib.check(L.vmv_ipc_import(hb, o, ctypes.byref(out)), f"vmv_ipc_import (arena of rank {q})")
            self.base.append(int(out.value))
        torch.cuda.synchronize()
        dist.barrier(group=group)                      # every rank has mapped every arena before anyone writes
        self.ctrl_off = 0
        self.data_off = self.CTRL_BYTES

    def begin_forward(self):
        self.ctrl_off = 0
        self.data_off = self.CTRL_BYTES

    def take_ctrl(self, nbytes: int = 64) -> int:
        o = self.ctrl_off
        self.ctrl_off += (nbytes + 63) // 64 * 64
        if self.ctrl_off > self.CTRL_BYTES:
            raise RuntimeError("videomv_b200: peer arena control region exhausted")
        return o

    def take_data(self, nbytes: int) -> int:
        o = self.data_off
        self.data_off += (nbytes + 255) // 256 * 256
        if self.data_off > self.buf.numel():
            raise RuntimeError("videomv_b200: peer arena exhausted; raise VMV_PEER_ARENA_MB")
        return o


class ShardCtx:
    """Process-group view for frame sharding: `world` ranks, this rank owns frames [rank*Fl, (rank+1)*Fl)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device=None, exchange: Optional[str] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.collectives = 0          # NCCL / gloo collectives issued per forward (reported by bench.py)
        self.peer_ops = 0             # peer-memory exchange / all-reduce kernels issued per forward
        mode = exchange or EXCHANGE or ("peer" if (device is not None and torch.device(device).type == "cuda") else "gather")
        if mode not in ("peer", "gather", "a2a"):
            raise ValueError(f"VMV_SHARD_EXCHANGE={mode!r}: expected peer, gather or a2a")
        self.mode = mode
        # 5-D GroupNorm in the pixel layout: "1" = one smem-resident kernel with the cross-GPU statistics sum inside
        # (vmv_groupnorm_fused_peer); "0" = statistics kernel + peer all-reduce kernel + apply kernel
        self.fused_gn = os.environ.get("VMV_SHARD_FUSED_GN", "1") != "0"
        self.peer: Optional[PeerArena] = None
        if mode == "peer":
            if self.world > 8:
                raise ValueError("peer-memory frame sharding supports up to 8 ranks (one NVSwitch domain)")
            mb = int(os.environ.get("VMV_PEER_ARENA_MB", "1536"))
            self.peer = PeerArena(group, self.world, self.rank, device, mb << 20)

    def begin_forward(self):
        self.collectives = 0
        self.peer_ops = 0
        if self.peer is not None:
            self.peer.begin_forward()

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


def _peer_exchange(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx, direction: int) -> torch.Tensor:
    from . import _lib, ops
    ar = ctx.peer
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
    x = x.contiguous()
    nbytes = x.numel() * 2
    d_off = ar.take_data(nbytes)                     # my output tensor: same offset in every rank's arena
    c_off = ar.take_ctrl(64)                         # [0,32) flags (8 x u32)  [32] epoch  [36] done
    p = _lib.PeerExchangeParams()
    p.src = x.data_ptr()
    for q in range(P):
        p.dst[q] = ar.base[q] + d_off
        p.flags[q] = ar.base[q] + c_off
    p.epoch = ar.base[ctx.rank] + c_off + 32
    p.done = ar.base[ctx.rank] + c_off + 36
    p.world, p.rank, p.direction = P, ctx.rank, direction
    p.B, p.Fl, p.HWl, p.C = B, Fl, HWl, C
    _lib.check(_lib.lib().vmv_peer_exchange(ctypes.byref(p), ops._stream()), "vmv_peer_exchange")
    ctx.peer_ops += 1
    return ar.buf[d_off:d_off + nbytes].view(torch.float16).view(x.shape[0], C)


def frames_to_pixels(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx) -> torch.Tensor:
    """layout A [B*Fl*HW, C] -> layout B [B*F*HWl, C]   (F = Fl*P, HWl = HW/P)."""
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
    if ctx.mode == "peer":
        return _peer_exchange(x, B, Fl, HW, ctx, 0)
    if ctx.mode == "gather":
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
    if ctx.mode == "peer":
        return _peer_exchange(x, B, Fl, HW, ctx, 1)
    if ctx.mode == "gather":
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
    if ctx.mode == "peer":
        from . import _lib, ops
        ar = ctx.peer
        n = stats.numel()
        d_off = ar.take_data(ctx.world * n * 8)
        c_off = ar.take_ctrl(64)
        p = _lib.PeerAllreduceParams()
        p.data = stats.data_ptr()
        for q in range(ctx.world):
            p.slots[q] = ar.base[q] + d_off
            p.flags[q] = ar.base[q] + c_off
        p.epoch = ar.base[ctx.rank] + c_off + 32
        p.world, p.rank, p.n = ctx.world, ctx.rank, n
        _lib.check(_lib.lib().vmv_peer_allreduce_f64(ctypes.byref(p), ops._stream()), "vmv_peer_allreduce_f64")
        ctx.peer_ops += 1
        return
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=ctx.group)
    ctx.collectives += 1


def gather_frames(out: torch.Tensor, ctx: ShardCtx) -> torch.Tensor:
    """[B, C, Fl, h, w] per rank -> [B, C, F, h, w] on every rank (final output only)."""
    g = _gather(out, ctx)                                                      # [P, B, C, Fl, h, w]
    return g.permute(1, 2, 0, 3, 4, 5).reshape(out.shape[0], out.shape[1], -1, out.shape[3], out.shape[4]).contiguous()
