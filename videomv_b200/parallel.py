"""ONE sample over P GPUs (BASELINE config 5): frame sharding, optionally x CFG splitting.

The reference has no such mode (its multi-GPU inference is independent replicas); this is new design.

    P ranks = cfg_ways (1 or 2) x frame_ways.

* CFG split (cfg_ways = 2).  The two UNet evaluations of a DDIM step (conditional / unconditional,
  diffusion_ddim.py:149-155) are independent: rank group 0 evaluates the conditional half, group 1 the unconditional
  one, each at batch 1.  The only exchange is the output: one all-gather per UNet call (393 KB per half).
* Frame sharding inside a group (frame_ways = P / cfg_ways).  Each rank owns F/frame_ways consecutive frames.  Every
  *spatial* op (3x3 convs, per-frame GroupNorm, spatial/cross attention, all Linear layers) is frame-local and runs
  unchanged on the local rows.  The three temporal couplings

    (1) temporal self-attention over the F frames of a pixel        (util.py:1061-1065, 17 blocks)
    (2) the four (3,1,1) temporal convolutions of every ResBlock tail (util.py:1381-1392, 22 blocks)
    (3) 5-D GroupNorm statistics over (C/32, F, H, W)                (util.py:1014,1358-1372)

  are handled by switching layouts around each temporal segment (Ulysses-style):

    layout A  "frame shard":  rows = (b, f_local, pixel)        x_A [B * F/P * HW, C]
    layout B  "pixel shard":  rows = (b, f, pixel_local)        x_B [B * F * HW/P, C]

  `frames_to_pixels` / `pixels_to_frames` are one exchange each; inside layout B (1) and (2) are local, and (3) needs
  one all-reduce of 2*32*B doubles between the statistics and apply kernels.

At 2 GPUs the CFG split alone halves the work with ONE exchange per call; at 4 / 8 GPUs it halves the number of ranks
that meet at each of the 78 layout exchanges (2 / 4 instead of 4 / 8).  `cfg_split=False` gives pure frame sharding.

Transport.  On GPUs the default is PEER MEMORY (`VMV_SHARD_EXCHANGE=peer`): every rank owns an arena all its peers map
through CUDA IPC, and an exchange is ONE kernel per rank (csrc/peer.cu: remote 16 B stores of the slices the other ranks
need into their output tensors + an epoch flag barrier) -- no NCCL call, no staging, no separate permute pass; the
GroupNorm statistics and the output all-gather use the same flags.  The payloads are <= 8 MB per rank, so what matters is
the per-exchange latency.  `gather` / `a2a` keep the torch.distributed baseline (NCCL on GPUs, gloo in the CPU tests).
Everything is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

# How layouts are exchanged.  "a2a": one all_to_all_single (minimal traffic, NCCL send/recv channels).  "gather": one
# all_gather_into_tensor + local slicing (P x the traffic, but only the collective transports that all_reduce uses).
EXCHANGE = os.environ.get("VMV_SHARD_EXCHANGE", "")          # "" = peer on CUDA, gather otherwise


class PeerArena:
    """This rank's IPC-shared arena + the peers' mappings of theirs (all ranks of `group`).  Bump allocation in the same
    order on every rank of a participant set, so an offset names the same object everywhere.
    Layout: [control region: 64-byte lines, zero at start and NEVER reset (epochs are monotonic)] [data region: exchange
    outputs, statistics slots; rewound every forward].  Every control line has ONE layout whatever op uses it -- u32
    flags[8] at +0, the owner's epoch at +32, its done counter at +36 -- and the region is split by PARTICIPANT SET
    (first half: ops among the ranks of a frame group, second half: ops among all ranks), so a line only ever sees one
    set of ranks running one protocol: any sequence of forwards (B = 1 / CFG batch 2, 256 / 512, fused GroupNorm on or
    off) keeps flags and epochs consistent."""
    CTRL_BYTES = 4 << 20

    def __init__(self, group, world: int, rank: int, device, data_bytes: int):
        from . import _lib
        self.world, self.rank, self.group = world, rank, group
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
        self.begin_forward()

    def begin_forward(self):
        self.ctrl_off = [0, self.CTRL_BYTES // 2]      # [frame-group ops, all-rank ops]
        self.data_off = self.CTRL_BYTES

    def take_ctrl(self, nbytes: int = 64, all_ranks: bool = False) -> int:
        k = 1 if all_ranks else 0
        o = self.ctrl_off[k]
        self.ctrl_off[k] += (nbytes + 63) // 64 * 64
        if self.ctrl_off[k] > (self.CTRL_BYTES // 2) * (k + 1):
            raise RuntimeError("videomv_b200: peer arena control region exhausted")
        return o

    def take_data(self, nbytes: int) -> int:
        o = self.data_off
        self.data_off += (nbytes + 255) // 256 * 256
        if self.data_off > self.buf.numel():
            raise RuntimeError("videomv_b200: peer arena exhausted; raise VMV_PEER_ARENA_MB")
        return o


# One arena per (process group, device): set_frame_sharding() may be called repeatedly (re-enable, other exchange mode)
# and must not leave IPC mappings of a dropped buffer behind on the peers.
_ARENAS: Dict[Tuple, PeerArena] = {}


def _arena_for(group, world: int, rank: int, device) -> PeerArena:
    key = (id(group) if group is not None else 0, torch.device(device).index)
    ar = _ARENAS.get(key)
    if ar is None:
        mb = int(os.environ.get("VMV_PEER_ARENA_MB", "1536"))
        ar = _ARENAS[key] = PeerArena(group, world, rank, device, mb << 20)
    return ar


class ShardCtx:
    """Process-group view for sharding one sample: `world_all` ranks = `cfg_ways` groups of `world` ranks each; inside its
    group this rank (`rank`) owns frames [rank*Fl, (rank+1)*Fl); group `cfg_index` evaluates that half of the CFG pair."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device=None, exchange: Optional[str] = None,
                 cfg_split: bool = False):
        self.group_all = group
        self.world_all = dist.get_world_size(group)
        self.rank_all = dist.get_rank(group)
        self.cfg_ways = 2 if (cfg_split and self.world_all % 2 == 0) else 1
        self.world = self.world_all // self.cfg_ways           # ranks per frame group
        self.cfg_index = self.rank_all // self.world
        self.rank = self.rank_all % self.world
        self.members = [self.cfg_index * self.world + i for i in range(self.world)]
        self.collectives = 0          # NCCL / gloo collectives issued per forward (reported by bench.py)
        self.peer_ops = 0             # peer-memory exchange / all-reduce / peer-GroupNorm / gather kernels issued per forward
        self.fused_ops = 0            # layout exchanges done inside a GEMM epilogue (no kernel of their own)
        mode = exchange or EXCHANGE or ("peer" if (device is not None and torch.device(device).type == "cuda") else "gather")
        if mode not in ("peer", "gather", "a2a"):
            raise ValueError(f"VMV_SHARD_EXCHANGE={mode!r}: expected peer, gather or a2a")
        self.mode = mode
        # 5-D GroupNorm in the pixel layout: "1" = one smem-resident kernel with the cross-GPU statistics sum inside
        # (vmv_groupnorm_fused_peer); "0" = statistics kernel + peer all-reduce kernel + apply kernel
        self.fused_gn = os.environ.get("VMV_SHARD_FUSED_GN", "1") != "0"
        # layout exchanges: "1" = inside the epilogue of the GEMM that produces the exchanged tensor (vmv_gemm scatter);
        # "0" = one vmv_peer_exchange kernel per exchange
        self.fused_exchange = os.environ.get("VMV_SHARD_FUSED_EXCHANGE", "1") != "0"
        self.peer: Optional[PeerArena] = None
        self.group = group            # the frame group (torch.distributed modes)
        if mode == "peer":
            if self.world_all > 8:
                raise ValueError("peer-memory sharding supports up to 8 ranks (one NVSwitch domain)")
            self.peer = _arena_for(group, self.world_all, self.rank_all, device)
        elif self.cfg_ways > 1:
            # every rank must create every subgroup (torch.distributed contract)
            subs = [dist.new_group([c * self.world + i for i in range(self.world)]) for c in range(self.cfg_ways)]
            self.group = subs[self.cfg_index]

    @property
    def frame_sharded(self) -> bool:
        return self.world > 1

    def base(self, q: int) -> int:
        """Arena base address of frame-group member q (as mapped in this process)."""
        return self.peer.base[self.members[q]]

    def begin_forward(self):
        self.collectives = 0
        self.peer_ops = 0
        self.fused_ops = 0
        if self.peer is not None:
            self.peer.begin_forward()

    def check(self, frames: int, hw_min: int):
        if frames % self.world != 0:
            raise ValueError(f"frame sharding needs F={frames} divisible by the {self.world} ranks of a frame group")
        if hw_min % self.world != 0:
            raise ValueError(f"frame sharding needs every level's H*W (min {hw_min}) divisible by {self.world}")

    def describe(self) -> str:
        s = f"frames/{self.world}" if self.world > 1 else "frames/1"
        return (f"cfg/2 x {s}" if self.cfg_ways > 1 else s) + f" via {self.mode}"


def _gather(x: torch.Tensor, ctx: ShardCtx, all_ranks: bool = False) -> torch.Tensor:
    x = x.contiguous()
    world = ctx.world_all if all_ranks else ctx.world
    flat = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=ctx.group_all if all_ranks else ctx.group)   # concatenation along dim 0
    ctx.collectives += 1
    return flat.view((world,) + tuple(x.shape))


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
        p.dst[q] = ctx.base(q) + d_off
        p.flags[q] = ctx.base(q) + c_off
    p.epoch = ctx.base(ctx.rank) + c_off + 32
    p.done = ctx.base(ctx.rank) + c_off + 36
    p.world, p.rank, p.direction = P, ctx.rank, direction
    p.B, p.Fl, p.HWl, p.C = B, Fl, HWl, C
    e0 = ops._prof_begin()
    _lib.check(_lib.lib().vmv_peer_exchange(ctypes.byref(p), ops._stream()), "vmv_peer_exchange")
    if e0 is not None:                                 # tools/sharded_breakdown.py: every rank replays the same launches in lockstep
        replay = lambda p=p, keep=x: _lib.check(_lib.lib().vmv_peer_exchange(ctypes.byref(p), ops._stream()), "vmv_peer_exchange")
        ops._prof_end(e0, "peer_exchange", 0.0, 2.0 * nbytes, f"dir{direction} B{B} Fl{Fl} HWl{HWl} C{C} P{P}", replay)
    ctx.peer_ops += 1
    return ar.buf[d_off:d_off + nbytes].view(torch.float16).view(x.shape[0], C)


def make_scatter(rows: int, C: int, B: int, Fl: int, HW: int, ctx: ShardCtx, direction: int):
    """For a GEMM whose output [rows, C] should land in the OTHER layout (direction 0: frames -> pixels, 1: pixels -> frames):
    allocate the destination tensor at the same arena offset on every rank + a control line and return
    (vmv_gemm_scatter struct, this rank's destination tensor).  The exchange then happens inside the GEMM's epilogue
    (`ops.gemm(..., out=dst, scatter=struct)`): no exchange kernel, no extra pass over the tensor."""
    from . import _lib
    ar = ctx.peer
    P = ctx.world
    nbytes = rows * C * 2
    d_off = ar.take_data(nbytes)
    c_off = ar.take_ctrl(64)
    sc = _lib.GemmScatter()
    sc.world, sc.rank, sc.direction = P, ctx.rank, direction
    sc.B, sc.Fl, sc.HWl = B, Fl, HW // P
    for q in range(P):
        sc.dst[q] = ctx.base(q) + d_off
        sc.flags[q] = ctx.base(q) + c_off
    sc.epoch = ctx.base(ctx.rank) + c_off + 32
    sc.done = ctx.base(ctx.rank) + c_off + 36
    ctx.fused_ops += 1
    return sc, ar.buf[d_off:d_off + nbytes].view(torch.float16).view(rows, C)


def frames_to_pixels(x: torch.Tensor, B: int, Fl: int, HW: int, ctx: ShardCtx) -> torch.Tensor:
    """layout A [B*Fl*HW, C] -> layout B [B*F*HWl, C]   (F = Fl*P, HWl = HW/P)."""
    P = ctx.world
    C = x.shape[1]
    HWl = HW // P
    if P == 1:
        return x
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
    if P == 1:
        return x
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
    if ctx.world == 1:
        return
    if ctx.mode == "peer":
        from . import _lib, ops
        ar = ctx.peer
        n = stats.numel()
        d_off = ar.take_data(ctx.world * n * 8)
        c_off = ar.take_ctrl(64)
        p = _lib.PeerAllreduceParams()
        p.data = stats.data_ptr()
        for q in range(ctx.world):
            p.slots[q] = ctx.base(q) + d_off
            p.flags[q] = ctx.base(q) + c_off
        p.epoch = ctx.base(ctx.rank) + c_off + 32
        p.world, p.rank, p.n = ctx.world, ctx.rank, n
        _lib.check(_lib.lib().vmv_peer_allreduce_f64(ctypes.byref(p), ops._stream()), "vmv_peer_allreduce_f64")
        ctx.peer_ops += 1
        return
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=ctx.group)
    ctx.collectives += 1


def gather_output(out: torch.Tensor, ctx: ShardCtx) -> torch.Tensor:
    """This rank's [B, C, Fl, h, w] fp32 output (its frames of its CFG half) -> [cfg_ways, B, C, F, h, w] on every rank:
    the only exchange of the CFG split, and the final frame all-gather of a frame group (one collective for both)."""
    B, C, Fl, h, w = out.shape
    P, Wc, Wa = ctx.world, ctx.cfg_ways, ctx.world_all
    F = Fl * P
    if Wa == 1:
        return out.unsqueeze(0)
    out = out.contiguous()
    if ctx.mode == "peer":
        from . import _lib, ops
        ar = ctx.peer
        es = out.element_size()
        nbytes = Wc * B * C * F * h * w * es
        d_off = ar.take_data(nbytes)
        c_off = ar.take_ctrl(64, all_ranks=True)
        p = _lib.PeerAllgatherParams()
        p.src = out.data_ptr()
        for q in range(Wa):
            p.dst[q] = ar.base[q] + d_off
            p.flags[q] = ar.base[q] + c_off
        p.epoch = ar.base[ctx.rank_all] + c_off + 32
        p.done = ar.base[ctx.rank_all] + c_off + 36
        p.world, p.rank = Wa, ctx.rank_all
        p.nouter, p.inner_bytes = B * C, Fl * h * w * es
        p.dst_offset_bytes = (ctx.cfg_index * B * C * F + ctx.rank * Fl) * h * w * es
        p.dst_outer_stride_bytes = F * h * w * es
        e0 = ops._prof_begin()
        _lib.check(_lib.lib().vmv_peer_allgather(ctypes.byref(p), ops._stream()), "vmv_peer_allgather")
        if e0 is not None:
            replay = lambda p=p, keep=out: _lib.check(_lib.lib().vmv_peer_allgather(ctypes.byref(p), ops._stream()), "vmv_peer_allgather")
            ops._prof_end(e0, "peer_gather", 0.0, 2.0 * out.numel() * es, f"out {tuple(out.shape)} over {Wa} ranks", replay)
        ctx.peer_ops += 1
        return ar.buf[d_off:d_off + nbytes].view(out.dtype).view(Wc, B, C, F, h, w)
    g = _gather(out, ctx, all_ranks=True)                                      # [Wc*P, B, C, Fl, h, w]
    return g.view(Wc, P, B, C, Fl, h, w).permute(0, 2, 3, 1, 4, 5, 6).reshape(Wc, B, C, F, h, w).contiguous()


def gather_frames(out: torch.Tensor, ctx: ShardCtx) -> torch.Tensor:
    """[B, C, Fl, h, w] per rank -> [B, C, F, h, w] on every rank (no CFG split)."""
    if ctx.cfg_ways != 1:
        raise RuntimeError("gather_frames: use gather_output when the CFG pair is split over rank groups")
    return gather_output(out, ctx)[0]
