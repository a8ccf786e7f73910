"""Thin torch-tensor front end over the C ABI (include/videomv_b200.h).

PyTorch is used for device memory and streams only; every op below launches the library's own sm_100a kernels on
the current CUDA stream (so they are captured by torch.cuda.graph).  Activations are channels-last fp16 `[rows, C]`.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import AttnParams, GemmParams, check

LINEAR, CONV3X3, TCONV3, CONV3X3_S2, UPCONV3X3 = 0, 1, 2, 3, 4
ACT_NONE, ACT_SILU, ACT_GEGLU = 0, 1, 2

# Optional per-call timing hook (bench.py roofline leg): when set to a list, every op appends
# (kernel family, algorithmic FLOPs, algorithmic bytes, start event, end event).
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, family, flops, nbytes, desc="", replay=None):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((family, flops, nbytes, e0, e1, desc, replay))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: torch.Tensor, what: str, dtype=torch.float16) -> None:
    if not (t.is_cuda and t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError(f"{what}: expected a CUDA {dtype} [rows, C] tensor with unit inner stride, got "
                         f"{tuple(t.shape)} {t.dtype} strides {t.stride()} on {t.device}")


def gemm(a1: torch.Tensor, w: torch.Tensor, *, a2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None, rows_per_group: int = 0,
         residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, mode: int = LINEAR,
         geom: Optional[Tuple[int, int, int, int]] = None, block_n: int = 0, stages: int = 0, split_k: int = 0,
         workspace: Optional[torch.Tensor] = None, variant: int = 0, ln_stats: Optional[torch.Tensor] = None,
         ln_colsum: Optional[torch.Tensor] = None, w_static: bool = False, ln_src: Optional[Tuple[int, int]] = None,
         ln_eps: float = 1e-5, rowstats_out: Optional[torch.Tensor] = None, scatter=None) -> torch.Tensor:
    """D = epilogue(A (*) W^T).  See `vmv_gemm` in include/videomv_b200.h.

    a1 [M,K1] (linear) or the channels-last activation [B*F*H*W, Cin] (conv modes, geom=(B,F,H,W));
    w [N,Ktot] fp16 packed by videomv_b200.packing.
    """
    _rows(a1, "gemm a1")
    _rows(w, "gemm w")
    M, K1 = a1.shape
    N = w.shape[0]
    if mode == CONV3X3_S2:                    # geom describes the input images; the output has a quarter of the pixels
        M //= 4
    elif mode == UPCONV3X3:                   # 4 output pixels per input pixel; w holds one [N, 4*Cin] block per phase
        M, N = M * 4, N // 4
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float16, device=a1.device)
    _rows(out, "gemm out")
    p = GemmParams()
    p.mode, p.M, p.N, p.K1 = mode, M, N, K1
    p.A1, p.lda1 = a1.data_ptr(), a1.stride(0)
    if a2 is not None:
        _rows(a2, "gemm a2")
        p.A2, p.lda2, p.K2 = a2.data_ptr(), a2.stride(0), a2.shape[1]
    p.W, p.ldw = w.data_ptr(), w.stride(0)
    p.D, p.ldd = out.data_ptr(), out.stride(0)
    if geom is not None:
        p.B, p.F, p.H, p.Wd = geom
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("gemm bias must be fp32 [N]")
        p.bias = bias.data_ptr()
    if rowbias is not None:
        _rows(rowbias, "gemm rowbias")
        p.rowbias, p.ld_rowbias, p.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), rows_per_group
    if residual is not None:
        _rows(residual, "gemm residual")
        p.residual, p.ldr = residual.data_ptr(), residual.stride(0)
    if ln_stats is not None:
        # ln_src = (N, block_n) of the upstream GEMM whose `rowstats_out` this is; None: {mean, rstd} rows
        nsl = 1 if ln_src is None else rowstats_slots(*ln_src)
        if ln_stats.dtype != torch.float32 or ln_stats.numel() != 2 * M * nsl or ln_colsum is None or ln_colsum.numel() != N:
            raise ValueError("gemm ln_stats must be fp32 [M,2] (or [M,slots,2] with ln_src) and ln_colsum fp32 [N]")
        p.ln_stats, p.ln_colsum, p.ln_eps = ln_stats.data_ptr(), ln_colsum.data_ptr(), float(ln_eps)
        if ln_src is not None:
            p.ln_stats_src_n, p.ln_stats_src_bn = int(ln_src[0]), int(ln_src[1])
    p.act, p.block_n, p.stages, p.split_k, p.variant = act, block_n, stages, split_k, variant
    if rowstats_out is not None:
        nsl = rowstats_slots(N, gemm_block_n(N, act, block_n, variant))
        if rowstats_out.dtype != torch.float32 or rowstats_out.numel() != 2 * M * nsl or not rowstats_out.is_contiguous():
            raise ValueError(f"gemm rowstats_out must be a contiguous fp32 [M,{nsl},2] tensor")
        p.rowstats_out = rowstats_out.data_ptr()
    p.w_static = 1 if w_static else 0
    if scatter is not None:                     # _lib.GemmScatter: rows go to the peers' tensors of the other sharding layout; `out` = mine
        p.scatter = ctypes.pointer(scatter)
    if split_k > 1:
        need = _lib.lib().vmv_gemm_workspace_bytes(ctypes.byref(p))
        if need < 0:
            check(1, "gemm (workspace query)")
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=a1.device)
        p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    e0 = _prof_begin()
    check(_lib.lib().vmv_gemm(ctypes.byref(p), _stream()), "vmv_gemm")
    if e0 is not None:
        ktot = 9 * K1 if mode == UPCONV3X3 else w.shape[1]          # algorithmic FLOPs: the 9-tap conv on the upsampled image
        keep = (a1, a2, w, out, bias, rowbias, residual, workspace, ln_stats, ln_colsum, rowstats_out, scatter)   # alive for replays
        replay = lambda p=p, keep=keep: check(_lib.lib().vmv_gemm(ctypes.byref(p), _stream()), "vmv_gemm")
        k_in = K1 + (a2.shape[1] if a2 is not None else 0)               # activation columns actually read (conv modes: Cin)
        _prof_end(e0, "gemm_tc", 2.0 * M * N * ktot, 2.0 * (M * k_in + N * ktot + M * n_out * (2 if residual is not None else 1)),
                  f"mode{mode} M{M} N{N} K{ktot} act{act} res{int(residual is not None)} rb{int(rowbias is not None)} "
                  f"split{split_k}" + ("" if scatter is None else f" scatter{scatter.direction}P{scatter.world}"), replay)
    return out


def gemm_block_n(N: int, act: int = ACT_NONE, block_n: int = 0, variant: int = 0) -> int:
    """The N-tile width vmv_gemm picks for these parameters (sizes `rowstats_out`)."""
    p = GemmParams()
    p.N, p.act, p.block_n, p.variant = N, act, block_n, variant
    bn = int(_lib.lib().vmv_gemm_block_n(ctypes.byref(p)))
    if bn <= 0:
        check(1, "vmv_gemm_block_n")
    return bn


def rowstats_slots(N: int, bn: int) -> int:
    """Partial-statistic slots per row written by a GEMM with N columns in tiles of bn: (N tile, epilogue warp of the quarter)."""
    return int(_lib.lib().vmv_gemm_epilogue_split()) * ((N + bn - 1) // bn)


class GnArena:
    """Bump-allocated scratch for the statistics of one forward.
      * `bar`  : arrival-barrier words of the GroupNorm kernels ({count, generation} per chunk).  Zeroed ONCE here; the
                 barriers reset themselves, so any later call may reuse any word (no per-forward memset).
      * `buf`  : uninitialised data: per-CTA GroupNorm partial sums (`take`) and the LayerNorm partial-statistics slots the
                 upstream GEMM's epilogue writes (`take_rowstats`); every word is written before it is read.
    `reset()` rewinds both -- call it once per forward."""

    def __init__(self, device, nbytes: int = 8 << 20, bar_bytes: int = 1 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.bar = torch.zeros(bar_bytes, dtype=torch.uint8, device=device)
        self.off = 0
        self.bar_off = 0

    def reset(self):
        self.off = 0
        self.bar_off = 0

    def _bump(self, n: int) -> int:
        if self.off + n > self.buf.numel():
            raise RuntimeError("GnArena exhausted: raise its size or call reset() once per forward")
        o = self.off
        self.off += (n + 255) // 256 * 256
        return o

    def take(self, C: int, rows_per_batch: int, nbatch: int) -> Tuple[int, int]:
        """(barriers pointer, scratch pointer) for one statistics-producing GroupNorm call."""
        nb = (nbatch * 8 + 63) // 64 * 64
        if self.bar_off + nb > self.bar.numel():
            raise RuntimeError("GnArena barrier region exhausted: call reset() once per forward")
        bo = self.bar_off
        self.bar_off += nb
        need = int(_lib.lib().vmv_groupnorm_scratch_bytes(C, rows_per_batch, nbatch))
        return self.bar.data_ptr() + bo, self.buf.data_ptr() + self._bump(need)

    def take_rowstats(self, rows: int, nslots: int) -> torch.Tensor:
        o = self._bump(rows * nslots * 8)
        return self.buf[o:o + rows * nslots * 8].view(torch.float32).view(rows, nslots, 2)


def groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, rows_per_batch: int, eps: float,
              silu: bool, x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              stats: Optional[torch.Tensor] = None, reduce_fn=None, stat_rows: int = 0,
              scratch: Optional["GnArena"] = None, peer=None) -> torch.Tensor:
    """GroupNorm(32) over [x1 | x2] rows, statistics per chunk of `rows_per_batch` rows, optional SiLU.
    `reduce_fn(stats)` (e.g. an all-reduce over pixel shards) runs between the statistics and apply kernels;
    `stat_rows` is then the global number of rows per chunk.  With a `scratch` arena (and no reduction) the
    single-launch kernel is used: statistics + in-kernel barrier + apply."""
    _rows(x1, "groupnorm x1")
    rows, C1 = x1.shape
    C2 = 0
    if x2 is not None:
        _rows(x2, "groupnorm x2")
        C2 = x2.shape[1]
    nbatch = rows // rows_per_batch
    if nbatch * rows_per_batch != rows:
        raise ValueError("groupnorm: rows not divisible by rows_per_batch")
    L = _lib.lib()
    st = _stream()
    if scratch is not None and peer is not None and _gn_smem_fits(x1.device, rows_per_batch, nbatch, C1 + C2):
        # multi-GPU pixel-sharded chunk: single-pass kernel with the cross-GPU statistics sum inside (NVLink peer memory)
        if out is None:
            out = torch.empty((rows, C1 + C2), dtype=torch.float16, device=x1.device)
        ar = peer.peer
        d_off = ar.take_data(peer.world * nbatch * 64 * 8)
        c_off = ar.take_ctrl(nbatch * 64)                    # one control line per chunk: flags +0, epoch +32
        gp = _lib.GnPeer()
        gp.world, gp.rank, gp.stat_rows = peer.world, peer.rank, int(stat_rows)
        for q in range(peer.world):
            gp.slots[q] = peer.base(q) + d_off
            gp.flags[q] = peer.base(q) + c_off
        gp.epoch = peer.base(peer.rank) + c_off + 32
        bars, scr = scratch.take(C1 + C2, rows_per_batch, nbatch)
        e0 = _prof_begin()

        def launch_peer():
            check(L.vmv_groupnorm_fused_peer(x1.data_ptr(), x1.stride(0), C1, _p(x2), 0 if x2 is None else x2.stride(0), C2,
                                             rows_per_batch, nbatch, bars, scr, gamma.data_ptr(), beta.data_ptr(),
                                             float(eps), int(silu), out.data_ptr(), out.stride(0), ctypes.byref(gp), _stream()),
                  "vmv_groupnorm_fused_peer")
        launch_peer()
        if e0 is not None:
            launch_peer.keep = (x1, x2, out, gamma, beta, scratch, gp)
            _prof_end(e0, "groupnorm_peer", 0.0, 2.0 * 2 * rows * (C1 + C2),
                      f"rows{rows} C{C1}+{C2} rpb{rows_per_batch} silu{int(silu)} P{peer.world}", launch_peer)
        peer.peer_ops += 1
        return out
    if scratch is not None and reduce_fn is None:
        if out is None:
            out = torch.empty((rows, C1 + C2), dtype=torch.float16, device=x1.device)
        e0 = _prof_begin()
        bars, scr = scratch.take(C1 + C2, rows_per_batch, nbatch)

        def launch():
            check(L.vmv_groupnorm_fused(x1.data_ptr(), x1.stride(0), C1, _p(x2), 0 if x2 is None else x2.stride(0), C2,
                                        rows_per_batch, nbatch, bars, scr, gamma.data_ptr(), beta.data_ptr(),
                                        float(eps), int(silu), out.data_ptr(), out.stride(0), _stream()), "vmv_groupnorm_fused")
        launch()
        if e0 is not None:
            launch.keep = (x1, x2, out, gamma, beta, scratch)
            # algorithmic bytes: one read + one write of the tensor (what the smem-resident kernel moves through HBM)
            _prof_end(e0, "groupnorm", 0.0, 2.0 * 2 * rows * (C1 + C2),
                      f"rows{rows} C{C1}+{C2} rpb{rows_per_batch} silu{int(silu)}", launch)
        return out
    if stats is None:
        stats = torch.empty(nbatch * 64, dtype=torch.float64, device=x1.device)
    if out is None:
        out = torch.empty((rows, C1 + C2), dtype=torch.float16, device=x1.device)
    e0 = _prof_begin()
    if scratch is not None:
        bars, scr = scratch.take(C1 + C2, rows_per_batch, nbatch)
        keep = None
    else:                                                    # stand-alone call: private barrier words + scratch
        kb = torch.zeros(nbatch * 2, dtype=torch.int32, device=x1.device)
        ks = torch.empty(int(L.vmv_groupnorm_scratch_bytes(C1 + C2, rows_per_batch, nbatch)), dtype=torch.uint8, device=x1.device)
        bars, scr, keep = kb.data_ptr(), ks.data_ptr(), (kb, ks)
    check(L.vmv_groupnorm_stats(x1.data_ptr(), x1.stride(0), C1, _p(x2), 0 if x2 is None else x2.stride(0), C2,
                                rows_per_batch, nbatch, stats.data_ptr(), bars, scr, st), "vmv_groupnorm_stats")
    if reduce_fn is not None:
        reduce_fn(stats)
    check(L.vmv_groupnorm_apply(x1.data_ptr(), x1.stride(0), C1, _p(x2), 0 if x2 is None else x2.stride(0), C2,
                                rows_per_batch, nbatch, stats.data_ptr(), stat_rows, gamma.data_ptr(), beta.data_ptr(),
                                float(eps), int(silu), out.data_ptr(), out.stride(0), st), "vmv_groupnorm_apply")
    _prof_end(e0, "groupnorm", 0.0, 2.0 * 3 * rows * (C1 + C2))     # stats read + apply read + write
    del keep
    return out


def _gn_smem_fits(device, rows_per_batch: int, nbatch: int, C: int) -> bool:
    """Does vmv_groupnorm_fused take the smem-resident single-pass kernel for this shape (asked of the library itself)?"""
    with torch.cuda.device(device):
        return bool(_lib.lib().vmv_groupnorm_fused_fits_smem(C, rows_per_batch, nbatch))


def layernorm_stats(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Per-row {mean, rstd} fp32 [M,2] for a LayerNorm folded into the consuming GEMM (see `gemm(ln_stats=...)`)."""
    _rows(x, "layernorm_stats x")
    out = torch.empty((x.shape[0], 2), dtype=torch.float32, device=x.device)
    e0 = _prof_begin()
    check(_lib.lib().vmv_layernorm_stats(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(eps), out.data_ptr(),
                                         _stream()), "vmv_layernorm_stats")
    _prof_end(e0, "layernorm", 0.0, 2.0 * x.shape[0] * x.shape[1])
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _rows(x, "layernorm x")
    if out is None:
        out = torch.empty_like(x)
    e0 = _prof_begin()
    check(_lib.lib().vmv_layernorm(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], gamma.data_ptr(),
                                   beta.data_ptr(), float(eps), out.data_ptr(), out.stride(0), _stream()),
          "vmv_layernorm")
    _prof_end(e0, "layernorm", 0.0, 2.0 * 2 * x.shape[0] * x.shape[1])
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, outer: int, inner: int,
              heads: int, nq: int, nk: int, q_strides, k_strides, v_strides, o_strides, kv_group: int = 1,
              scale: float = 0.125, impl: int = 0) -> torch.Tensor:
    """softmax(q k^T scale) v with explicit (outer, inner, row) element strides; q/k/v/out are any fp16 CUDA
    tensors whose data_ptr is the first element of head 0."""
    p = AttnParams()
    p.q, p.k, p.v, p.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    p.outer, p.inner, p.heads, p.nq, p.nk = outer, inner, heads, nq, nk
    p.q_bs_outer, p.q_bs_inner, p.q_rs = q_strides
    p.k_bs_outer, p.k_bs_inner, p.k_rs = k_strides
    p.v_bs_outer, p.v_bs_inner, p.v_rs = v_strides
    p.o_bs_outer, p.o_bs_inner, p.o_rs = o_strides
    p.kv_group, p.scale, p.impl = kv_group, scale, impl
    e0 = _prof_begin()
    check(_lib.lib().vmv_attention(ctypes.byref(p), _stream()), "vmv_attention")
    if e0 is not None:
        nb = outer * inner * heads
        keep = (q, k, v, out)
        replay = lambda p=p, keep=keep: check(_lib.lib().vmv_attention(ctypes.byref(p), _stream()), "vmv_attention")
        _prof_end(e0, "attention", 4.0 * nb * nq * nk * 64, 2.0 * 64 * nb * (2 * nq + 2 * nk / kv_group),
                  f"outer{outer} inner{inner} h{heads} nq{nq} nk{nk} kvg{kv_group}", replay)
    return out


def upsample_nearest2x(x: torch.Tensor, n: int, H: int, W: int) -> torch.Tensor:
    _rows(x, "upsample x")
    C = x.shape[1]
    assert x.is_contiguous()
    out = torch.empty((n * 4 * H * W, C), dtype=torch.float16, device=x.device)
    check(_lib.lib().vmv_upsample_nearest2x(x.data_ptr(), n, H, W, C, out.data_ptr(), _stream()), "vmv_upsample_nearest2x")
    return out


def im2col_3x3_s2(x: torch.Tensor, n: int, H: int, W: int) -> torch.Tensor:
    _rows(x, "im2col x")
    C = x.shape[1]
    assert x.is_contiguous()
    out = torch.empty((n * (H // 2) * (W // 2), 9 * C), dtype=torch.float16, device=x.device)
    check(_lib.lib().vmv_im2col_3x3_s2(x.data_ptr(), n, H, W, C, out.data_ptr(), _stream()), "vmv_im2col_3x3_s2")
    return out


def conv3x3_in(x1: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x1 fp32 [B,C1,F,H,W] (+x2 [B,C2,F,H,W]) -> fp16 [B*F*H*W, Cout]."""
    assert x1.dtype == torch.float32 and x1.is_contiguous() and x1.is_cuda
    B, C1, F, H, W = x1.shape
    C2 = 0
    if x2 is not None:
        assert x2.dtype == torch.float32 and x2.is_contiguous() and x2.shape[0] == B and x2.shape[2:] == x1.shape[2:]
        C2 = x2.shape[1]
    Cout = w.shape[0]
    assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[1] == C1 + C2
    out = torch.empty((B * F * H * W, Cout), dtype=torch.float16, device=x1.device)
    check(_lib.lib().vmv_conv3x3_in(x1.data_ptr(), C1, _p(x2), C2, B, F, H, W, w.data_ptr(), bias.data_ptr(), Cout,
                                    out.data_ptr(), _stream()), "vmv_conv3x3_in")
    return out


def conv3x3_out(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, B: int, F: int, H: int, W: int) -> torch.Tensor:
    """x fp16 [B*F*H*W, C] -> fp32 [B,Cout,F,H,W]."""
    _rows(x, "conv3x3_out x")
    assert x.is_contiguous() and w.dtype == torch.float32 and w.is_contiguous()
    Cout = w.shape[0]
    out = torch.empty((B, Cout, F, H, W), dtype=torch.float32, device=x.device)
    check(_lib.lib().vmv_conv3x3_out(x.data_ptr(), B, F, H, W, x.shape[1], w.data_ptr(), bias.data_ptr(), Cout,
                                     out.data_ptr(), _stream()), "vmv_conv3x3_out")
    return out


def softmax_rows(x: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(scale * x) over the last dim of fp16 rows."""
    _rows(x, "softmax_rows x")
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().vmv_softmax_rows(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(scale), out.data_ptr(),
                                      out.stride(0), _stream()), "vmv_softmax_rows")
    return out


def rows_to_ncfhw(x: torch.Tensor, B: int, F: int, H: int, W: int, cout: int) -> torch.Tensor:
    """fp16 rows [B*F*H*W, >=cout] -> fp32 [B,cout,F,H,W]."""
    _rows(x, "rows_to_ncfhw x")
    out = torch.empty((B, cout, F, H, W), dtype=torch.float32, device=x.device)
    check(_lib.lib().vmv_rows_to_ncfhw(x.data_ptr(), x.stride(0), B, F, H, W, cout, out.data_ptr(), _stream()),
          "vmv_rows_to_ncfhw")
    return out


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    assert t.dtype == torch.int64 and t.is_cuda and t.is_contiguous()
    out = torch.empty((t.shape[0], dim), dtype=torch.float16, device=t.device)
    check(_lib.lib().vmv_sinusoidal_embedding(t.data_ptr(), t.shape[0], dim, out.data_ptr(), _stream()),
          "vmv_sinusoidal_embedding")
    return out


def embed_combine_silu(t_emb: torch.Tensor, t_emb2: Optional[torch.Tensor], cam_emb: Optional[torch.Tensor], B: int,
                       F: int) -> torch.Tensor:
    _rows(t_emb, "embed t_emb")
    E = t_emb.shape[1]
    assert t_emb.is_contiguous() and (cam_emb is None or cam_emb.is_contiguous())
    out = torch.empty((B * F, E), dtype=torch.float16, device=t_emb.device)
    check(_lib.lib().vmv_embed_combine_silu(t_emb.data_ptr(), _p(t_emb2), _p(cam_emb), B, F, E, out.data_ptr(),
                                            _stream()), "vmv_embed_combine_silu")
    return out


def cfg_ddim_step(xt: torch.Tensor, y_out: torch.Tensor, u_out: torch.Tensor, coef7: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    for t in (xt, y_out, u_out):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    if out is None:
        out = torch.empty_like(xt)
    check(_lib.lib().vmv_cfg_ddim_step(xt.data_ptr(), y_out.data_ptr(), u_out.data_ptr(), coef7.data_ptr(),
                                       xt.numel(), out.data_ptr(), _stream()), "vmv_cfg_ddim_step")
    return out


def launch_count() -> int:
    return int(_lib.lib().vmv_launch_count())
