"""UNetEngine: executes the VideoMV video-UNet forward with the library's sm_100a kernels.

Data layout in HBM: every activation is channels-last fp16 `x[(b*F + f)*H*W + h*W + w, c]`, i.e. ONE row-major
[M, C] matrix that is simultaneously
  * the NHWC image batch the 3x3 convs tile with 4-D TMA boxes,
  * the (frame, token, channel) sequence batch of the spatial transformers,
  * the (pixel, frame, channel) sequence batch of the temporal transformers / temporal convs (strided views),
so the reference's rearrange(...).contiguous() transposes (util.py:363,370,1054-1083,726-729) and torch.cat skip
concats (unet_t2v.py:361) never materialise.  Weights are repacked once to K-major fp16 (packing.py).

Reference call graph being replaced: unet_t2v.py:283-403 / unet_i2vgen.py:287-439 and every block in util.py they
dispatch to.  Per-op reference citations live in include/videomv_b200.h.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib, ops, packing, parallel


SM_COUNT = 148


class _W:
    """Packed fp16 weight [N,K] + fp32 bias (+ column sums when a LayerNorm is folded in)."""
    __slots__ = ("w", "b", "bn", "colsum")

    def __init__(self, w, b=None, bn=0, colsum=None):
        self.w, self.b, self.bn, self.colsum = w, b, bn, colsum


def _f32(p):
    return None if p is None else p.detach().to(torch.float32).contiguous()


class UNetEngine:
    def __init__(self, module):
        self.m = module
        self.device = next(module.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("videomv_b200: move the UNet to a CUDA device before calling it (no CPU fallback)")
        self.head_dim = module.head_dim
        if self.head_dim != 64:
            raise ValueError(f"videomv_b200: the attention kernels are built for head_dim=64 (got {self.head_dim}); "
                             "both shipped configs use 64 (configs/t2v_infer.yaml:27)")
        self.variant = module.variant
        self._ws: Optional[torch.Tensor] = None        # split-K workspace
        self._ws_retired: List[torch.Tensor] = []      # outgrown workspaces: captured graphs hold their addresses
        self._ctx_cache: Dict[Tuple, torch.Tensor] = {}
        self._i2v_cache: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._graphs: Dict[Tuple, "_Graph"] = {}
        self.use_graphs = False
        self.shard: Optional[parallel.ShardCtx] = None     # frame sharding of one sample over the ranks of a group
        self._gn_arena: Optional[ops.GnArena] = None       # statistics scratch of the current forward (GroupNorm partials, LayerNorm slots)
        self._gn_arenas: Dict[Tuple[int, int, int], ops.GnArena] = {}
        self.fused_groupnorm = os.environ.get("VMV_GN_FUSED", "1") != "0"
        # VMV_RESAMPLE_LEGACY=1: Downsample through an explicit patch gather, Upsample through the materialised 4x tensor (A/B)
        self.legacy_resample = os.environ.get("VMV_RESAMPLE_LEGACY", "0") == "1"
        # LayerNorm row sums accumulated by the epilogue of the GEMM that produces the rows (no separate statistics pass);
        # needs the CTA-pair kernel's register epilogue and the per-forward arena
        self.fused_ln_stats = (self.fused_groupnorm and os.environ.get("VMV_LN_FUSED", "1") != "0"
                               and os.environ.get("VMV_GEMM_VARIANT", "2") != "1")
        self._pack()

    # ------------------------------------------------------------------------------------------------------------
    # weight packing (once per load)
    # ------------------------------------------------------------------------------------------------------------
    def _lin(self, layer, bias=True) -> _W:
        return _W(packing.pack_linear(layer.weight.detach()), _f32(layer.bias) if bias and layer.bias is not None else None)

    def _folded(self, ws, bias, norm, geglu=False) -> _W:
        """Linear(s) `ws` (row-concatenated) consuming LayerNorm `norm`: fold gamma/beta into the weights/bias."""
        w = torch.cat([z.detach().reshape(z.shape[0], -1) for z in ws], 0)
        wg, bf = packing.fold_layernorm(w, bias, norm.weight, norm.bias)
        if geglu:
            wp, bp, bn = packing.pack_geglu(wg, bf)
        else:
            wp, bp, bn = wg.to(torch.float16).contiguous(), bf, 0
        return _W(wp, bp, bn, wp.float().sum(1).contiguous())

    def _pack_tblock(self, tb, cross: bool):
        d = {}
        a1, a2 = tb.attn1, tb.attn2
        d["qkv1"] = self._folded([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], None, tb.norm1)
        d["o1"] = self._lin(a1.to_out[0])
        if cross:
            d["q2"] = self._folded([a2.to_q.weight], None, tb.norm2)
            d["kv2_w"] = packing.pack_cat(a2.to_k.weight.detach(), a2.to_v.weight.detach())   # gathered into kv_all
        else:
            d["qkv2"] = self._folded([a2.to_q.weight, a2.to_k.weight, a2.to_v.weight], None, tb.norm2)
        d["o2"] = self._lin(a2.to_out[0])
        d["ff1"] = self._folded([tb.ff.net[0].proj.weight], tb.ff.net[0].proj.bias, tb.norm3, geglu=True)
        d["ff2"] = self._lin(tb.ff.net[2])
        return d

    def _pack_block(self, mod):
        kind = getattr(mod, "kind", None)
        if kind == "res":
            d = {"kind": "res", "cin": mod.channels, "cout": mod.out_channels}
            d["gn1"] = (_f32(mod.in_layers[0].weight), _f32(mod.in_layers[0].bias))
            d["c1"] = _W(packing.pack_conv3x3(mod.in_layers[2].weight.detach()), _f32(mod.in_layers[2].bias))
            d["emb_off"] = self._emb_total
            self._emb_w.append(packing.pack_linear(mod.emb_layers[1].weight.detach()))
            self._emb_b.append(_f32(mod.emb_layers[1].bias))
            self._emb_total += mod.out_channels
            d["gn2"] = (_f32(mod.out_layers[0].weight), _f32(mod.out_layers[0].bias))
            d["c2"] = _W(packing.pack_conv3x3(mod.out_layers[3].weight.detach()), _f32(mod.out_layers[3].bias))
            d["skip"] = None
            if not isinstance(mod.skip_connection, torch.nn.Identity):
                d["skip"] = self._lin(mod.skip_connection)
            tc = mod.temopral_conv
            d["t"] = []
            for st in (tc.conv1, tc.conv2, tc.conv3, tc.conv4):
                d["t"].append(((_f32(st[0].weight), _f32(st[0].bias)),
                               _W(packing.pack_tconv3(st[-1].weight.detach()), _f32(st[-1].bias))))
            return d
        if kind == "spatial":
            d = {"kind": "spatial", "c": mod.in_channels, "heads": mod.heads}
            d["gn"] = (_f32(mod.norm.weight), _f32(mod.norm.bias))
            d["pin"], d["pout"] = self._lin(mod.proj_in), self._lin(mod.proj_out)
            d["tb"] = self._pack_tblock(mod.transformer_blocks[0], cross=True)
            d["kv_off"] = self._kv_total
            self._kv_w.append(d["tb"].pop("kv2_w"))
            self._kv_total += 2 * mod.heads * self.head_dim
            return d
        if kind == "temporal":
            d = {"kind": "temporal", "c": mod.in_channels, "heads": mod.heads}
            d["gn"] = (_f32(mod.norm.weight), _f32(mod.norm.bias))
            d["pin"], d["pout"] = self._lin(mod.proj_in), self._lin(mod.proj_out)
            d["tb"] = self._pack_tblock(mod.transformer_blocks[0], cross=False)
            return d
        if kind == "down":
            return {"kind": "down", "c": mod.op.in_channels,
                    "w": _W(packing.pack_conv3x3(mod.op.weight.detach()), _f32(mod.op.bias))}
        if kind == "up":
            if self.legacy_resample:
                return {"kind": "up", "c": mod.conv.in_channels,
                        "w": _W(packing.pack_conv3x3(mod.conv.weight.detach()), _f32(mod.conv.bias))}
            # nearest x2 + 3x3 conv as four 2x2 phase convs on the original image (weights pre-summed per phase)
            return {"kind": "up", "c": mod.conv.in_channels,
                    "w": _W(packing.pack_upconv3x3(mod.conv.weight.detach()), _f32(mod.conv.bias))}
        if isinstance(mod, torch.nn.Conv2d):
            return {"kind": "stem", "w": _f32(mod.weight), "b": _f32(mod.bias)}
        if isinstance(mod, torch.nn.ModuleList):
            return {"kind": "list", "items": [self._pack_block(s) for s in mod]}
        raise RuntimeError(f"videomv_b200: unknown block {type(mod)}")

    def _pack_mlp(self, seq, pad_k: int = 0):
        w0 = seq[0].weight.detach()
        if pad_k and w0.shape[1] < pad_k:
            w0 = F.pad(w0, (0, pad_k - w0.shape[1]))
        return (_W(packing.pack_linear(w0), _f32(seq[0].bias)), self._lin(seq[2]))

    @torch.no_grad()
    def _pack(self):
        m = self.m
        self._emb_w: List[torch.Tensor] = []
        self._emb_b: List[torch.Tensor] = []
        self._emb_total = 0
        self._kv_w: List[torch.Tensor] = []
        self._kv_total = 0
        self.time_mlp = self._pack_mlp(m.time_embed)
        self.cam_mlp = self._pack_mlp(m.camera_embedding, pad_k=64) if hasattr(m, "camera_embedding") else None
        self.fps_mlp = self._pack_mlp(m.fps_embedding) if hasattr(m, "fps_embedding") else None
        self.enc = [self._pack_block(b) for b in m.input_blocks]
        self.mid = [self._pack_block(b) for b in m.middle_block]
        self.dec = [self._pack_block(b) for b in m.output_blocks]
        self.head_gn = (_f32(m.out[0].weight), _f32(m.out[0].bias))
        self.head_w, self.head_b = _f32(m.out[2].weight), _f32(m.out[2].bias)
        # head conv on the tensor cores: Cout (4) zero-padded to 16 GEMM columns
        hw = m.out[2].weight.detach()
        self.head_cout = hw.shape[0]
        hpad = torch.zeros((16, hw.shape[1], 3, 3), dtype=hw.dtype, device=hw.device)
        hpad[: hw.shape[0]] = hw
        bpad = torch.zeros(16, dtype=torch.float32, device=hw.device)
        bpad[: hw.shape[0]] = m.out[2].bias.detach().float()
        self.head_gemm = _W(packing.pack_conv3x3(hpad), bpad, 128) if hw.shape[0] <= 16 and hw.shape[1] % 64 == 0 else None
        # one GEMM produces every ResBlock's emb projection / every cross-attention's K,V
        self.emb_all = _W(torch.cat(self._emb_w, 0).contiguous(), torch.cat(self._emb_b, 0).contiguous())
        self.kv_all = _W(torch.cat(self._kv_w, 0).contiguous())
        del self._emb_w, self._emb_b, self._kv_w

    # ------------------------------------------------------------------------------------------------------------
    # GEMM helper with the small-M split-K heuristic
    # ------------------------------------------------------------------------------------------------------------
    def _gemm(self, a, w: _W, **kw):
        M = a.shape[0]
        N = w.w.shape[0]
        mode = kw.get("mode", ops.LINEAR)
        if mode == ops.CONV3X3_S2:
            M //= 4                                    # output rows
        elif mode == ops.UPCONV3X3:
            M, N = 4 * M, N // 4
        bn = w.bn or (256 if N > 320 else (160 if N % 160 == 0 else 128))      # mirrors pick_block_n in gemm_tc.cu
        # The persistent kernel runs one CTA pair per 2 SMs (74 pairs) on 256 x bn tiles.  When the tile count is a
        # poor fit for 74 (few tiles at the 8x8 / 4x4 levels) split K so the last wave is not mostly idle.
        tiles = ((M + 255) // 256) * ((N + bn - 1) // bn)
        nkb = w.w.shape[1] // 64
        pairs = SM_COUNT // 2
        split = 0
        if kw.get("act", 0) != ops.ACT_GEGLU and nkb >= 32 and tiles < 4 * pairs and mode != ops.UPCONV3X3:
            # time(s) ~ T_full / utilisation(s) + cost of the fp32 partials (write + read back + extra launch)
            t_full = 2.0 * M * N * nkb * 64 / 0.9e15

            def est(s_):
                t_ = tiles * s_
                u = t_ / (-(-t_ // pairs) * pairs)
                extra = 0.0 if s_ == 1 else (2.0 * s_ * M * N * 4 + M * N * 2) / 5e12 + 5e-6
                return t_full / u + extra
            best = min(range(1, min(nkb // 8, 8) + 1), key=est)
            if best >= 2 and est(best) < 0.9 * est(1):
                split = best
                need = split * M * N * 4
                if self._ws is None or self._ws.numel() < need:
                    if self._ws is not None:
                        self._ws_retired.append(self._ws)          # never freed while a graph may replay into it
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                kw["workspace"] = self._ws
        # to_layout = (direction, S): this GEMM's output feeds a layout exchange of the frame-sharded multi-GPU mode
        # (0: frames -> pixels before a temporal segment, 1: pixels -> frames after it).  With peer memory the exchange
        # happens INSIDE the kernel that produces the rows (the GEMM's register epilogue, or the split-K finish kernel):
        # they are stored straight into the peers' tensors and the kernel ends with the flag rendezvous.
        to_layout = kw.pop("to_layout", None)
        fs = self._fs if to_layout is not None else None
        if fs is not None:
            direction, S_ = to_layout
            B_, Fl_, HW_ = S_["B"], S_["F"], S_["H"] * S_["W"]
            if (fs.mode == "peer" and fs.fused_exchange and kw.get("act", 0) != ops.ACT_GEGLU and N % 32 == 0
                    and not kw.get("want_stats", False)):
                sc, dst = parallel.make_scatter(M, N, B_, Fl_, HW_, fs, direction)
                ln = kw.pop("ln", None)
                if ln is not None:
                    kw.update(ln_stats=ln[0], ln_src=ln[1], ln_colsum=w.colsum)
                return ops.gemm(a, w.w, bias=w.b, block_n=w.bn, split_k=split, w_static=True, out=dst, scatter=sc, **kw)
            out = self._gemm(a, w, **kw)
            if direction == 0:
                return parallel.frames_to_pixels(out, B_, Fl_, HW_, fs)
            return parallel.pixels_to_frames(out, B_, Fl_, HW_, fs)
        # w_static: every W here is a packed model weight, so the kernel may prefetch W tiles ahead of its PDL wait
        ln = kw.pop("ln", None)                    # (stats tensor, src): statistics of the rows of `a` for the folded LayerNorm
        if ln is not None:
            kw.update(ln_stats=ln[0], ln_src=ln[1], ln_colsum=w.colsum)
        want_stats = kw.pop("want_stats", False)   # the output feeds a LayerNorm: also return its row statistics
        if not want_stats:
            return ops.gemm(a, w.w, bias=w.b, block_n=w.bn, split_k=split, w_static=True, **kw)
        if self.fused_ln_stats and split == 0 and N % 32 == 0 and self._gn_arena is not None:
            src_bn = ops.gemm_block_n(N, kw.get("act", 0), w.bn)
            nsl = ops.rowstats_slots(N, src_bn)
            if nsl <= 16:
                rs = self._gn_arena.take_rowstats(M, nsl)
                out = ops.gemm(a, w.w, bias=w.b, block_n=w.bn, split_k=0, w_static=True, rowstats_out=rs, **kw)
                return out, (rs, (N, src_bn))
        out = ops.gemm(a, w.w, bias=w.b, block_n=w.bn, split_k=split, w_static=True, **kw)
        return out, (ops.layernorm_stats(out), None)

    # ------------------------------------------------------------------------------------------------------------
    # blocks
    # ------------------------------------------------------------------------------------------------------------
    def _res(self, d, x, skip, S):
        B, Fr, H, W = S["B"], S["F"], S["H"], S["W"]
        HW = H * W
        emb = S["emb_all"][:, d["emb_off"]:d["emb_off"] + d["cout"]]
        a0 = ops.groupnorm(x, *d["gn1"], rows_per_batch=HW, eps=1e-5, silu=True, x2=skip, scratch=self._gn_arena)
        h = self._gemm(a0, d["c1"], mode=ops.CONV3X3, geom=(1, B * Fr, H, W), rowbias=emb, rows_per_group=HW)
        a1 = ops.groupnorm(h, *d["gn2"], rows_per_batch=HW, eps=1e-5, silu=True, scratch=self._gn_arena)
        if d["skip"] is not None:
            res = self._gemm(x, d["skip"], a2=skip)
        else:
            res = x
        # temporal tail (util.py:1381-1392) on the pixel-sharded layout when the sample is spread over ranks: conv2's
        # epilogue delivers its rows in that layout, the last temporal conv's epilogue brings them back
        h2t = self._gemm(a1, d["c2"], mode=ops.CONV3X3, geom=(1, B * Fr, H, W), residual=res, to_layout=(0, S))
        Ff, HWt = self._temporal_geom(S)
        cur = h2t
        for i, (gn, wt) in enumerate(d["t"]):
            a = ops.groupnorm(cur, *gn, rows_per_batch=Ff * HWt, eps=1e-5, silu=True, scratch=self._gn_arena, **self._gn5d_kw(S))
            cur = self._gemm(a, wt, mode=ops.TCONV3, geom=(B, Ff, HWt, 1), residual=h2t if i == 3 else None,
                             to_layout=(1, S) if i == 3 else None)
        return cur

    # ---- frame-shard <-> pixel-shard plumbing (identity on a single GPU) -----------------------------------------
    @property
    def _fs(self) -> Optional[parallel.ShardCtx]:
        """The sharding context when the frames of a sample are actually spread over ranks (else None)."""
        sh = self.shard
        return sh if (sh is not None and sh.frame_sharded) else None

    def _temporal_geom(self, S):
        """(frames, pixels per rank) of a temporal segment: all frames of HW/P pixels when the sample is frame-sharded."""
        HW = S["H"] * S["W"]
        ctx = self._fs
        if ctx is None:
            return S["F"], HW
        return S["F"] * ctx.world, HW // ctx.world

    def _gn5d_kw(self, S):
        ctx = self._fs
        if ctx is None:
            return {}
        kw = dict(reduce_fn=lambda st: parallel.allreduce_stats(st, ctx), stat_rows=S["F"] * ctx.world * S["H"] * S["W"])
        if ctx.mode == "peer" and ctx.fused_gn:
            kw["peer"] = ctx            # single-pass GroupNorm with the cross-GPU sum inside the kernel (when it fits smem)
        return kw

    def _tblock(self, tb, h, hst, S, heads, temporal: bool, kv=None, Fr=None, HW=None):
        """BasicTransformerBlock (util.py:536-540) on rows `h` whose LayerNorm statistics are `hst`."""
        B = S["B"]
        Fr = S["F"] if Fr is None else Fr
        HW = S["H"] * S["W"] if HW is None else HW
        C = heads * self.head_dim
        M = h.shape[0]

        def self_attn(qkv):
            o = torch.empty((M, C), dtype=torch.float16, device=h.device)
            ld = 3 * C
            if temporal:
                st = (Fr * HW * ld, ld, HW * ld)
                ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, outer=B, inner=HW, heads=heads, nq=Fr, nk=Fr,
                              q_strides=st, k_strides=st, v_strides=st, o_strides=(Fr * HW * C, C, HW * C))
            else:
                st = (HW * ld, 0, ld)
                ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, outer=B * Fr, inner=1, heads=heads, nq=HW, nk=HW,
                              q_strides=st, k_strides=st, v_strides=st, o_strides=(HW * C, 0, C))
            return o

        # The three LayerNorms are folded into the GEMMs that consume them, and their row statistics come out of the
        # epilogue of the GEMM that produced the rows (proj_in, attn1.to_out, attn2.to_out): no LayerNorm kernel at all.
        h, hst = self._gemm(self_attn(self._gemm(h, tb["qkv1"], ln=hst)), tb["o1"], residual=h, want_stats=True)
        if temporal:
            h, hst = self._gemm(self_attn(self._gemm(h, tb["qkv2"], ln=hst)), tb["o2"], residual=h, want_stats=True)
        else:
            q = self._gemm(h, tb["q2"], ln=hst)
            kview, vview, L, ldkv = kv
            o = torch.empty((M, C), dtype=torch.float16, device=h.device)
            ops.attention(q, kview, vview, o, outer=B * Fr, inner=1, heads=heads, nq=HW, nk=L,
                          q_strides=(HW * C, 0, C), k_strides=(L * ldkv, 0, ldkv), v_strides=(L * ldkv, 0, ldkv),
                          o_strides=(HW * C, 0, C), kv_group=Fr)
            h, hst = self._gemm(o, tb["o2"], residual=h, want_stats=True)
        g = self._gemm(h, tb["ff1"], act=ops.ACT_GEGLU, ln=hst)
        return self._gemm(g, tb["ff2"], residual=h)

    def _spatial(self, d, x, S, to_pixels=False):
        HW = S["H"] * S["W"]
        a = ops.groupnorm(x, *d["gn"], rows_per_batch=HW, eps=1e-6, silu=False, scratch=self._gn_arena)
        h, hst = self._gemm(a, d["pin"], want_stats=True)
        C = d["heads"] * self.head_dim
        kvall = S["kv_all"]
        off = d["kv_off"]
        kv = (kvall[:, off:off + C], kvall[:, off + C:off + 2 * C], S["L"], kvall.stride(0))
        h = self._tblock(d["tb"], h, hst, S, d["heads"], temporal=False, kv=kv)
        # a TemporalTransformer follows: proj_out's epilogue hands the rows over in the pixel-sharded layout
        return self._gemm(h, d["pout"], residual=x, to_layout=(0, S) if to_pixels else None)

    def _temporal(self, d, xt, S):
        """`xt` is already in the temporal layout (== the frame layout on one GPU)."""
        Ff, HWt = self._temporal_geom(S)
        a = ops.groupnorm(xt, *d["gn"], rows_per_batch=Ff * HWt, eps=1e-6, silu=False, scratch=self._gn_arena, **self._gn5d_kw(S))
        h, hst = self._gemm(a, d["pin"], want_stats=True)
        h = self._tblock(d["tb"], h, hst, S, d["heads"], temporal=True, Fr=Ff, HW=HWt)
        return self._gemm(h, d["pout"], residual=xt, to_layout=(1, S))

    def _run_block(self, d, x, skip, S):
        k = d["kind"]
        if k == "list":
            items = d["items"]
            for i, it in enumerate(items):
                if it["kind"] == "spatial":
                    x = self._spatial(it, x, S, to_pixels=i + 1 < len(items) and items[i + 1]["kind"] == "temporal")
                elif it["kind"] == "temporal":
                    if i == 0 or items[i - 1]["kind"] != "spatial":
                        fs = self._fs
                        if fs is not None:
                            x = parallel.frames_to_pixels(x, S["B"], S["F"], S["H"] * S["W"], fs)
                    x = self._temporal(it, x, S)
                else:
                    x = self._run_block(it, x, skip, S)
                skip = None
            return x
        if k == "res":
            return self._res(d, x, skip, S)
        if k == "down":
            n, H, W = S["B"] * S["F"], S["H"], S["W"]
            S["H"], S["W"] = H // 2, W // 2
            if self.legacy_resample:
                return self._gemm(ops.im2col_3x3_s2(x, n, H, W), d["w"])
            # Downsample.op (util.py:749): stride-2 windows read by element-strided TMA boxes, no patch gather
            return self._gemm(x, d["w"], mode=ops.CONV3X3_S2, geom=(1, n, H, W))
        if k == "up":
            n, H, W = S["B"] * S["F"], S["H"], S["W"]
            S["H"], S["W"] = 2 * H, 2 * W
            if self.legacy_resample:
                up = ops.upsample_nearest2x(x, n, H, W)
                return self._gemm(up, d["w"], mode=ops.CONV3X3, geom=(1, n, 2 * H, 2 * W))
            # Upsample (util.py:604-606) without the 4x tensor: four 2x2 phase convs, 2.25x fewer FLOPs
            return self._gemm(x, d["w"], mode=ops.UPCONV3X3, geom=(1, n, H, W))
        if k == "stem":
            return ops.conv3x3_in(S["x_in"], d["w"], d["b"], S.get("x_in2"))
        raise RuntimeError(k)

    # ------------------------------------------------------------------------------------------------------------
    # conditioning
    # ------------------------------------------------------------------------------------------------------------
    def _mlp(self, mlp, x16):
        h = self._gemm(x16, mlp[0], act=ops.ACT_SILU)
        return self._gemm(h, mlp[1])

    def _embeddings(self, t, fps, cam, B, Fr):
        dim = self.m.dim
        te = self._mlp(self.time_mlp, ops.sinusoidal_embedding(t.to(torch.int64).contiguous(), dim))
        fe = None
        if fps is not None and self.fps_mlp is not None:
            fe = self._mlp(self.fps_mlp, ops.sinusoidal_embedding(fps.to(torch.int64).contiguous(), dim))
        ce = None
        if cam is not None and self.cam_mlp is not None:
            c16 = torch.zeros((B * Fr, 64), dtype=torch.float16, device=self.device)
            c16[:, :cam.shape[-1]] = cam.reshape(B * Fr, -1)
            ce = self._mlp(self.cam_mlp, c16)
        e_silu = ops.embed_combine_silu(te, fe, ce, B, Fr)
        return self._gemm(e_silu, self.emb_all)

    def _context_kv(self, ctx: torch.Tensor):
        """K/V of every cross-attention layer for context [B, L, 1024]: one GEMM, B*L rows (the reference recomputes
        them per frame per layer per step: util.py:233-234 on a context repeated F times, unet_t2v.py:346)."""
        Bc, L, Dc = ctx.shape
        c16 = ctx.reshape(Bc * L, Dc).to(torch.float16).contiguous()
        return self._gemm(c16, self.kv_all), L

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, y, camera_data, fps, image=None, local_image=None):
        """One UNet evaluation with the reference's tensor conventions: x [B,C,F,h,w],
        t [B] int64, y [B,L,1024], camera_data [B,F,16] (may arrive on CPU), fps [B] int64."""
        if x.dim() != 5:
            raise ValueError("videomv_b200: x must be [B, C, F, h, w]")
        dev = self.device
        out_dtype = x.dtype
        x32 = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        cam = None if camera_data is None else camera_data.to(device=dev, dtype=torch.float32).contiguous()
        fps = None if fps is None else fps.to(device=dev, dtype=torch.int64).contiguous()
        kv, concat = self.prepare_condition(x32.shape, y, image, local_image)
        out = self.forward_core(x32, t, kv, cam, fps, concat)
        if self.shard is not None:
            out = out[self.shard.cfg_index]            # [cfg_ways, B, C, F, h, w]: every group evaluated the same call
        return out if out_dtype == torch.float32 else out.to(out_dtype)

    def prepare_condition(self, xshape, y, image=None, local_image=None):
        """Step-invariant conditioning: the K/V of every cross-attention layer for the context tokens [B,L,1024] (one GEMM
        per SAMPLE -- the reference recomputes them per frame, per layer, per step: util.py:233-234) + the I2V concat
        planes.  Returns ((kv_all fp16 [B*L, sum 2*C], L), concat).  Cached on the
        identity of the caller's tensors, which the sampler passes unchanged for all 50 steps.  The entry holds strong
        references to the keyed tensors: while it exists neither their id() nor their storage address can be recycled
        for another prompt's tensors (the reference engines build fresh y / image tensors per caption:
        inference_text2video_entrance.py:170,240), and `_version` catches in-place updates."""
        key = tuple((id(z), z.data_ptr(), z._version, tuple(z.shape)) if z is not None else None
                    for z in (y, image, local_image)) + (tuple(xshape),)
        hit = self._ctx_cache.get(key)
        if hit is not None:
            return hit[0], hit[1]
        dev = self.device
        if self.variant == "i2v":
            concat, ctx = self._i2v_condition(xshape, y.to(dev), image, local_image)
        else:
            concat, ctx = None, y.to(device=dev, dtype=torch.float32)
        kv = self._context_kv(ctx.contiguous())
        if len(self._ctx_cache) >= 8:
            self._ctx_cache.clear()
        self._ctx_cache[key] = (kv, concat, (y, image, local_image))
        return kv, concat

    def forward_core(self, x32, t, kv, cam, fps, concat):
        """The per-step hot path on device-resident inputs. Replays a captured CUDA graph when enabled.
        Returns [B, C, F, h, w]; with a sharding context [cfg_ways, B, C, F, h, w] (the gathered halves of a split CFG pair)."""
        kv_all, L = kv
        if not self.use_graphs:
            out = self._forward_impl(x32, t, kv_all, cam, fps, concat, L)
            return out if self.shard is None else out.clone()     # the gathered output lives in the peer arena: hand out a copy
        key = (tuple(x32.shape), tuple(kv_all.shape), L, cam is not None, fps is not None, concat is not None)
        g = self._graphs.get(key)
        if g is None:
            g = _Graph(self, L, x32, t, kv_all, cam, fps, concat)
            self._graphs[key] = g
        return g.run(x32, t, kv_all, cam, fps, concat)

    def _forward_impl(self, x32, t, kv_all, cam, fps, concat, L):
        B, _, Fr, H, W = x32.shape
        if self.fused_groupnorm:
            # one arena per (B*F) size class, kept alive for the CUDA graphs that captured pointers into it
            nb = B * Fr
            arena = self._gn_arenas.get((nb, H, W))
            if arena is None:
                per_call = int(_lib.lib().vmv_groupnorm_scratch_bytes(self.m.dim, H * W, nb)) + 256
                # 166 GroupNorms per forward (per-CTA partial sums) + the LayerNorm partial-statistics slots (99 LayerNorms:
                # 30 at each of the three transformer levels with 4 / 6 / 10 slots per row, 9 in the middle block
                # = 187 x (level-0 rows) x 8 B; sized with margin).  Nothing here is ever memset.
                arena = self._gn_arenas[(nb, H, W)] = ops.GnArena(x32.device, 192 * per_call + 336 * nb * H * W * 8,
                                                                  bar_bytes=192 * ((nb * 8 + 63) // 64 * 64))
            self._gn_arena = arena
            arena.reset()                              # rewind only: every call of this forward takes a fresh region
        sh = self.shard
        if sh is not None:
            sh.begin_forward()
        if sh is not None and sh.frame_sharded:
            # every rank receives the full [B,C,F,h,w] latent (as the sampler holds it) and computes its F/P frames
            sh.check(Fr, (H >> (len(self.m.dim_mult) - 1)) * (W >> (len(self.m.dim_mult) - 1)))
            Fl = Fr // sh.world
            lo = sh.rank * Fl
            x32 = x32[:, :, lo:lo + Fl].contiguous()
            cam = None if cam is None else cam[:, lo:lo + Fl].contiguous()
            concat = None if concat is None else concat[:, :, lo:lo + Fl].contiguous()
            Fr = Fl
        S = {"B": B, "F": Fr, "H": H, "W": W, "x_in": x32}
        if concat is not None:
            S["x_in2"] = concat
        S["emb_all"] = self._embeddings(t, fps, cam, B, Fr)
        S["kv_all"], S["L"] = kv_all, L                 # step-invariant: computed once per sample in prepare_condition
        skips = []
        h = None
        for blk in self.enc:
            h = self._run_block(blk, h, None, S)
            skips.append(h)
        h = self._run_block({"kind": "list", "items": self.mid}, h, None, S)
        for blk in self.dec:
            h = self._run_block(blk, h, skips.pop(), S)
        a = ops.groupnorm(h, *self.head_gn, rows_per_batch=S["H"] * S["W"], eps=1e-5, silu=True, scratch=self._gn_arena)
        if self.head_gemm is not None:
            o16 = self._gemm(a, self.head_gemm, mode=ops.CONV3X3, geom=(1, B * Fr, S["H"], S["W"]))
            out = ops.rows_to_ncfhw(o16, B, Fr, S["H"], S["W"], self.head_cout)
        else:
            out = ops.conv3x3_out(a, self.head_w, self.head_b, B, Fr, S["H"], S["W"])
        return out if sh is None else parallel.gather_output(out, sh)

    # ------------------------------------------------------------------------------------------------------------
    # I2V conditioning glue (step-invariant; depends only on local_image / image / y)
    # ------------------------------------------------------------------------------------------------------------
    def _i2v_condition(self, xshape, y, image, local_image):
        """unet_i2vgen.py:314-346 (concat branch) and :361-382 (context).  Tiny 4..32-channel convolutions and a
        2-head d=4 transformer over frames whose inputs do not change across the 50 DDIM steps: evaluated with torch
        ops as host-side glue (like the VAE/CLIP), not part of the per-step hot path."""
        m = self.m
        B, _, Fr, H, W = xshape
        li = local_image.to(device=self.device, dtype=torch.float32)
        if li.ndim == 5 and li.size(2) > 1:
            li = li[:, :, :1]
        elif li.ndim != 5:
            li = li.unsqueeze(2)
        if Fr > 1:
            pos = torch.cat([torch.ones_like(li[:, :, :1]) * ((tp + 1) / (Fr - 1)) for tp in range(Fr - 1)], dim=2)
            ximg = torch.cat([li[:, :, :1], pos], dim=2)
        else:
            ximg = li
        ximg = ximg.permute(0, 2, 1, 3, 4).reshape(B * ximg.shape[2], -1, H, W)

        def run_seq(seq, z):                       # eval semantics (Dropout = identity), true fp32 convolutions
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                for layer in seq:
                    if not isinstance(layer, torch.nn.Dropout):
                        z = layer(z)
            return z

        ximg = run_seq(m.local_image_concat, ximg)
        cd = ximg.shape[1]
        tok = ximg.reshape(B, Fr, cd, H, W).permute(0, 3, 4, 1, 2).reshape(B * H * W, Fr, cd)
        enc = m.local_temporal_encoder
        for attn, ff in enc.layers:
            n = attn.norm(tok)
            q, k, v = attn.fn.to_qkv(n).chunk(3, dim=-1)
            hd = attn.fn.heads

            def sp(z):
                return z.reshape(z.shape[0], z.shape[1], hd, -1).transpose(1, 2)
            o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(tok.shape[0], Fr, -1)
            if not isinstance(attn.fn.to_out, torch.nn.Identity):
                o = attn.fn.to_out[0](o)
            tok = o + tok
            tok = ff.net[2](ff.net[0](tok)) + tok
        ximg = tok.reshape(B, H, W, Fr, cd).permute(0, 4, 3, 1, 2)
        concat = (ximg + ximg).contiguous()                                   # :345-346 (kept double add)
        ctx = y.to(torch.float32)
        lc = run_seq(m.local_image_embedding, li[:, :, 0])
        ctx = torch.cat([ctx, lc.flatten(2).transpose(1, 2)], dim=1)
        if image is not None:
            ic = run_seq(m.context_embedding, image.to(device=self.device, dtype=torch.float32))
            ctx = torch.cat([ctx, ic.view(-1, m.num_tokens, m.context_dim)], dim=1)
        return concat, ctx


class _Graph:
    """One captured CUDA graph of `_forward_impl` for a fixed set of shapes; inputs are copied into static buffers.
    Every library launch lands on the capturing stream (ops._stream), TMA descriptors are baked into the kernel
    parameters, and all intermediates live in the graph's private memory pool.  Inputs that are the very same tensor as
    at the previous replay (the per-sample conditioning: context K/V, cameras, fps, I2V planes) are not copied again."""

    def __init__(self, eng: UNetEngine, L: int, *inputs):
        self.L = L
        self.static = [None if z is None else z.clone() for z in inputs]
        self.last = [None] * len(inputs)            # (tensor, _version) last copied into each static buffer (strong refs)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm up: func attributes, workspace, allocator
                eng._forward_impl(*self.static, L)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = eng._forward_impl(*self.static, L)
        self.launches = ops.launch_count() - n0     # kernels per replay (bench `gpu_launches`)

    def run(self, *inputs):
        self.replays = getattr(self, "replays", 0) + 1
        for i, (dst, src) in enumerate(zip(self.static, inputs)):
            if dst is None:
                continue
            prev = self.last[i]
            if prev is not None and prev[0] is src and prev[1] == src._version:
                continue
            dst.copy_(src, non_blocking=True)
            self.last[i] = (src, src._version)
        self.graph.replay()
        return self.out.clone()
