"""Drop-in video-UNet modules for VideoMV: `UNetSD_T2VBase` and `UNetSD_I2VGen`.

Boundary: same class names (registered in the reference `MODEL` registry when it is
importable), same constructor kwargs, same `forward` signature, and a parameter tree whose `state_dict()` keys and
shapes equal the reference's (tools/modules/unet/unet_t2v.py:56-265, unet_i2vgen.py:28-285, util.py) so released
checkpoints load with `load_state_dict(strict=False)` and `configs/*_infer.yaml` drop in unchanged.

The modules below only *hold* parameters (standard nn layers are used as typed containers so names, shapes, init and
`.to()` behave as usual); none of their `forward`s run.  `forward()` hands the call to `engine.UNetEngine`, which
executes the whole network with the library's sm_100a kernels in channels-last fp16.  There is no PyTorch or CPU
fallback: without CUDA + the built library the call raises.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .engine import UNetEngine

GN_GROUPS = 32


def _seq(*mods) -> nn.Sequential:
    return nn.Sequential(*mods)


def _mlp(i: int, h: int, o: int) -> nn.Sequential:
    return _seq(nn.Linear(i, h), nn.SiLU(), nn.Linear(h, o))


class _Holder(nn.Module):
    """Parameter container with a `kind` tag the engine dispatches on (never called)."""
    kind = "holder"

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("videomv_b200 parameter holders are not callable; the UNet runs through UNetEngine")


class TemporalConvParams(_Holder):
    """util.py:1347-1379 TemporalConvBlock_v2: 4 x [GN, SiLU, (Dropout,) Conv3d(3,1,1)]."""
    kind = "tconv"

    def __init__(self, c: int):
        super().__init__()
        def stage(first):
            layers = [nn.GroupNorm(GN_GROUPS, c), nn.SiLU()]
            if not first:
                layers.append(nn.Dropout(0.1))
            layers.append(nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0)))
            return _seq(*layers)
        self.conv1, self.conv2, self.conv3, self.conv4 = stage(True), stage(False), stage(False), stage(False)
        nn.init.zeros_(self.conv4[-1].weight)
        nn.init.zeros_(self.conv4[-1].bias)


class ResBlockParams(_Holder):
    """util.py:610-701 ResBlock (use_scale_shift_norm=False, no up/down) + temporal conv tail."""
    kind = "res"

    def __init__(self, cin: int, emb: int, cout: int, dropout: float):
        super().__init__()
        self.channels, self.out_channels = cin, cout
        self.in_layers = _seq(nn.GroupNorm(GN_GROUPS, cin), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = _seq(nn.SiLU(), nn.Linear(emb, cout))
        self.out_layers = _seq(nn.GroupNorm(GN_GROUPS, cout), nn.SiLU(), nn.Dropout(dropout),
                               nn.Conv2d(cout, cout, 3, padding=1))
        for p in self.out_layers[-1].parameters():
            nn.init.zeros_(p)
        self.skip_connection = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)
        self.temopral_conv = TemporalConvParams(cout)          # [sic] util.py:691


class _AttnParams(nn.Module):
    """util.py:212-228 MemoryEfficientCrossAttention parameters."""

    def __init__(self, qdim: int, ctx: Optional[int], heads: int, dh: int):
        super().__init__()
        inner = heads * dh
        ctx = ctx or qdim
        self.heads, self.dim_head = heads, dh
        self.to_q = nn.Linear(qdim, inner, bias=False)
        self.to_k = nn.Linear(ctx, inner, bias=False)
        self.to_v = nn.Linear(ctx, inner, bias=False)
        self.to_out = _seq(nn.Linear(inner, qdim), nn.Dropout(0.0))


class _GEGLUParams(nn.Module):
    def __init__(self, i: int, o: int):
        super().__init__()
        self.proj = nn.Linear(i, o * 2)


class _FFParams(nn.Module):
    """util.py:560-577 FeedForward(glu=True): GEGLU(dim -> 4 dim) ; Dropout ; Linear(4 dim -> dim)."""

    def __init__(self, dim: int):
        super().__init__()
        self.net = _seq(_GEGLUParams(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim))


class _TBlockParams(nn.Module):
    """util.py:510-531 BasicTransformerBlock: attn1 (self), attn2 (ctx or self), GEGLU ff, 3 LayerNorms."""

    def __init__(self, dim: int, heads: int, dh: int, ctx: Optional[int]):
        super().__init__()
        self.attn1 = _AttnParams(dim, None, heads, dh)
        self.ff = _FFParams(dim)
        self.attn2 = _AttnParams(dim, ctx, heads, dh)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformerParams(_Holder):
    """util.py:311-352 SpatialTransformer(use_linear=True, depth=1)."""
    kind = "spatial"

    def __init__(self, c: int, heads: int, dh: int, ctx: int):
        super().__init__()
        inner = heads * dh
        self.in_channels, self.heads = c, heads
        self.norm = nn.GroupNorm(GN_GROUPS, c, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(c, inner)
        self.transformer_blocks = nn.ModuleList([_TBlockParams(inner, heads, dh, ctx)])
        self.proj_out = nn.Linear(c, inner)
        for p in self.proj_out.parameters():
            nn.init.zeros_(p)


class TemporalTransformerParams(_Holder):
    """util.py:992-1041 TemporalTransformer(only_self_att=True, use_linear=False): Conv1d k=1 projections."""
    kind = "temporal"

    def __init__(self, c: int, heads: int, dh: int):
        super().__init__()
        inner = heads * dh
        self.in_channels, self.heads = c, heads
        self.norm = nn.GroupNorm(GN_GROUPS, c, eps=1e-6, affine=True)
        self.proj_in = nn.Conv1d(c, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([_TBlockParams(inner, heads, dh, None)])
        self.proj_out = nn.Conv1d(inner, c, kernel_size=1)
        for p in self.proj_out.parameters():
            nn.init.zeros_(p)


class DownsampleParams(_Holder):
    """util.py:732-756 Downsample(use_conv=True): Conv2d 3x3 stride 2."""
    kind = "down"

    def __init__(self, c: int):
        super().__init__()
        self.op = nn.Conv2d(c, c, 3, stride=2, padding=1)


class UpsampleParams(_Holder):
    """util.py:579-607 Upsample(use_conv=True): nearest x2 then Conv2d 3x3."""
    kind = "up"

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _VideoUNetBase(nn.Module):
    """Trunk shared by both models: encoder / middle / decoder / head parameter tree + engine plumbing."""

    variant = "t2v"

    def _build_trunk(self, first_in: int, dim: int, out_dim: int, dim_mult: List[int], num_heads: int, head_dim: int,
                     num_res_blocks: int, attn_scales, dropout: float, context_dim: int, temporal_attention: bool):
        embed_dim = dim * 4
        widths = [dim * m for m in dim_mult]
        enc = [dim] + widths
        dec = [widths[-1]] + widths[::-1]
        skips: List[int] = []
        scale = 1.0

        def attn_pair(c):
            mods = [SpatialTransformerParams(c, c // head_dim, head_dim, context_dim)]
            if temporal_attention:
                mods.append(TemporalTransformerParams(c, c // head_dim, head_dim))
            return mods

        # encoder (unet_t2v.py:168-206)
        self.input_blocks = nn.ModuleList()
        stem = [nn.Conv2d(first_in, dim, 3, padding=1)]
        if temporal_attention:
            stem.append(TemporalTransformerParams(dim, num_heads, head_dim))    # 8 heads x 64 = 512 inner
        self.input_blocks.append(nn.ModuleList(stem))
        skips.append(dim)
        for i, (cin, cout) in enumerate(zip(enc[:-1], enc[1:])):
            for j in range(num_res_blocks):
                blk = [ResBlockParams(cin, embed_dim, cout, dropout)]
                if scale in attn_scales:
                    blk += attn_pair(cout)
                cin = cout
                self.input_blocks.append(nn.ModuleList(blk))
                skips.append(cout)
                if i != len(dim_mult) - 1 and j == num_res_blocks - 1:
                    self.input_blocks.append(DownsampleParams(cout))
                    skips.append(cout)
                    scale /= 2.0
        # middle (unet_t2v.py:208-227)
        mid = [ResBlockParams(cout, embed_dim, cout, dropout), SpatialTransformerParams(cout, cout // head_dim, head_dim, context_dim)]
        if temporal_attention:
            mid.append(TemporalTransformerParams(cout, cout // head_dim, head_dim))
        mid.append(ResBlockParams(cout, embed_dim, cout, dropout))
        self.middle_block = nn.ModuleList(mid)
        # decoder (unet_t2v.py:229-258); decoder cross-attention hard-codes context_dim=1024 (:237)
        self.output_blocks = nn.ModuleList()
        for i, (cin, cout) in enumerate(zip(dec[:-1], dec[1:])):
            for j in range(num_res_blocks + 1):
                blk = [ResBlockParams(cin + skips.pop(), embed_dim, cout, dropout)]
                if scale in attn_scales:
                    blk.append(SpatialTransformerParams(cout, cout // head_dim, head_dim, 1024))
                    if temporal_attention:
                        blk.append(TemporalTransformerParams(cout, cout // head_dim, head_dim))
                cin = cout
                if i != len(dim_mult) - 1 and j == num_res_blocks:
                    blk.append(UpsampleParams(cout))
                    scale *= 2.0
                self.output_blocks.append(nn.ModuleList(blk))
        # head (unet_t2v.py:261-265)
        self.out = _seq(nn.GroupNorm(GN_GROUPS, cout), nn.SiLU(), nn.Conv2d(cout, out_dim, 3, padding=1))
        nn.init.zeros_(self.out[-1].weight)

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    def _engine(self) -> UNetEngine:
        eng = self.__dict__.get("_eng")
        if eng is None:
            eng = UNetEngine(self)
            self.__dict__["_eng"] = eng
        return eng

    def invalidate_engine(self):
        """Drop packed weights / captured graphs (called automatically after load_state_dict and .to())."""
        self.__dict__["_eng"] = None
        self.__dict__["_pair_cache"] = {}
        self.__dict__["_loop_graphs"] = {}

    def _apply(self, fn, *a, **k):
        self.invalidate_engine()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **k):
        self.invalidate_engine()
        return super().load_state_dict(state_dict, strict=strict, **k)

    def enable_cuda_graphs(self, on: bool = True):
        """Replay the whole forward as one CUDA graph per input-shape set (kills ~1.2k launch overheads per call)."""
        self._engine().use_graphs = bool(on)
        return self

    def set_frame_sharding(self, group=None, enable: bool = True, exchange=None, cfg_split: bool = False):
        """Spread ONE sample over the ranks of `group` (default: WORLD); see videomv_b200/parallel.py.
        cfg_split=False: pure frame sharding, rank r computes frames [r*F/P, (r+1)*F/P) of both halves of a CFG pair.
        cfg_split=True (even P): the first P/2 ranks evaluate the conditional half of `forward_cfg_pair`, the others the
        unconditional half, each group frame-sharding its half over P/2 ranks (P = 2: no frame sharding at all).
        Every rank must make the same calls with the same inputs and receives the full output."""
        from . import parallel
        eng = self._engine()
        # exchange: None (= VMV_SHARD_EXCHANGE, default "peer": one kernel per exchange over NVLink peer memory) | "gather" | "a2a"
        eng.shard = parallel.ShardCtx(group, device=eng.device, exchange=exchange, cfg_split=cfg_split) if enable else None
        eng._graphs.clear()
        self.__dict__["_pair_cache"] = {}
        self.__dict__["_loop_graphs"] = {}
        return self

    def graph_launches(self) -> int:
        """Kernels per graph replay, summed over captured graphs (0 when none)."""
        eng = self.__dict__.get("_eng")
        return 0 if eng is None else sum(g.launches for g in eng._graphs.values())

    def _pair_condition(self, x, kw_cond, kw_uncond):
        """Step-invariant inputs of a CFG pair, cached per (kwargs tensors, latent shape): (kv, cam, fps, concat, split)."""
        eng = self._engine()
        b = x.shape[0]
        key = tuple((k, id(v), v.data_ptr(), v._version, tuple(v.shape)) for kw in (kw_cond, kw_uncond)
                    for k, v in sorted(kw.items()) if torch.is_tensor(v)) + (tuple(x.shape),)
        cache = self.__dict__.setdefault("_pair_cache", {})
        hit = cache.get(key)
        split = eng.shard is not None and eng.shard.cfg_ways == 2      # this rank evaluates ONE half at batch b
        if hit is None:
            def both(name, dtype=None):
                a, c = kw_cond.get(name), kw_uncond.get(name)
                if a is None or c is None:
                    return None
                if split:
                    z = (a if eng.shard.cfg_index == 0 else c).to(eng.device)
                else:
                    z = torch.cat([a.to(eng.device), c.to(eng.device)], dim=0)
                return z if dtype is None else z.to(dtype)
            y2 = both("y")
            if y2 is None:
                raise ValueError("videomv_b200: forward_cfg_pair needs `y` in both kwargs dicts")
            cam2 = both("camera_data", torch.float32) if self.use_camera_condition else None
            fps2 = both("fps", torch.int64) if (self.use_fps_condition or self.variant == "i2v") else None
            img2, loc2 = both("image"), both("local_image")
            kv, concat = eng.prepare_condition(((1 if split else 2) * b,) + tuple(x.shape[1:]), y2, img2, loc2)
            # the entry keeps the caller's tensors alive (their id() is the key: it must not be recycled for another
            # prompt's tensors while the entry exists) as well as the cat'ed copies prepare_condition keys on
            held = [v for kw in (kw_cond, kw_uncond) for v in kw.values() if torch.is_tensor(v)]
            hit = (kv, None if cam2 is None else cam2.contiguous(), None if fps2 is None else fps2.contiguous(), concat,
                   (y2, img2, loc2, held))
            if len(cache) >= 4:
                cache.clear()
                self.__dict__["_loop_graphs"] = {}
            cache[key] = hit
        kv, cam2, fps2, concat, _ = hit
        return kv, cam2, fps2, concat, split, key

    @torch.no_grad()
    def forward_cfg_pair(self, x, t, kw_cond, kw_uncond):
        """Classifier-free-guidance pair (diffusion_ddim.py:149-155) as ONE batch-2B evaluation: rows [0,B) use the
        conditional kwargs, rows [B,2B) the unconditional ones. Per-sample arithmetic is unchanged (all norms and
        attentions are per sample). Returns (y_out, u_out) fp32."""
        eng = self._engine()
        b = x.shape[0]
        kv, cam2, fps2, concat, split, _ = self._pair_condition(x, kw_cond, kw_uncond)
        if split:
            out = eng.forward_core(x.to(device=eng.device, dtype=torch.float32).contiguous(),
                                   t.to(device=eng.device, dtype=torch.int64).contiguous(), kv, cam2, fps2, concat)
            return out[0], out[1]                      # [cfg half, b, C, F, h, w] gathered from both rank groups
        x2 = torch.cat([x, x], dim=0).to(device=eng.device, dtype=torch.float32).contiguous()
        t2 = torch.cat([t, t], dim=0).to(device=eng.device, dtype=torch.int64).contiguous()
        out = eng.forward_core(x2, t2, kv, cam2, fps2, concat)
        if eng.shard is not None:
            out = out[0]
        return out[:b].contiguous(), out[b:].contiguous()

    @torch.no_grad()
    def cfg_sample_loop_graph(self, noise, t_all, coef, kw_cond, kw_uncond):
        """The WHOLE guided DDIM loop (diffusion_ddim.py:247-260: len(t_all) steps of [CFG pair -> fused guidance + DDIM
        update]) captured as ONE CUDA graph and replayed: no per-step host work at all.
        noise [b,C,F,h,w] fp32; t_all int64 [steps] (device); coef fp32 [steps,7] (sampler.step_coefficients).
        The graph is cached per (conditioning tensors, latent shape, step count); a replay costs one memcpy of the noise."""
        from . import ops
        eng = self._engine()
        b = noise.shape[0]
        kv, cam2, fps2, concat, split, ckey = self._pair_condition(noise, kw_cond, kw_uncond)
        graphs = self.__dict__.setdefault("_loop_graphs", {})
        key = (ckey, int(t_all.numel()))
        g = graphs.get(key)
        if g is None:
            kv_all, L = kv
            st_x = noise.detach().to(device=eng.device, dtype=torch.float32).contiguous().clone()
            st_t = t_all.to(device=eng.device, dtype=torch.int64).contiguous().clone()
            st_c = coef.to(device=eng.device, dtype=torch.float32).contiguous().clone()

            nb2 = b if split else 2 * b
            # the UNet inputs of every step live in two persistent buffers (like the static inputs of the per-step graphs)
            st_x2 = torch.empty((nb2,) + tuple(st_x.shape[1:]), dtype=torch.float32, device=eng.device)
            st_t2 = torch.empty((nb2,), dtype=torch.int64, device=eng.device)

            def one_step(xt, i):
                for h in range(nb2 // b):
                    st_x2[h * b:(h + 1) * b].copy_(xt)
                st_t2.copy_(st_t[i].expand(nb2))
                out = eng._forward_impl(st_x2, st_t2, kv_all, cam2, fps2, concat, L)
                if split:
                    y_out, u_out = out[0], out[1]
                else:
                    if eng.shard is not None:
                        out = out[0]
                    y_out, u_out = out[:b], out[b:]
                return ops.cfg_ddim_step(xt, y_out.contiguous(), u_out.contiguous(), st_c[i])

            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):                     # warm up (function attributes, workspaces, allocator)
                    one_step(st_x, 0)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                xt = st_x
                for i in range(int(st_t.numel())):
                    xt = one_step(xt, i)
            # every tensor the graph reads must outlive it: the timestep table and the UNet input buffers too
            g = graphs[key] = (graph, st_x, st_c, xt, ops.launch_count() - n0, (st_t, st_x2, st_t2, kv_all, cam2, fps2, concat))
        graph, st_x, st_c, out = g[:4]
        st_x.copy_(noise, non_blocking=True)
        st_c.copy_(coef, non_blocking=True)
        graph.replay()
        return out.clone()

    def _check_call(self, x, masked, autoencoder, x0):
        assert self.inpainting or masked is None, "inpainting is not supported"
        if autoencoder is not None or x0 is not None:
            raise NotImplementedError(
                "videomv_b200: the LGM 3D-aware refinement tail (reference unet_t2v.py:370-433) is outside the "
                "accelerated hot path (SURVEY.md section 8f row N4); call with autoencoder=None / x0=None")
        if not x.is_cuda:
            raise RuntimeError("videomv_b200: the UNet hot path only runs on a CUDA (sm_100a) device; there is no CPU fallback")


class UNetSD_T2VBase(_VideoUNetBase):
    """Text -> multi-view video UNet. Reference: tools/modules/unet/unet_t2v.py:56 (ctor :57-265, forward :283-403)."""

    variant = "t2v"

    def __init__(self, config=None, in_dim=4, dim=512, y_dim=512, context_dim=512, hist_dim=156, dim_condition=4,
                 out_dim=6, num_tokens=4, dim_mult=(1, 2, 3, 4), num_heads=None, head_dim=64, camera_dim=16,
                 num_res_blocks=3, attn_scales=(1 / 2, 1 / 4, 1 / 8), use_scale_shift_norm=True, dropout=0.1,
                 temporal_attn_times=1, temporal_attention=True, use_checkpoint=False, use_image_dataset=False,
                 use_sim_mask=False, training=True, inpainting=True, use_fps_condition=False,
                 use_camera_condition=False, use_lgm_refine=False, p_all_zero=0.1, p_all_keep=0.1, zero_y=None,
                 adapter_transformer_layers=1, **kwargs):
        super().__init__()
        embed_dim = dim * 4
        num_heads = num_heads if num_heads else dim // 32
        self.zero_y = zero_y
        self.in_dim, self.dim, self.y_dim, self.context_dim = in_dim, dim, y_dim, context_dim
        self.embed_dim, self.out_dim, self.dim_mult = embed_dim, out_dim, list(dim_mult)
        self.num_heads, self.head_dim, self.num_res_blocks = num_heads, head_dim, num_res_blocks
        self.attn_scales = list(attn_scales)
        self.temporal_attention = temporal_attention
        self.use_checkpoint = use_checkpoint            # accepted, irrelevant (no autograd on this path)
        self.inpainting = inpainting
        self.use_fps_condition, self.use_camera_condition = use_fps_condition, use_camera_condition
        self.camera_dim = camera_dim
        # The LGM refine network (core/models.py) is glue outside this path; keep the flag for callers that read it
        # (diffusion_ddim.py:390) but do not construct it here.
        self.use_lgm_refine = use_lgm_refine
        self.time_embed = _mlp(dim, embed_dim, embed_dim)
        if use_camera_condition:
            self.camera_embedding = _mlp(camera_dim, embed_dim, embed_dim)
            nn.init.zeros_(self.camera_embedding[-1].weight)
            nn.init.zeros_(self.camera_embedding[-1].bias)
        if use_fps_condition:
            self.fps_embedding = _mlp(dim, embed_dim, embed_dim)
            nn.init.zeros_(self.fps_embedding[-1].weight)
            nn.init.zeros_(self.fps_embedding[-1].bias)
        self._build_trunk(in_dim, dim, out_dim, list(dim_mult), num_heads, head_dim, num_res_blocks,
                          list(attn_scales), dropout, context_dim, temporal_attention)

    def forward(self, x, t, x0=None, gs_data=None, sqrt_alphas_cumprod=None, sqrt_one_minus_alphas_cumprod=None,
                sqrt_recip_alphas_cumprod=None, sqrt_recipm1_alphas_cumprod=None, autoencoder=None, y=None, fps=None,
                masked=None, camera_data=None, video_mask=None, focus_present_mask=None, prob_focus_present=0.,
                mask_last_frame_num=0, **kwargs):
        """Same call contract as the reference (diffusion_ddim.py:149-155). Returns [B, out_dim, F, h, w] like x."""
        self._check_call(x, masked, autoencoder, x0)
        if y is None:
            if self.zero_y is None:
                raise ValueError("videomv_b200: forward needs `y` (or a `zero_y` given at construction)")
            y = self.zero_y.repeat(x.shape[0], 1, 1)[:, :1, :]          # unet_t2v.py:344
        use_fps = self.use_fps_condition and fps is not None
        use_cam = self.use_camera_condition and camera_data is not None
        return self._engine().forward(x, t, y, camera_data if use_cam else None, fps if use_fps else None)


class UNetSD_I2VGen(_VideoUNetBase):
    """Image -> multi-view video UNet. Reference: tools/modules/unet/unet_i2vgen.py:28 (forward :287-439)."""

    variant = "i2v"

    def __init__(self, config=None, in_dim=7, dim=512, y_dim=512, context_dim=512, hist_dim=156, concat_dim=8,
                 dim_condition=4, out_dim=6, num_tokens=4, dim_mult=(1, 2, 3, 4), num_heads=None, head_dim=64,
                 num_res_blocks=3, attn_scales=(1 / 2, 1 / 4, 1 / 8), use_scale_shift_norm=True, dropout=0.1,
                 temporal_attn_times=1, camera_dim=16, temporal_attention=True, use_checkpoint=False,
                 use_image_dataset=False, use_sim_mask=False, use_camera_condition=False, use_lgm_refine=False,
                 training=True, inpainting=True, p_all_zero=0.1, p_all_keep=0.1, zero_y=None,
                 adapter_transformer_layers=1, **kwargs):
        super().__init__()
        embed_dim = dim * 4
        num_heads = num_heads if num_heads else dim // 32
        self.zero_y = zero_y
        self.in_dim, self.dim, self.y_dim, self.context_dim = in_dim, dim, y_dim, context_dim
        self.num_tokens = num_tokens
        self.embed_dim, self.out_dim, self.dim_mult = embed_dim, out_dim, list(dim_mult)
        self.num_heads, self.head_dim, self.num_res_blocks = num_heads, head_dim, num_res_blocks
        self.attn_scales = list(attn_scales)
        self.temporal_attention = temporal_attention
        self.use_checkpoint = use_checkpoint
        self.inpainting = inpainting
        self.use_fps_condition = True                     # fps embedding is unconditional here (unet_i2vgen.py:349)
        self.use_camera_condition, self.camera_dim = use_camera_condition, camera_dim
        self.use_lgm_refine = use_lgm_refine
        cd = self.concat_dim = in_dim                     # unet_i2vgen.py:93 overwrites concat_dim with in_dim
        self.time_embed = _mlp(dim, embed_dim, embed_dim)
        self.context_embedding = _mlp(y_dim, embed_dim, context_dim * num_tokens)
        if use_camera_condition:
            self.camera_embedding = _mlp(camera_dim, embed_dim, embed_dim)
            nn.init.zeros_(self.camera_embedding[-1].weight)
            nn.init.zeros_(self.camera_embedding[-1].bias)
        self.fps_embedding = _mlp(dim, embed_dim, embed_dim)
        nn.init.zeros_(self.fps_embedding[-1].weight)
        nn.init.zeros_(self.fps_embedding[-1].bias)
        # conditioning glue, step-invariant (unet_i2vgen.py:145-162); executed once per sample by the engine
        self.local_image_concat = _seq(nn.Conv2d(4, cd * 4, 3, padding=1), nn.SiLU(),
                                       nn.Conv2d(cd * 4, cd * 4, 3, stride=1, padding=1), nn.SiLU(),
                                       nn.Conv2d(cd * 4, cd, 3, stride=1, padding=1))
        self.local_temporal_encoder = _LocalTemporalEncoderParams(cd, heads=2, depth=adapter_transformer_layers)
        self.local_image_embedding = _seq(nn.Conv2d(4, cd * 8, 3, padding=1), nn.SiLU(),
                                          nn.AdaptiveAvgPool2d((32, 32)),
                                          nn.Conv2d(cd * 8, cd * 16, 3, stride=2, padding=1), nn.SiLU(),
                                          nn.Conv2d(cd * 16, 1024, 3, stride=2, padding=1))
        self._build_trunk(in_dim + cd, dim, out_dim, list(dim_mult), num_heads, head_dim, num_res_blocks,
                          list(attn_scales), dropout, context_dim, temporal_attention)

    def forward(self, x, t, x0=None, gs_data=None, sqrt_alphas_cumprod=None, sqrt_one_minus_alphas_cumprod=None,
                sqrt_recip_alphas_cumprod=None, sqrt_recipm1_alphas_cumprod=None, autoencoder=None, y=None, image=None,
                local_image=None, camera_data=None, masked=None, fps=None, video_mask=None, focus_present_mask=None,
                prob_focus_present=0., mask_last_frame_num=0, **kwargs):
        self._check_call(x, masked, autoencoder, x0)
        if local_image is None or fps is None:
            raise ValueError("videomv_b200: UNetSD_I2VGen.forward needs `local_image` and `fps` (unet_i2vgen.py:314,349)")
        if y is None:
            if self.zero_y is None:
                raise ValueError("videomv_b200: forward needs `y` (or a `zero_y` given at construction)")
            y = self.zero_y.repeat(x.shape[0], 1, 1)[:, :1, :]
        use_cam = self.use_camera_condition and camera_data is not None
        return self._engine().forward(x, t, y, camera_data if use_cam else None, fps, image=image, local_image=local_image)


class _PreNormAttnParams(nn.Module):
    def __init__(self, dim: int, heads: int, dh: int):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        fn = nn.Module()
        fn.to_qkv = nn.Linear(dim, heads * dh * 3, bias=False)
        fn.to_out = _seq(nn.Linear(heads * dh, dim), nn.Dropout(0.05)) if not (heads == 1 and dh == dim) else nn.Identity()
        fn.heads = heads
        self.fn = fn


class _PlainFFParams(nn.Module):
    """util.py:560-577 FeedForward(glu=False): Linear -> GELU -> Dropout -> Linear."""

    def __init__(self, dim: int, dim_out: int):
        super().__init__()
        self.net = _seq(_seq(nn.Linear(dim, dim * 4), nn.GELU()), nn.Dropout(0.05), nn.Linear(dim * 4, dim_out))


class _LocalTemporalEncoderParams(nn.Module):
    """util.py:1129-1148 TransformerV2(heads=2, dim=dim_head=mlp_dim=concat_dim)."""

    def __init__(self, dim: int, heads: int, depth: int):
        super().__init__()
        self.depth = depth
        self.layers = nn.ModuleList([nn.ModuleList([_PreNormAttnParams(dim, heads, dim), _PlainFFParams(dim, dim)])
                                     for _ in range(depth)])


def register_with_reference() -> bool:
    """Register both classes in the reference's MODEL registry (utils/registry_class.py:16) under the reference
    names, replacing the stock implementations (utils/registry.py:116-119 allows re-registration). Returns False when
    the reference package is not importable (e.g. on the GPU box)."""
    try:
        from utils.registry_class import MODEL  # type: ignore
    except Exception:
        return False
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")               # "already registered ... will be replaced" is the point
        for cls in (UNetSD_T2VBase, UNetSD_I2VGen):
            MODEL.register_class()(cls)
    return True
