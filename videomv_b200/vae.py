"""AutoencoderKL with a B200-native DECODER.

Drop-in for the reference `tools/modules/autoencoder.py:32` (`AUTO_ENCODER` registry): same constructor
(`ddconfig`, `embed_dim`, ...), same `state_dict()` keys / shapes / order, same `decode(z)` contract
(z [N, 4, h, w] already divided by scale_factor -> images [N, 3, 8h, 8w], inference_text2video_entrance.py:279-290).

`decode` runs the reference's `Decoder.forward` (autoencoder.py:654-691) on the library's sm_100a kernels, channels-last fp16:
  conv_in            vmv_conv3x3_in        (z_channels -> 512, CUDA cores: K = 36)
  ResnetBlock        vmv_groupnorm_fused(+SiLU) -> vmv_gemm CONV3X3 -> GN+SiLU -> CONV3X3 (+ residual / 1x1 nin_shortcut)
  mid.attn_1         GroupNorm -> fused q|k projection and V^T projection (vmv_gemm) -> QK^T (vmv_gemm) ->
                     vmv_softmax_rows -> PV (vmv_gemm, + v bias as a column bias: softmax rows sum to 1) -> proj_out + x
  Upsample           vmv_gemm UPCONV3X3 (nearest x2 + 3x3 conv as four phase convs: no 4x tensor)
  norm_out + conv_out  GN+SiLU -> CONV3X3 padded to 16 output columns -> vmv_rows_to_ncfhw
The 4 -> 4 channel 1x1 `post_quant_conv` (16 MACs per pixel) is evaluated with a torch op on the way in.
`encode` is not part of the accelerated path: the encoder parameters are held (checkpoints load with strict=True) and
evaluated with plain torch ops, exactly as the reference does.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, packing


def _norm(c: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)                  # autoencoder.py:15-16


class _ResnetBlockParams(nn.Module):
    """autoencoder.py:277-336 (temb_channels = 0)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.norm1 = _norm(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = _norm(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):                                              # encoder only (torch ops)
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.nin_shortcut(x) if self.in_channels != self.out_channels else x) + h


class _AttnBlockParams(nn.Module):
    """autoencoder.py:392-443."""

    def __init__(self, c: int):
        super().__init__()
        self.in_channels = c
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):                                              # encoder only (torch ops)
        b, c, h, w = x.shape
        n = self.norm(x)
        q, k, v = (m(n).reshape(b, c, h * w) for m in (self.q, self.k, self.v))
        p = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * c ** -0.5, dim=2)
        return x + self.proj_out(torch.bmm(v, p.permute(0, 2, 1)).reshape(b, c, h, w))


class _UpsampleParams(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.with_conv = True
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _DownsampleParams(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.with_conv = True
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):                                              # autoencoder.py:475-482
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class _Encoder(nn.Module):
    """autoencoder.py:484-579; evaluated with torch ops (not on the accelerated path)."""

    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 double_z=True, **ignore):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i], ch * ch_mult[i]
            for _ in range(num_res_blocks):
                block.append(_ResnetBlockParams(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(_AttnBlockParams(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i != self.num_resolutions - 1:
                down.downsample = _DownsampleParams(block_in)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = _ResnetBlockParams(block_in, block_in)
        self.mid.attn_1 = _AttnBlockParams(block_in)
        self.mid.block_2 = _ResnetBlockParams(block_in, block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for i in range(self.num_resolutions):
            for j in range(self.num_res_blocks):
                h = self.down[i].block[j](h)
                if len(self.down[i].attn) > 0:
                    h = self.down[i].attn[j](h)
            if i != self.num_resolutions - 1:
                h = self.down[i].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return self.conv_out(F.silu(self.norm_out(h)))


class _DecoderParams(nn.Module):
    """Parameter container with the reference Decoder's names (autoencoder.py:582-652); executed by `_DecoderEngine`."""

    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels, **ignore):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = _ResnetBlockParams(block_in, block_in)
        self.mid.attn_1 = _AttnBlockParams(block_in)
        self.mid.block_2 = _ResnetBlockParams(block_in, block_in)
        self.up = nn.ModuleList()
        for i in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i]
            for _ in range(num_res_blocks + 1):
                block.append(_ResnetBlockParams(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(_AttnBlockParams(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i != 0:
                up.upsample = _UpsampleParams(block_in)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)


def _f32(p):
    return p.detach().float().contiguous()


class _DecoderEngine:
    """Packed weights + the kernel sequence of Decoder.forward (autoencoder.py:654-691)."""

    def __init__(self, dec: _DecoderParams):
        self.dev = dec.conv_in.weight.device
        if self.dev.type != "cuda":
            raise RuntimeError("videomv_b200: move the autoencoder to a CUDA device before decode() (no CPU fallback)")
        self.cin_w, self.cin_b = _f32(dec.conv_in.weight), _f32(dec.conv_in.bias)
        self.mid1, self.mid2 = self._res(dec.mid.block_1), self._res(dec.mid.block_2)
        self.attn = self._attn(dec.mid.attn_1)
        self.levels: List[Dict] = []
        for i in reversed(range(dec.num_resolutions)):
            up = dec.up[i]
            lv = {"blocks": [self._res(b) for b in up.block], "attn": [self._attn(a) for a in up.attn], "up": None}
            if i != 0:
                lv["up"] = (packing.pack_upconv3x3(up.upsample.conv.weight), _f32(up.upsample.conv.bias))
            self.levels.append(lv)
        self.gn_out = (_f32(dec.norm_out.weight), _f32(dec.norm_out.bias))
        w = dec.conv_out.weight.detach()
        self.out_ch = w.shape[0]
        if self.out_ch > 16 or w.shape[1] % 64:
            raise ValueError("videomv_b200 vae: conv_out needs <= 16 output channels and a multiple of 64 input channels")
        wp = torch.zeros((16, w.shape[1], 3, 3), dtype=w.dtype, device=w.device)
        wp[: self.out_ch] = w
        bp = torch.zeros(16, dtype=torch.float32, device=w.device)
        bp[: self.out_ch] = dec.conv_out.bias.detach().float()
        self.cout_w, self.cout_b = packing.pack_conv3x3(wp), bp
        self.arena: Optional[ops.GnArena] = None

    @staticmethod
    def _res(b: _ResnetBlockParams) -> Dict:
        if b.in_channels % 64 or b.out_channels % 64:
            raise ValueError("videomv_b200 vae: channel counts must be multiples of 64 (ch = 128 in every shipped config)")
        d = {"gn1": (_f32(b.norm1.weight), _f32(b.norm1.bias)), "gn2": (_f32(b.norm2.weight), _f32(b.norm2.bias)),
             "c1": (packing.pack_conv3x3(b.conv1.weight), _f32(b.conv1.bias)),
             "c2": (packing.pack_conv3x3(b.conv2.weight), _f32(b.conv2.bias)), "skip": None}
        if b.in_channels != b.out_channels:
            d["skip"] = (packing.pack_linear(b.nin_shortcut.weight.detach()), _f32(b.nin_shortcut.bias))
        return d

    @staticmethod
    def _attn(a: _AttnBlockParams) -> Dict:
        c = a.in_channels
        lin = lambda m: m.weight.detach().reshape(c, c)
        return {"c": c, "gn": (_f32(a.norm.weight), _f32(a.norm.bias)),
                "qk": (torch.cat([lin(a.q), lin(a.k)], 0).to(torch.float16).contiguous(),
                       torch.cat([_f32(a.q.bias), _f32(a.k.bias)])),
                "v_w": lin(a.v).to(torch.float16).contiguous(), "v_b": _f32(a.v.bias),
                "o": (lin(a.proj_out).to(torch.float16).contiguous(), _f32(a.proj_out.bias))}

    # ---- blocks -------------------------------------------------------------------------------------------------
    def _gn(self, x, gn, hw, silu=True):
        return ops.groupnorm(x, *gn, rows_per_batch=hw, eps=1e-6, silu=silu, scratch=self.arena)

    def _run_res(self, d, x, n, H, W):
        h = ops.gemm(self._gn(x, d["gn1"], H * W), d["c1"][0], bias=d["c1"][1], mode=ops.CONV3X3, geom=(1, n, H, W), w_static=True)
        res = x if d["skip"] is None else ops.gemm(x, d["skip"][0], bias=d["skip"][1], w_static=True)
        return ops.gemm(self._gn(h, d["gn2"], H * W), d["c2"][0], bias=d["c2"][1], mode=ops.CONV3X3, geom=(1, n, H, W),
                        residual=res, w_static=True)

    def _run_attn(self, d, x, n, H, W):
        """Single-head attention over the H*W tokens of each image, head dim = C (autoencoder.py:419-441)."""
        c, hw = d["c"], H * W
        a = self._gn(x, d["gn"], hw, silu=False)
        qk = ops.gemm(a, d["qk"][0], bias=d["qk"][1], w_static=True)                    # [n*hw, 2c] = q | k
        out = torch.empty((n * hw, c), dtype=torch.float16, device=x.device)
        for i in range(n):
            ai, qi, ki = a[i * hw:(i + 1) * hw], qk[i * hw:(i + 1) * hw, :c], qk[i * hw:(i + 1) * hw, c:]
            s = ops.gemm(qi, ki)                                                          # [hw, hw] = q k^T
            p = ops.softmax_rows(s, c ** -0.5)
            vt = ops.gemm(d["v_w"], ai)                                                   # [c, hw] = (W_v a^T): V^T without a transpose
            # P (V + 1 b_v^T) = P V + b_v^T because the rows of P sum to one: the v bias becomes a column bias here
            ops.gemm(p, vt, bias=d["v_b"], out=out[i * hw:(i + 1) * hw])
        return ops.gemm(out, d["o"][0], bias=d["o"][1], residual=x, w_static=True)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z fp32 [N, zc, h, w] (after post_quant_conv) -> fp32 [N, out_ch, 8h, 8w]."""
        n, zc, H, W = z.shape
        if self.arena is None:
            self.arena = ops.GnArena(self.dev, 64 << 20)
        self.arena.reset()
        x5 = z.reshape(1, n, zc, H, W).permute(0, 2, 1, 3, 4).contiguous()              # [1, zc, N, h, w]: images as "frames"
        h = ops.conv3x3_in(x5, self.cin_w, self.cin_b)
        h = self._run_res(self.mid1, h, n, H, W)
        h = self._run_attn(self.attn, h, n, H, W)
        h = self._run_res(self.mid2, h, n, H, W)
        for lv in self.levels:
            for j, blk in enumerate(lv["blocks"]):
                h = self._run_res(blk, h, n, H, W)
                if lv["attn"]:
                    h = self._run_attn(lv["attn"][j], h, n, H, W)
            if lv["up"] is not None:
                h = ops.gemm(h, lv["up"][0], bias=lv["up"][1], mode=ops.UPCONV3X3, geom=(1, n, H, W), w_static=True)
                H, W = 2 * H, 2 * W
        o16 = ops.gemm(self._gn(h, self.gn_out, H * W), self.cout_w, bias=self.cout_b, mode=ops.CONV3X3, geom=(1, n, H, W),
                       block_n=128, w_static=True)
        return ops.rows_to_ncfhw(o16, n, 1, H, W, self.out_ch).reshape(n, self.out_ch, H, W)


class AutoencoderKL(nn.Module):
    """Reference: tools/modules/autoencoder.py:32-105.  `decode` is native; `encode` is plain torch (outside the path)."""

    def __init__(self, ddconfig, embed_dim, pretrained=None, ignore_keys=(), image_key="image", colorize_nlabels=None,
                 monitor=None, ema_decay=None, learn_logvar=False, use_vid_decoder=False, **kwargs):
        super().__init__()
        assert ddconfig["double_z"]
        self.learn_logvar, self.image_key, self.embed_dim = learn_logvar, image_key, embed_dim
        self.encoder = _Encoder(**ddconfig)
        self.decoder = _DecoderParams(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if pretrained is not None:
            self.init_from_ckpt(pretrained, ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=()):
        """autoencoder.py:65-74: keys under `first_stage_model.` of a Stable-Diffusion style checkpoint."""
        sd = torch.load(path, map_location="cpu")["state_dict"]
        self.load_state_dict({k.split("first_stage_model.")[-1]: v for k, v in sd.items() if "first_stage_model" in k}, strict=True)

    def _apply(self, fn, *a, **k):
        self.__dict__["_dec_eng"] = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **k):
        self.__dict__["_dec_eng"] = None
        return super().load_state_dict(state_dict, strict=strict, **k)

    @torch.no_grad()
    def decode(self, z, **kwargs):
        """autoencoder.py:101-104.  z [N, embed_dim, h, w] -> [N, out_ch, 8h, 8w], dtype / device of z."""
        if not z.is_cuda:
            raise RuntimeError("videomv_b200: AutoencoderKL.decode only runs on a CUDA (sm_100a) device; there is no CPU fallback")
        eng = self.__dict__.get("_dec_eng")
        if eng is None:
            eng = self.__dict__["_dec_eng"] = _DecoderEngine(self.decoder)
        zq = F.conv2d(z.float(), self.post_quant_conv.weight.float(), self.post_quant_conv.bias.float())
        out = eng.decode(zq.contiguous())
        return out if z.dtype == torch.float32 else out.to(z.dtype)

    def encode(self, x):
        """Posterior moments [N, 2*embed_dim, h, w] (mean | logvar), autoencoder.py:80-84; torch ops."""
        return self.quant_conv(self.encoder(x))

    def encode_firsr_stage(self, x, scale_factor=1.0):                  # sic (autoencoder.py:86)
        mean, logvar = self.encode(x).chunk(2, dim=1)
        std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
        return scale_factor * (mean + std * torch.randn_like(mean))


def register_with_reference() -> bool:
    """Register under the reference name in its AUTO_ENCODER registry (utils/registry_class.py)."""
    try:
        from utils.registry_class import AUTO_ENCODER  # type: ignore
    except Exception:
        return False
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        AUTO_ENCODER.register_class()(AutoencoderKL)
    return True
