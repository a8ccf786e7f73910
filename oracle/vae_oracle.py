"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference VAE decoder.

Plain-PyTorch fp32 functional form of `AutoencoderKL.decode` (tools/modules/autoencoder.py:101-104) =
`post_quant_conv` + `Decoder.forward` (:654-691) with `ResnetBlock` (:316-336, temb = None), `AttnBlock` (:419-443),
`Upsample` (:456-460) and `Normalize` = GroupNorm(32, eps 1e-6) (:15-16), driven by a reference-format state_dict.
Pinned against the unmodified reference class by oracle/gen_golden.py (`vae` cases).  Only tests / smoke / the bench's
baseline legs may import this module; the product (videomv_b200/vae.py) never does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _conv(sd, p, x, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def _swish(x):
    return x * torch.sigmoid(x)                                          # autoencoder.py:11-13


def _resnet(sd, p, x):
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)     # dropout = identity in eval
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    return x + h


def _attn(sd, p, x):
    b, c, h, w = x.shape
    n = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{name}", n).reshape(b, c, h * w) for name in ("q", "k", "v"))
    w_ = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** (-0.5)), dim=2)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", h_)


@torch.no_grad()
def vae_decode(sd, z: torch.Tensor) -> torch.Tensor:
    """sd: reference AutoencoderKL state_dict (fp32); z [N, embed_dim, h, w] -> [N, 3, 8h, 8w]."""
    z = _conv(sd, "post_quant_conv", z)
    h = _conv(sd, "decoder.conv_in", z, 1)
    h = _resnet(sd, "decoder.mid.block_1", h)
    h = _attn(sd, "decoder.mid.attn_1", h)
    h = _resnet(sd, "decoder.mid.block_2", h)
    levels = sorted({int(k.split(".")[2]) for k in sd if k.startswith("decoder.up.")})
    for i in reversed(levels):
        j = 0
        while f"decoder.up.{i}.block.{j}.norm1.weight" in sd:
            h = _resnet(sd, f"decoder.up.{i}.block.{j}", h)
            if f"decoder.up.{i}.attn.{j}.norm.weight" in sd:
                h = _attn(sd, f"decoder.up.{i}.attn.{j}", h)
            j += 1
        if f"decoder.up.{i}.upsample.conv.weight" in sd:
            h = _conv(sd, f"decoder.up.{i}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), 1)
    return _conv(sd, "decoder.conv_out", _swish(_gn(sd, "decoder.norm_out", h)), 1)
