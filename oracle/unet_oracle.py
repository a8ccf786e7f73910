"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the VideoMV video-UNet forward.

A plain-PyTorch fp32, purely functional restatement of the reference's hot path.
It is driven by a reference-format ``state_dict`` (same key names / shapes as
``UNetSD_T2VBase`` / ``UNetSD_I2VGen``): block types are recognised from the keys,
so the oracle shares no structure code with the product in ``videomv_b200/``.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Pinned against the reference itself by oracle/gen_golden.py
(the reference has no golden vectors of its own: SURVEY.md section 4).

eval-mode semantics only: every nn.Dropout is the identity.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

HEAD_DIM_DEFAULT = 64


# --------------------------------------------------------------------------- #
# leaf arithmetic
# --------------------------------------------------------------------------- #
def sinusoidal_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """tools/modules/unet/util.py:177-189 -- [cos | sin], freq 10000^(-i/half)."""
    half = dim // 2
    timesteps = timesteps.float()
    freqs = torch.pow(10000, -torch.arange(half).to(timesteps).div(half))
    sinusoid = torch.outer(timesteps, freqs)
    x = torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)
    if dim % 2 != 0:
        x = torch.cat([x, torch.zeros_like(x[:, :1])], dim=1)
    return x


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    w = sd[p + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[p + ".bias"], 1e-5)


def attention_core(q, k, v, heads):
    """util.py:237-268 (MemoryEfficientCrossAttention head split + xformers MEA).

    xformers.ops.memory_efficient_attention(q,k,v) == softmax(q k^T / sqrt(d)) v, no
    mask/bias/dropout (xformers==0.0.13, requirements.txt:10); the in-repo pure-torch
    statement of the same arithmetic is CrossAttention.forward util.py:396-427.
    The max_bs chunking (util.py:247-256) is batch-wise and does not change results.
    """
    b, nq, inner = q.shape
    d = inner // heads
    nk = k.shape[1]
    q = q.reshape(b, nq, heads, d).permute(0, 2, 1, 3)
    k = k.reshape(b, nk, heads, d).permute(0, 2, 1, 3)
    v = v.reshape(b, nk, heads, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v)
    return o.permute(0, 2, 1, 3).reshape(b, nq, inner)


def mea(sd, p, x, context, head_dim):
    """util.py:230-268: to_q/to_k/to_v (no bias), attention, to_out.0 (+bias)."""
    q = _lin(sd, p + ".to_q", x)
    ctx = x if context is None else context
    k = _lin(sd, p + ".to_k", ctx)
    v = _lin(sd, p + ".to_v", ctx)
    heads = q.shape[-1] // head_dim
    o = attention_core(q, k, v, heads)
    return _lin(sd, p + ".to_out.0", o)


def feed_forward_geglu(sd, p, x):
    """util.py:543-577: Linear C->8C ; value * gelu_erf(gate) ; Linear 4C->C."""
    h = _lin(sd, p + ".net.0.proj", x)
    val, gate = h.chunk(2, dim=-1)
    h = val * F.gelu(gate)
    return _lin(sd, p + ".net.2", h)


def basic_transformer_block(sd, p, x, context, head_dim):
    """util.py:536-540 (disable_self_attn=False => attn1 is always self-attn)."""
    x = mea(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, head_dim) + x
    x = mea(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, head_dim) + x
    x = feed_forward_geglu(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x


# --------------------------------------------------------------------------- #
# blocks
# --------------------------------------------------------------------------- #
def temporal_conv_block_v2(sd, p, x5):
    """util.py:1381-1392 on [b,c,f,h,w]; GN stats span (c/32,f,h,w); eps 1e-5."""
    identity = x5
    x = x5
    for i, conv_idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = f"{p}.conv{i}"
        x = _gn(sd, q + ".0", x, 1e-5)
        x = F.silu(x)
        x = F.conv3d(x, sd[f"{q}.{conv_idx}.weight"], sd[f"{q}.{conv_idx}.bias"], padding=(1, 0, 0))
    return identity + x


def res_block(sd, p, x, emb, batch):
    """util.py:703-730 (use_scale_shift_norm=False, no up/down, use_temporal_conv)."""
    h = _gn(sd, p + ".in_layers.0", x, 1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = _gn(sd, p + ".out_layers.0", h, 1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    h = x + h
    bf, c, hh, ww = h.shape
    h5 = h.reshape(batch, bf // batch, c, hh, ww).permute(0, 2, 1, 3, 4)
    h5 = temporal_conv_block_v2(sd, p + ".temopral_conv", h5)   # [sic] util.py:691
    return h5.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def spatial_transformer(sd, p, x, context, head_dim):
    """util.py:354-373 with use_linear=True; GN eps 1e-6 (util.py:329)."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = _lin(sd, p + ".proj_in", x)
    x = basic_transformer_block(sd, p + ".transformer_blocks.0", x, context, head_dim)
    x = _lin(sd, p + ".proj_out", x)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


def temporal_transformer(sd, p, x, batch, head_dim):
    """util.py:1043-1089 with only_self_att=True, use_linear=False.

    x arrives as [(b f),c,h,w]; the caller-side rearranges of unet_t2v.py:453-455 are
    folded in.  GN (eps 1e-6, util.py:1014) is over [b,c,f,h,w] => stats span frames.
    proj_in/out are Conv1d k=1 (weights [out,in,1]).  Both attentions are self-attn
    over the f axis (context=None).
    """
    bf, c, h, w = x.shape
    f = bf // batch
    x5 = x.reshape(batch, f, c, h, w).permute(0, 2, 1, 3, 4)           # b c f h w
    x_in = x5
    y = _gn(sd, p + ".norm", x5, 1e-6)
    y = y.permute(0, 3, 4, 2, 1).reshape(batch * h * w, f, c)          # (b h w) f c
    y = F.linear(y, sd[p + ".proj_in.weight"][:, :, 0], sd[p + ".proj_in.bias"])
    y = basic_transformer_block(sd, p + ".transformer_blocks.0", y, None, head_dim)
    y = F.linear(y, sd[p + ".proj_out.weight"][:, :, 0], sd[p + ".proj_out.bias"])
    y = y.reshape(batch, h, w, f, c).permute(0, 4, 3, 1, 2)            # b c f h w
    y = y + x_in
    return y.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def _sub_indices(sd, p):
    idx = set()
    pl = len(p) + 1
    for k in sd:
        if k.startswith(p + "."):
            head = k[pl:].split(".", 1)[0]
            if head.isdigit():
                idx.add(int(head))
    return sorted(idx)


def forward_block(sd, p, x, emb, context, batch, head_dim):
    """unet_t2v.py:436-523 `_forward_single`: type dispatch, here by state_dict keys."""
    if (p + ".op.weight") in sd:                      # Downsample util.py:749 (conv s2 p1)
        return F.conv2d(x, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
    if (p + ".conv.weight") in sd:                    # Upsample util.py:604-606
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    if (p + ".in_layers.0.weight") in sd:
        return res_block(sd, p, x, emb, batch)
    if (p + ".transformer_blocks.0.attn1.to_q.weight") in sd:
        if sd[p + ".proj_in.weight"].ndim == 2:
            return spatial_transformer(sd, p, x, context, head_dim)
        return temporal_transformer(sd, p, x, batch, head_dim)
    if (p + ".weight") in sd and sd[p + ".weight"].ndim == 4:   # plain Conv2d (unet_t2v.py:169)
        return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    subs = _sub_indices(sd, p)                        # nn.ModuleList
    assert subs, f"oracle: unknown block at {p}"
    for j in subs:
        x = forward_block(sd, f"{p}.{j}", x, emb, context, batch, head_dim)
    return x


def _mlp(sd, p, x):
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", x)))


def _trunk(sd, x, emb, context, batch, head_dim):
    """unet_t2v.py:348-368 / unet_i2vgen.py:384-414: encoder, middle, decoder, head."""
    xs = []
    for i in _sub_indices(sd, "input_blocks"):
        x = forward_block(sd, f"input_blocks.{i}", x, emb, context, batch, head_dim)
        xs.append(x)
    for i in _sub_indices(sd, "middle_block"):
        x = forward_block(sd, f"middle_block.{i}", x, emb, context, batch, head_dim)
    for i in _sub_indices(sd, "output_blocks"):
        x = torch.cat([x, xs.pop()], dim=1)
        x = forward_block(sd, f"output_blocks.{i}", x, emb, context, batch, head_dim)
    x = _gn(sd, "out.0", x, 1e-5)
    x = F.silu(x)
    x = F.conv2d(x, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return x


@torch.no_grad()
def unet_t2v_forward(sd, x, t, y, camera_data=None, fps=None, head_dim=HEAD_DIM_DEFAULT):
    """UNetSD_T2VBase.forward, tools/modules/unet/unet_t2v.py:283-403 (autoencoder=None).

    x [B,4,F,h,w] fp32, t [B] int64, y [B,L,context_dim], camera_data [B,F,16].
    """
    batch, c, f, h, w = x.shape
    dim = sd["time_embed.0.weight"].shape[1]
    emb = _mlp(sd, "time_embed", sinusoidal_embedding(t, dim))                       # :326
    if fps is not None and "fps_embedding.0.weight" in sd:                           # :323-324
        emb = emb + _mlp(sd, "fps_embedding", sinusoidal_embedding(fps, dim))
    emb = emb.repeat_interleave(repeats=f, dim=0)                                    # :327
    if camera_data is not None and "camera_embedding.0.weight" in sd:                # :330-335
        emb = emb + _mlp(sd, "camera_embedding", camera_data.reshape(batch * f, -1))
    context = y.repeat_interleave(repeats=f, dim=0)                                  # :339-346
    x = x.permute(0, 2, 1, 3, 4).reshape(batch * f, c, h, w)                         # :348
    x = _trunk(sd, x, emb, context, batch, head_dim)
    return x.reshape(batch, f, -1, h, w).permute(0, 2, 1, 3, 4).contiguous()         # :368


def _transformer_v2(sd, p, x, heads):
    """util.py:1129-1148 TransformerV2 depth 1: PreNorm(Attention) + residual, FF + residual.

    Attention (util.py:1091-1120): fused to_qkv (no bias), scale dim_head^-0.5, to_out Linear.
    FeedForward glu=False: Linear -> GELU -> Linear (util.py:560-577).
    """
    q = f"{p}.layers.0.0"
    xn = _ln(sd, q + ".norm", x)
    qkv = F.linear(xn, sd[q + ".fn.to_qkv.weight"])
    qq, kk, vv = qkv.chunk(3, dim=-1)
    o = attention_core(qq, kk, vv, heads)
    if (q + ".fn.to_out.0.weight") in sd:
        o = _lin(sd, q + ".fn.to_out.0", o)
    x = o + x
    r = f"{p}.layers.0.1"
    hdn = F.gelu(_lin(sd, r + ".net.0.0", x))
    x = _lin(sd, r + ".net.2", hdn) + x
    return x


@torch.no_grad()
def unet_i2v_forward(sd, x, t, y, image, local_image, camera_data=None, fps=None,
                     head_dim=HEAD_DIM_DEFAULT, num_tokens=4):
    """UNetSD_I2VGen.forward, tools/modules/unet/unet_i2vgen.py:287-439 (autoencoder=None)."""
    batch, c, f, h, w = x.shape
    dim = sd["time_embed.0.weight"].shape[1]
    context_dim = sd["context_embedding.2.weight"].shape[0] // num_tokens
    if local_image.ndim == 5 and local_image.size(2) > 1:                            # :314-317
        local_image = local_image[:, :, :1]
    elif local_image.ndim != 5:
        local_image = local_image.unsqueeze(2)
    # [Concat] :330-346
    if f > 1:
        mask_pos = torch.cat([torch.ones_like(local_image[:, :, :1]) * ((tpos + 1) / (f - 1))
                              for tpos in range(f - 1)], dim=2)
        ximg = torch.cat([local_image[:, :, :1], mask_pos], dim=2)
    else:
        ximg = local_image
    ximg = ximg.permute(0, 2, 1, 3, 4).reshape(batch * ximg.shape[2], -1, h, w)
    p = "local_image_concat"
    ximg = F.conv2d(ximg, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1)
    ximg = F.conv2d(F.silu(ximg), sd[p + ".2.weight"], sd[p + ".2.bias"], padding=1)
    ximg = F.conv2d(F.silu(ximg), sd[p + ".4.weight"], sd[p + ".4.bias"], padding=1)
    cd = ximg.shape[1]
    ximg = ximg.reshape(batch, f, cd, h, w).permute(0, 3, 4, 1, 2).reshape(batch * h * w, f, cd)
    ximg = _transformer_v2(sd, "local_temporal_encoder", ximg, heads=2)
    ximg = ximg.reshape(batch, h, w, f, cd).permute(0, 4, 3, 1, 2)
    concat = ximg + ximg                                                             # :345-346 (kept bug)
    # [Embeddings] :349-357
    emb = _mlp(sd, "time_embed", sinusoidal_embedding(t, dim)) + \
        _mlp(sd, "fps_embedding", sinusoidal_embedding(fps, dim))
    emb = emb.repeat_interleave(repeats=f, dim=0)
    if camera_data is not None and "camera_embedding.0.weight" in sd:
        emb = emb + _mlp(sd, "camera_embedding", camera_data.reshape(batch * f, -1))
    # [Context] :361-382
    context = y
    lc = local_image[:, :, 0]
    p = "local_image_embedding"
    lc = F.silu(F.conv2d(lc, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1))
    lc = F.adaptive_avg_pool2d(lc, (32, 32))
    lc = F.silu(F.conv2d(lc, sd[p + ".3.weight"], sd[p + ".3.bias"], stride=2, padding=1))
    lc = F.conv2d(lc, sd[p + ".5.weight"], sd[p + ".5.bias"], stride=2, padding=1)
    lc = lc.flatten(2).transpose(1, 2)
    context = torch.cat([context, lc], dim=1)
    if image is not None:
        ic = _mlp(sd, "context_embedding", image).view(-1, num_tokens, context_dim)
        context = torch.cat([context, ic], dim=1)
    context = context.repeat_interleave(repeats=f, dim=0)
    x = torch.cat([x, concat], dim=1)                                                # :384
    x = x.permute(0, 2, 1, 3, 4).reshape(batch * f, x.shape[1], h, w)
    x = _trunk(sd, x, emb, context, batch, head_dim)
    return x.reshape(batch, f, -1, h, w).permute(0, 2, 1, 3, 4).contiguous()
