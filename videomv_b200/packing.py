"""Weight repacking: reference `state_dict` layouts -> the K-major fp16 operands `vmv_gemm` consumes.

All functions are pure tensor reshapes (exact apart from the fp32->fp16 cast), done once after
`load_state_dict`; the fp32 reference-named parameters stay the module's source of truth.
"""
from __future__ import annotations

import torch


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """nn.Linear [N,K], nn.Conv1d k=1 [N,K,1], nn.Conv2d 1x1 [N,K,1,1] -> fp16 [N,K]."""
    return w.reshape(w.shape[0], w.shape[1]).to(torch.float16).contiguous()


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d [Cout,Cin,3,3] -> fp16 [Cout, 9*Cin] with K ordered (ky, kx, c)."""
    co, ci = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).to(torch.float16).contiguous()


def pack_upconv3x3(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d [Cout,Cin,3,3] applied AFTER a nearest x2 upsample (util.py:604-606) -> fp16 [4*Cout, 4*Cin]: one 2x2 conv per
    output phase (py, px) on the ORIGINAL image.  Output pixel (2y+py, 2x+px), 3x3 tap k reads upsampled row 2y+py+k-1, i.e.
    input row y + floor((py+k-1)/2): for py = 0 the taps {0 | 1,2} fall on input rows {y-1 | y}, for py = 1 the taps {0,1 | 2}
    on {y | y+1} (same along x), so taps that share an input pixel are summed (in fp32) once here.  Row = phase*Cout + co with
    phase = 2*py + px; K ordered (ty, tx, c) with input offset (ty - 1 + py, tx - 1 + px)."""
    co, ci = w.shape[:2]
    w = w.detach().float()
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}           # phase -> (taps of ty/tx = 0, taps of ty/tx = 1)
    out = torch.empty((4, co, 2, 2, ci), dtype=torch.float32, device=w.device)
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = torch.zeros((co, ci), dtype=torch.float32, device=w.device)
                    for ky in groups[py][ty]:
                        for kx in groups[px][tx]:
                            acc += w[:, :, ky, kx]
                    out[2 * py + px, :, ty, tx] = acc
    return out.reshape(4 * co, 4 * ci).to(torch.float16).contiguous()


def pack_tconv3(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv3d [Cout,Cin,3,1,1] -> fp16 [Cout, 3*Cin] with K ordered (kt, c)."""
    co, ci = w.shape[:2]
    return w[:, :, :, 0, 0].permute(0, 2, 1).reshape(co, 3 * ci).to(torch.float16).contiguous()


def geglu_block_n(n: int) -> int:
    """N tile used for a GEGLU GEMM with N = 8C columns: must divide N and BN/2 must be a multiple of the 32-column
    epilogue block of the CTA-pair kernel (mirrors pick_block_n in gemm_tc.cu)."""
    return 256 if n % 256 == 0 else 128


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU.proj (util.py:546): W [8C,C] = [value rows | gate rows], bias [8C].

    Re-order rows so every BN-row tile holds BN/2 value rows followed by the BN/2 matching gate rows; the GEMM
    epilogue then computes value*gelu(gate) inside one tile.  Returns (fp16 W, fp32 bias, block_n).
    """
    n = w.shape[0]
    half = n // 2
    bn = geglu_block_n(n)
    hb = bn // 2
    assert n % bn == 0 and half % hb == 0
    idx = []
    for t in range(n // bn):
        idx.extend(range(t * hb, (t + 1) * hb))
        idx.extend(range(half + t * hb, half + (t + 1) * hb))
    idx = torch.tensor(idx, dtype=torch.long, device=w.device)
    return (w.index_select(0, idx).to(torch.float16).contiguous(),
            b.index_select(0, idx).to(torch.float32).contiguous(), bn)


def pack_cat(*ws: torch.Tensor) -> torch.Tensor:
    """Row-concatenate several [Ni,K] projections (fused QKV / KV)."""
    return torch.cat([pack_linear(w) for w in ws], dim=0).contiguous()


def fold_layernorm(w: torch.Tensor, b, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm(gamma, beta) followed by Linear(w, b)  ==  rstd*(x @ (w*gamma)^T - mean*colsum) + (b + w @ beta).
    Returns (w*gamma fp32 [N,K], fp32 bias b + w@beta). The caller packs the weight and takes colsum of the PACKED fp16
    values (so the mean correction cancels exactly what the tensor cores accumulated)."""
    w32 = w.detach().reshape(w.shape[0], -1).float()
    bias = w32 @ beta.detach().float()
    if b is not None:
        bias = bias + b.detach().float()
    return w32 * gamma.detach().float()[None, :], bias.contiguous()
