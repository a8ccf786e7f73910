"""Parity of the non-GEMM kernels vs plain fp32 torch on identical fp16 inputs (through the C ABI)."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _r(*shape, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale + shift).half()


@pytest.mark.parametrize("nb,rows,C1,C2,eps,silu", [
    (24, 1024, 320, 0, 1e-5, True), (1, 24 * 1024, 320, 0, 1e-6, False), (24, 256, 1280, 640, 1e-5, True),
    (2, 24 * 16, 1280, 0, 1e-5, True), (24, 16, 1280, 1280, 1e-5, True), (4, 256, 64, 0, 1e-5, True),
    (24, 64, 1280, 640, 1e-5, True), (4, 64, 128, 64, 1e-6, False), (48, 1024, 640, 320, 1e-5, True),
    # smem-resident single-pass kernel at its capacity limit (213 / 214 KB per CTA), odd row counts, more chunks than SMs
    (2, 24 * 1024, 320, 0, 1e-5, True), (48, 1024, 320, 0, 1e-5, True), (3, 1000, 320, 0, 1e-5, True),
    (160, 16, 64, 0, 1e-5, True), (48, 16, 1280, 0, 1e-5, True), (2, 24 * 64, 1280, 0, 1e-6, False),
])
@pytest.mark.parametrize("fused", [False, True], ids=["split", "fused"])
def test_groupnorm(nb, rows, C1, C2, eps, silu, fused):
    from videomv_b200 import ops
    x1 = _r(nb * rows, C1, seed=1, shift=0.5, scale=2.0)
    x2 = _r(nb * rows, C2, seed=2, shift=-0.3) if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * torch.randn(C, device="cuda")
    beta = 0.1 * torch.randn(C, device="cuda")
    arena = ops.GnArena("cuda", 1 << 20) if fused else None       # single-launch kernel (stats + barrier + apply)
    out = ops.groupnorm(x1, gamma, beta, rows_per_batch=rows, eps=eps, silu=silu, x2=x2, scratch=arena)
    x = x1.float() if x2 is None else torch.cat([x1, x2], 1).float()
    xi = x.reshape(nb, rows, C).permute(0, 2, 1)
    ref = F.group_norm(xi, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(nb * rows, C)
    assert_close(f"groupnorm nb{nb} rows{rows} C{C1}+{C2}", out, ref)
    if fused:                                                     # second call on a fresh region, in place over x1
        if C2 == 0:
            out2 = ops.groupnorm(x1, gamma, beta, rows_per_batch=rows, eps=eps, silu=silu, out=x1, scratch=arena)
            assert_close(f"groupnorm in-place nb{nb} rows{rows} C{C1}", out2, ref)


def test_groupnorm_strided_views():
    """Column slices of wider tensors as inputs / output (row stride != C): the smem kernel copies row by row."""
    from videomv_b200 import ops
    nb, rows, C1, C2 = 24, 256, 320, 192
    big1, big2 = _r(nb * rows, C1 + 64, seed=1, shift=0.4), _r(nb * rows, C2 + 128, seed=2)
    x1, x2 = big1[:, 64:], big2[:, :C2]
    C = C1 + C2
    gamma, beta = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    big_out = torch.zeros(nb * rows, C + 32, dtype=torch.float16, device="cuda")
    arena = ops.GnArena("cuda", 1 << 20)
    ops.groupnorm(x1, gamma, beta, rows_per_batch=rows, eps=1e-5, silu=True, x2=x2, out=big_out[:, 32:], scratch=arena)
    x = torch.cat([x1, x2], 1).float().reshape(nb, rows, C).permute(0, 2, 1)
    ref = F.silu(F.group_norm(x, 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(nb * rows, C)
    assert_close("groupnorm strided", big_out[:, 32:], ref)
    assert (big_out[:, :32] == 0).all()


@pytest.mark.parametrize("M,C", [(24576, 320), (6144, 640), (1536, 1280), (100, 512), (7, 64), (33, 2048)])
def test_layernorm(M, C):
    from videomv_b200 import ops
    x = _r(M, C, seed=1, shift=0.2, scale=1.5)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda")
    beta = 0.1 * torch.randn(C, device="cuda")
    out = ops.layernorm(x, gamma, beta)
    assert_close(f"layernorm {M}x{C}", out, F.layer_norm(x.float(), (C,), gamma, beta, 1e-5))


def _attn_ref(q, k, v, scale):
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * scale
    return torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, -1), v.float())


@pytest.mark.parametrize("NF,HW,heads", [(24, 1024, 5), (24, 256, 10), (24, 64, 20), (24, 16, 20), (48, 64, 20), (48, 16, 20),
                                         (3, 64, 2), (5, 16, 3), (7, 32, 1), (2, 4096, 5), (3, 100, 2), (3, 300, 2), (5, 128, 1),
                                         (2, 1000, 3)])
@pytest.mark.parametrize("impl", [1, 2], ids=["mma.sync", "tcgen05"])
def test_attention_spatial_fused_qkv(NF, HW, heads, impl):
    from videomv_b200 import ops
    if impl == 2 and HW < 128 and (HW & (HW - 1)):
        pytest.skip("tcgen05 attention packs short sequences only when their length is a power of two")
    C = heads * 64
    qkv = _r(NF * HW, 3 * C, seed=1)
    out = torch.empty(NF * HW, C, dtype=torch.float16, device="cuda")
    ld = 3 * C
    ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], out, outer=NF, inner=1, heads=heads, nq=HW, nk=HW,
                  q_strides=(HW * ld, 0, ld), k_strides=(HW * ld, 0, ld), v_strides=(HW * ld, 0, ld),
                  o_strides=(HW * C, 0, C), impl=impl)
    t = qkv.reshape(NF, HW, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(t[0], t[1], t[2], 0.125).permute(0, 2, 1, 3).reshape(NF * HW, C)
    assert_close(f"attn spatial NF{NF} HW{HW} h{heads}", out, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,Fr,HW,heads,L", [(1, 24, 1024, 5, 77), (2, 24, 64, 20, 77), (2, 24, 16, 20, 77), (1, 24, 16, 20, 145),
                                              (1, 4, 256, 10, 145), (2, 3, 256, 10, 77), (2, 2, 4096, 5, 145)])
@pytest.mark.parametrize("impl", [1, 2], ids=["mma.sync", "tcgen05"])
def test_attention_cross(B, Fr, HW, heads, L, impl):
    from videomv_b200 import ops
    if impl == 2 and HW < 128 and Fr % (128 // HW):
        pytest.skip("packed cross-attention needs the frames of a tile to share one context")
    C = heads * 64
    q = _r(B * Fr * HW, C, seed=1)
    kv = _r(B * L, 2 * C, seed=2)
    out = torch.empty_like(q)
    ops.attention(q, kv, kv[:, C:], out, outer=B * Fr, inner=1, heads=heads, nq=HW, nk=L,
                  q_strides=(HW * C, 0, C), k_strides=(L * 2 * C, 0, 2 * C), v_strides=(L * 2 * C, 0, 2 * C),
                  o_strides=(HW * C, 0, C), kv_group=Fr, impl=impl)
    qh = q.reshape(B * Fr, HW, heads, 64).permute(0, 2, 1, 3)
    kvh = kv.reshape(B, L, 2, heads, 64).permute(2, 0, 3, 1, 4).repeat_interleave(Fr, dim=1)
    ref = _attn_ref(qh, kvh[0], kvh[1], 0.125).permute(0, 2, 1, 3).reshape(B * Fr * HW, C)
    assert_close(f"attn cross B{B} F{Fr} HW{HW} L{L}", out, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,Fr,HW,heads", [(1, 24, 1024, 5), (2, 24, 256, 10), (2, 24, 16, 20), (1, 4, 64, 8), (1, 24, 1024, 8),
                                           (2, 24, 1024, 5), (1, 24, 7, 3), (2, 6, 33, 2), (1, 128, 5, 1), (1, 24, 512, 5)])
@pytest.mark.parametrize("impl", [1, 2], ids=["mma.sync", "tcgen05"])
def test_attention_temporal(B, Fr, HW, heads, impl):
    """Temporal attention: a sequence = the F frames of one pixel (row stride HW * ld).  impl 2 = the tcgen05 kernel with
    G = 128 // F pixels packed per tile by a 4-D TMA box and the j % G == r % G mask."""
    from videomv_b200 import ops
    C = heads * 64
    ld = 3 * C
    qkv = _r(B * Fr * HW, ld, seed=1)
    out = torch.empty(B * Fr * HW, C, dtype=torch.float16, device="cuda")
    st = (Fr * HW * ld, ld, HW * ld)
    ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], out, outer=B, inner=HW, heads=heads, nq=Fr, nk=Fr,
                  q_strides=st, k_strides=st, v_strides=st, o_strides=(Fr * HW * C, C, HW * C), impl=impl)
    t = qkv.reshape(B, Fr, HW, 3, heads, 64).permute(3, 0, 2, 4, 1, 5).reshape(3, B * HW, heads, Fr, 64)
    ref = _attn_ref(t[0], t[1], t[2], 0.125)                              # [B*HW, h, F, 64]
    ref = ref.reshape(B, HW, heads, Fr, 64).permute(0, 3, 1, 2, 4).reshape(B * Fr * HW, C)
    assert_close(f"attn temporal B{B} F{Fr} HW{HW} h{heads}", out, ref, rtol=2e-3, atol=2e-3)


def test_upsample_and_im2col():
    from videomv_b200 import ops
    n, H, W, C = 3, 8, 8, 64
    x = _r(n * H * W, C, seed=1)
    up = ops.upsample_nearest2x(x, n, H, W)
    ref = F.interpolate(x.float().reshape(n, H, W, C).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert torch.equal(up.float(), ref.permute(0, 2, 3, 1).reshape(-1, C))
    col = ops.im2col_3x3_s2(x, n, H, W)
    unf = F.unfold(x.float().reshape(n, H, W, C).permute(0, 3, 1, 2), 3, padding=1, stride=2)   # [n, C*9, L]
    unf = unf.reshape(n, C, 9, -1).permute(0, 3, 2, 1).reshape(n * (H // 2) * (W // 2), 9 * C)
    assert torch.equal(col.float(), unf)


def test_downsample_conv_via_im2col():
    from videomv_b200 import ops, packing
    n, H, W, C = 24, 16, 16, 640
    x = _r(n * H * W, C, seed=1)
    w = torch.randn(C, C, 3, 3, device="cuda") * (9 * C) ** -0.5
    b = torch.randn(C, device="cuda")
    out = ops.gemm(ops.im2col_3x3_s2(x, n, H, W), packing.pack_conv3x3(w), bias=b)
    ref = F.conv2d(x.float().reshape(n, H, W, C).permute(0, 3, 1, 2), w.half().float(), b, stride=2, padding=1)
    assert_close("downsample conv", out, ref.permute(0, 2, 3, 1).reshape(-1, C))


@pytest.mark.parametrize("C2", [0, 4])
def test_conv_in_out(C2):
    from videomv_b200 import ops
    B, Fr, H, W, Cout = 2, 3, 16, 16, 64
    x1 = torch.randn(B, 4, Fr, H, W, device="cuda")
    x2 = torch.randn(B, C2, Fr, H, W, device="cuda") if C2 else None
    w = torch.randn(Cout, 4 + C2, 3, 3, device="cuda") * 0.2
    b = torch.randn(Cout, device="cuda")
    out = ops.conv3x3_in(x1, w, b, x2)
    xin = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.conv2d(xin.permute(0, 2, 1, 3, 4).reshape(B * Fr, -1, H, W), w, b, padding=1)
    assert_close("conv_in", out, ref.permute(0, 2, 3, 1).reshape(-1, Cout))
    # head conv
    C = 320
    xh = _r(B * Fr * H * W, C, seed=3)
    wo = torch.randn(4, C, 3, 3, device="cuda") * (9 * C) ** -0.5
    bo = torch.randn(4, device="cuda")
    o = ops.conv3x3_out(xh, wo, bo, B, Fr, H, W)
    ref = F.conv2d(xh.float().reshape(B * Fr, H, W, C).permute(0, 3, 1, 2), wo, bo, padding=1)
    ref = ref.reshape(B, Fr, 4, H, W).permute(0, 2, 1, 3, 4)
    assert_close("conv_out", o, ref, rtol=1e-4, atol=1e-5)


def test_embeddings_and_ddim():
    from videomv_b200 import ops
    t = torch.tensor([981, 1, 500], device="cuda")
    e = ops.sinusoidal_embedding(t, 320)
    half = 160
    sinus = torch.outer(t.float(), torch.pow(10000, -torch.arange(half, device="cuda").float().div(half)))
    ref = torch.cat([torch.cos(sinus), torch.sin(sinus)], 1)
    assert_close("sinusoidal", e, ref, rtol=1e-3, atol=6e-4)       # fp16 storage of values in [-1,1]
    B, Fr, E = 2, 5, 1280
    te, te2, cam = _r(B, E, seed=1), _r(B, E, seed=2), _r(B * Fr, E, seed=3)
    o = ops.embed_combine_silu(te, te2, cam, B, Fr)
    ref = F.silu(te.float().repeat_interleave(Fr, 0) + te2.float().repeat_interleave(Fr, 0) + cam.float())
    assert_close("embed_combine", o, ref)
    xt, y, u = (torch.randn(1, 4, 24, 32, 32, device="cuda") for _ in range(3))
    a_t, a_prev, gs = 0.37, 0.52, 9.0
    coef = torch.tensor([a_t ** -0.5, (1 / a_t - 1) ** 0.5, a_t ** -0.5, (1 / a_t - 1) ** 0.5, a_prev ** 0.5,
                         (1 - a_prev) ** 0.5, gs], device="cuda")
    xp = ops.cfg_ddim_step(xt, y, u, coef)
    eps = u + gs * (y - u)
    x0 = coef[0] * xt - coef[1] * eps
    ref = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * ((coef[2] * xt - x0) / coef[3])
    assert_close("cfg_ddim", xp, ref, rtol=1e-5, atol=1e-5)
