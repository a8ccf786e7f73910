"""tcgen05 GEMM / implicit conv parity vs plain fp32 torch on identical fp16 inputs (through the C ABI)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2], ids=["cta1", "cta2-persistent"])
def variant(request):
    """1 = one tile per CTA (cta_group::1); 2 = persistent CTA pairs (cta_group::2, double-buffered TMEM)."""
    return request.param


def _r(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).half()


@pytest.mark.parametrize("M,N,K,bn,stages", [
    (128, 128, 64, 128, 3), (128, 160, 64, 160, 3), (256, 320, 320, 0, 0), (24576, 320, 320, 0, 0),
    (384, 1280, 1280, 0, 0), (1000, 512, 320, 0, 0), (24, 1280, 320, 0, 0), (1536, 1280, 1280, 160, 6),
    (300, 64, 128, 64, 0), (512, 256, 2560, 256, 0), (6144, 640, 640, 128, 6),
    (49152, 320, 320, 0, 0), (49152, 2560, 320, 0, 0), (385, 480, 1280, 0, 0), (129, 160, 4096, 0, 0),
    (777, 48, 192, 0, 0), (2048, 80, 64, 0, 0), (1536, 1280, 640, 128, 0),
    (12288, 640, 640, 0, 0), (12288, 1920, 640, 0, 0), (49152, 960, 320, 0, 0), (3072, 3840, 1280, 0, 0), (600, 672, 128, 0, 0),
])
def test_linear(M, N, K, bn, stages, variant):
    from videomv_b200 import ops
    a, w = _r(M, K, seed=1), _r(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    out = ops.gemm(a, w, bias=bias, block_n=bn, stages=stages, variant=variant)
    ref = a.float() @ w.float().t() + bias
    assert_close(f"linear M{M} N{N} K{K} bn{bn}", out, ref)


def test_linear_dual_source_residual_rowbias(variant):
    from videomv_b200 import ops
    M, N, K1, K2, rpg = 6144, 640, 1280, 640, 256
    a1, a2 = _r(M, K1, seed=1), _r(M, K2, seed=2)
    w = _r(N, K1 + K2, scale=(K1 + K2) ** -0.5, seed=3)
    bias = torch.randn(N, device="cuda")
    res = _r(M, N, seed=4)
    rb = _r(M // rpg, N, seed=5)
    out = ops.gemm(a1, w, a2=a2, bias=bias, residual=res, rowbias=rb, rows_per_group=rpg, variant=variant)
    ref = torch.cat([a1, a2], 1).float() @ w.float().t() + bias + rb.float().repeat_interleave(rpg, 0) + res.float()
    assert_close("linear dual+res+rowbias", out, ref)


def test_linear_silu_and_strided_views(variant):
    from videomv_b200 import ops
    M, N, K = 500, 320, 192
    big_a = _r(M, 3 * K, seed=1)
    a = big_a[:, K:2 * K]                      # strided view (lda = 3K)
    w = _r(N, K, scale=K ** -0.5, seed=2)
    big_out = torch.zeros(M, 2 * N, dtype=torch.float16, device="cuda")
    ops.gemm(a, w, out=big_out[:, N:], act=ops.ACT_SILU, variant=variant)
    ref = F.silu(a.float() @ w.float().t())
    assert_close("linear silu strided", big_out[:, N:], ref)
    assert (big_out[:, :N] == 0).all()


@pytest.mark.parametrize("C", [320, 512])
def test_geglu(C, variant):
    from videomv_b200 import ops, packing
    M = 1024
    a = _r(M, C, seed=1)
    w = torch.randn(8 * C, C, device="cuda") * C ** -0.5
    b = torch.randn(8 * C, device="cuda")
    wp, bp, bn = packing.pack_geglu(w, b)
    out = ops.gemm(a, wp, bias=bp, act=ops.ACT_GEGLU, block_n=bn, variant=variant)
    h = a.float() @ w.half().float().t() + b
    val, gate = h.chunk(2, dim=-1)
    assert_close(f"geglu C{C}", out, val * F.gelu(gate))


@pytest.mark.parametrize("NF,H,W,Cin,Cout", [
    (4, 32, 32, 64, 64), (24, 32, 32, 320, 320), (24, 16, 16, 640, 640), (24, 8, 8, 1280, 1280),
    (24, 4, 4, 1280, 1280), (4, 4, 4, 128, 64), (3, 8, 8, 64, 128), (2, 64, 64, 64, 64), (5, 2, 2, 64, 64),
])
def test_conv3x3(NF, H, W, Cin, Cout, variant):
    from videomv_b200 import ops, packing
    x = _r(NF * H * W, Cin, seed=1)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (9 * Cin) ** -0.5
    b = torch.randn(Cout, device="cuda")
    rb = _r(NF, Cout, seed=3)
    out = ops.gemm(x, packing.pack_conv3x3(w), bias=b, mode=ops.CONV3X3, geom=(1, NF, H, W), rowbias=rb,
                   rows_per_group=H * W, variant=variant)
    xi = x.float().reshape(NF, H, W, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xi, w.half().float(), b, padding=1) + rb.float()[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(NF * H * W, Cout)
    assert_close(f"conv3x3 {NF}x{H}x{W} {Cin}->{Cout}", out, ref)


@pytest.mark.parametrize("NF,H,W,Cin,Cout", [
    (48, 32, 32, 320, 320), (24, 16, 16, 640, 640), (24, 8, 8, 1280, 1280), (4, 64, 64, 320, 320), (3, 16, 16, 64, 128),
    (5, 4, 4, 64, 64), (2, 8, 8, 128, 64),
])
def test_conv3x3_stride2_without_patch_gather(NF, H, W, Cin, Cout):
    """Downsample.op (util.py:749): Conv2d 3x3 stride 2 pad 1, the windows read by element-strided TMA boxes."""
    from videomv_b200 import ops, packing
    x = _r(NF * H * W, Cin, seed=1)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (9 * Cin) ** -0.5
    b = torch.randn(Cout, device="cuda")
    out = ops.gemm(x, packing.pack_conv3x3(w), bias=b, mode=ops.CONV3X3_S2, geom=(1, NF, H, W))
    assert out.shape == (NF * (H // 2) * (W // 2), Cout)
    xi = x.float().reshape(NF, H, W, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xi, w.half().float(), b, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert_close(f"conv3x3 stride 2 {NF}x{H}x{W} {Cin}->{Cout}", out, ref)
    # same result as the explicit patch gather it replaces
    old = ops.gemm(ops.im2col_3x3_s2(x, NF, H, W), packing.pack_conv3x3(w), bias=b)
    assert_close("vs im2col path", out, old.float(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("NF,H,W,Cin,Cout", [
    (48, 16, 16, 640, 640), (24, 8, 8, 1280, 1280), (24, 4, 4, 1280, 1280), (4, 32, 32, 640, 640), (3, 8, 8, 64, 128),
    (5, 2, 2, 64, 64), (2, 16, 16, 128, 64),
])
def test_upsample_conv3x3_as_four_phase_convs(NF, H, W, Cin, Cout):
    """Upsample (util.py:604-606): nearest x2 then Conv2d 3x3, computed as four 2x2 convs on the original image with the
    taps that fall on the same input pixel pre-summed -- no 4x tensor, 2.25x fewer FLOPs."""
    from videomv_b200 import ops, packing
    x = _r(NF * H * W, Cin, seed=1)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (9 * Cin) ** -0.5
    b = torch.randn(Cout, device="cuda")
    out = ops.gemm(x, packing.pack_upconv3x3(w), bias=b, mode=ops.UPCONV3X3, geom=(1, NF, H, W))
    assert out.shape == (NF * 4 * H * W, Cout)
    xi = x.float().reshape(NF, H, W, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(F.interpolate(xi, scale_factor=2.0, mode="nearest"), w, b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    # the phase weights are sums of up to four fp16-representable taps rounded once: same error class as rounding each tap
    assert_close(f"upsample+conv3x3 {NF}x{H}x{W} {Cin}->{Cout}", out, ref, rtol=2e-3, atol=2e-3)
    if H * 2 <= 64:
        old = ops.gemm(ops.upsample_nearest2x(x, NF, H, W), packing.pack_conv3x3(w), bias=b, mode=ops.CONV3X3, geom=(1, NF, 2 * H, 2 * W))
        assert_close("vs materialised 4x tensor", out, old.float(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,Fr,H,W,C", [
    (1, 24, 32, 32, 320), (2, 24, 16, 16, 640), (1, 24, 8, 8, 1280), (2, 24, 4, 4, 1280), (1, 4, 4, 4, 64),
    (2, 5, 2, 2, 64), (1, 4, 16, 16, 128),
])
def test_tconv3(B, Fr, H, W, C, variant):
    from videomv_b200 import ops, packing
    M = B * Fr * H * W
    x = _r(M, C, seed=1)
    w = torch.randn(C, C, 3, 1, 1, device="cuda") * (3 * C) ** -0.5
    b = torch.randn(C, device="cuda")
    res = _r(M, C, seed=2)
    out = ops.gemm(x, packing.pack_tconv3(w), bias=b, mode=ops.TCONV3, geom=(B, Fr, H, W), residual=res, variant=variant)
    x5 = x.float().reshape(B, Fr, H, W, C).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(x5, w.half().float(), b, padding=(1, 0, 0)).permute(0, 2, 3, 4, 1).reshape(M, C) + res.float()
    assert_close(f"tconv3 B{B} F{Fr} {H}x{W} C{C}", out, ref)


@pytest.mark.parametrize("M,N,K,split", [(384, 1280, 11520, 6), (768, 1280, 3840, 4), (130, 64, 640, 3)])
def test_split_k(M, N, K, split, variant):
    from videomv_b200 import ops
    a, w = _r(M, K, seed=1), _r(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = _r(M, N, seed=3)
    out = ops.gemm(a, w, bias=bias, residual=res, split_k=split, variant=variant)
    ref = a.float() @ w.float().t() + bias + res.float()
    assert_close(f"splitk M{M} N{N} K{K} s{split}", out, ref)


@pytest.mark.parametrize("M,C,N,geglu", [(24576, 320, 960, False), (1536, 1280, 1280, False), (1000, 512, 4096, True),
                                         (6144, 640, 5120, True)])
def test_folded_layernorm(M, C, N, geglu, variant):
    """LayerNorm folded into the consumer GEMM == nn.LayerNorm -> nn.Linear (-> GEGLU) on the same fp16 rows."""
    from videomv_b200 import ops, packing
    h = _r(M, C, seed=1, scale=1.7) + 0.3
    w = torch.randn(N, C, device="cuda") * C ** -0.5
    b = torch.randn(N, device="cuda")
    gamma = 1 + 0.2 * torch.randn(C, device="cuda")
    beta = 0.2 * torch.randn(C, device="cuda")
    wg, bf = packing.fold_layernorm(w, b, gamma, beta)
    if geglu:
        wp, bp, bn = packing.pack_geglu(wg, bf)
    else:
        wp, bp, bn = wg.half().contiguous(), bf, 0
    colsum = wp.float().sum(1).contiguous()
    out = ops.gemm(h, wp, bias=bp, ln_stats=ops.layernorm_stats(h), ln_colsum=colsum, block_n=bn,
                   act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, variant=variant)
    ref = F.linear(F.layer_norm(h.float(), (C,), gamma, beta, 1e-5), w, b)
    if geglu:
        val, gate = ref.chunk(2, dim=-1)
        ref = val * F.gelu(gate)
    # The reference here is fp32 end to end, so unlike the other tests the bound must also cover the fp16 rounding of
    # the weights (W*gamma here, W and LN(h) in the unfused pair): ~2x fp16 eps on each GEMM output, and GEGLU multiplies
    # two such outputs (value * gelu(gate)), hence the doubled bound for it.
    tol = 6e-3 if geglu else 2e-3
    assert_close(f"folded LN M{M} C{C} N{N} geglu{int(geglu)}", out, ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("M,N,K,split", [(49152, 320, 320, 0), (12288, 1920, 640, 0), (3072, 1280, 11520, 0), (256, 128, 64, 0),
                                         (768, 1280, 3840, 4), (200, 2560, 128, 0)])
def test_static_weight_prefetch(M, N, K, split):
    """w_static: the CTA-pair kernel issues the W tiles of its first ring pass BEFORE griddepcontrol.wait (programmatic
    dependent launch).  Same result as the ordered path, including when fewer K blocks than stages exist."""
    from videomv_b200 import ops
    a, w = _r(M, K, seed=1), _r(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = _r(M, N, seed=3)
    torch.cuda.synchronize()
    base = ops.gemm(a, w, bias=bias, residual=res, split_k=split, variant=2)
    out = ops.gemm(a, w, bias=bias, residual=res, split_k=split, variant=2, w_static=True)
    assert_close(f"w_static M{M} N{N} K{K}", out, a.float() @ w.float().t() + bias + res.float())
    if split == 0:
        assert torch.equal(out, base)


def test_dependent_launch_chain():
    """A chain of kernels each consuming the previous one's output, launched back to back (every launch carries the
    programmatic-stream-serialization attribute): GEMM -> LayerNorm stats -> folded GEMM -> GroupNorm -> conv -> GEMM, 20 times,
    eagerly and from a CUDA graph.  Any kernel reading before its griddepcontrol.wait would see stale data."""
    from videomv_b200 import ops, packing
    M, C, HW = 24 * 256, 640, 256
    x0 = _r(M, C, seed=1)
    w1, w2, w3 = (_r(C, C, scale=C ** -0.5, seed=s) for s in (2, 3, 4))
    wc = packing.pack_conv3x3(torch.randn(C, C, 3, 3, device="cuda") * (9 * C) ** -0.5)
    gamma, beta = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    cs = w2.float().sum(1).contiguous()
    arena = ops.GnArena("cuda")

    def chain(x):
        arena.reset()
        for _ in range(5):
            h = ops.gemm(x, w1, residual=x, w_static=True)
            h = ops.gemm(h, w2, ln_stats=ops.layernorm_stats(h), ln_colsum=cs, w_static=True)
            g = ops.groupnorm(h, gamma, beta, rows_per_batch=HW, eps=1e-5, silu=True, scratch=arena)
            h = ops.gemm(g, wc, mode=ops.CONV3X3, geom=(1, 24, 16, 16), residual=h, w_static=True)
            x = ops.gemm(h, w3, act=ops.ACT_SILU, w_static=True)
        return x

    def ref_chain(x):
        x = x.float()
        w1f, w2f, w3f = w1.float(), w2.float(), w3.float()
        wcf = wc.float().reshape(C, 3, 3, C).permute(0, 3, 1, 2)
        for _ in range(5):
            h = (x @ w1f.t() + x).half().float()
            h = (F.layer_norm(h, (C,)) @ w2f.t()).half().float()
            g = F.silu(F.group_norm(h.reshape(24, 16, 16, C).permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).half().float()
            h = (F.conv2d(g, wcf, padding=1).permute(0, 2, 3, 1).reshape(M, C) + h).half().float()
            x = F.silu(h @ w3f.t()).half().float()
        return x

    ref = ref_chain(x0)
    outs = [chain(x0) for _ in range(4)]
    torch.cuda.synchronize()
    for o in outs:
        rel = ((o.float() - ref).norm() / ref.norm()).item()
        assert rel < 5e-3, rel
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain(x0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        og = chain(x0)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    rel = ((og.float() - ref).norm() / ref.norm()).item()
    assert rel < 5e-3, rel
    # no floating-point atomics anywhere on the path: eager runs and graph replays are bit-identical
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert torch.equal(og, outs[0])


def _merge_slots(rs, N, bn):
    """Host restatement of the consumer's slot merge: rs [M, slots, 2] of {mean_k, M2_k} -> (mean, biased variance)."""
    from videomv_b200 import _lib
    S = int(_lib.lib().vmv_gemm_epilogue_split())
    M, nsl, _ = rs.shape
    n = torch.zeros(M, dtype=torch.float64, device=rs.device)
    mean, m2 = torch.zeros_like(n), torch.zeros_like(n)
    for s_ in range(nsl):
        nt, hh = s_ // S, s_ % S
        nvalid = max(min(bn // 32, (N - nt * bn + 31) // 32), 0)
        nk = 32 * max((nvalid - hh + S - 1) // S, 0)
        if nk <= 0:
            continue
        mk, m2k = rs[:, s_, 0].double(), rs[:, s_, 1].double()
        nn = n + nk
        d = mk - mean
        mean = mean + d * nk / nn
        m2 = m2 + m2k + d * d * n * nk / nn
        n = nn
    return mean, m2 / n


@pytest.mark.parametrize("M,N,K", [(24576, 320, 320), (6144, 640, 2560), (1536, 1280, 1280), (1000, 512, 320), (130, 128, 64)])
def test_rowstats_feed_folded_layernorm(M, N, K):
    """rowstats_out: the producing GEMM writes per-slot partial LayerNorm statistics {mean_k, M2_k} of its output rows; the
    consuming GEMM merges them in slot order and folds the LayerNorm.  Same result as nn.LayerNorm -> nn.Linear on the
    producer's fp16 output; bit-identical from run to run (one writer per slot, no atomics)."""
    from videomv_b200 import ops, packing
    a, w = _r(M, K, seed=1), _r(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda") + 0.4
    res = _r(M, N, seed=3)
    bn = ops.gemm_block_n(N, variant=2)
    nsl = ops.rowstats_slots(N, bn)
    rs = torch.full((M, nsl, 2), float("nan"), device="cuda")                 # no initialisation required
    h = ops.gemm(a, w, bias=bias, residual=res, rowstats_out=rs, variant=2)
    ref_h = a.float() @ w.float().t() + bias + res.float()
    assert_close(f"rowstats producer M{M} N{N} K{K}", h, ref_h)
    mean, var = _merge_slots(rs, N, bn)
    assert torch.allclose(mean.float(), ref_h.mean(1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(var.float(), ref_h.var(1, unbiased=False), rtol=1e-3, atol=1e-4)
    N2 = 384
    w2 = torch.randn(N2, N, device="cuda") * N ** -0.5
    b2 = torch.randn(N2, device="cuda")
    gamma, beta = 1 + 0.2 * torch.randn(N, device="cuda"), 0.2 * torch.randn(N, device="cuda")
    wg, bf = packing.fold_layernorm(w2, b2, gamma, beta)
    wp = wg.half().contiguous()
    colsum = wp.float().sum(1).contiguous()
    out = ops.gemm(h, wp, bias=bf, ln_stats=rs, ln_colsum=colsum, ln_src=(N, bn), ln_eps=1e-5, variant=2)
    base = ops.gemm(h, wp, bias=bf, ln_stats=ops.layernorm_stats(h), ln_colsum=colsum, variant=2)
    ref = F.linear(F.layer_norm(h.float(), (N,), gamma, beta, 1e-5), w2, b2)
    assert_close(f"rowstats consumer M{M} N{N}", out, ref, rtol=2e-3, atol=2e-3)
    assert_close(f"rowstats consumer vs stats kernel M{M} N{N}", out, base.float(), rtol=1e-3, atol=1e-3)
    rs2 = torch.empty_like(rs)
    h2 = ops.gemm(a, w, bias=bias, residual=res, rowstats_out=rs2, variant=2)
    out2 = ops.gemm(h2, wp, bias=bf, ln_stats=rs2, ln_colsum=colsum, ln_src=(N, bn), ln_eps=1e-5, variant=2)
    assert torch.equal(rs, rs2) and torch.equal(h, h2) and torch.equal(out, out2)
    rs1 = torch.empty(M, ops.rowstats_slots(N, ops.gemm_block_n(N, variant=1)), 2, device="cuda")
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, rowstats_out=rs1, variant=1)


@pytest.mark.parametrize("mean,std", [(50.0, 1.0), (-120.0, 0.5), (8.0, 4.0)])
def test_folded_layernorm_survives_large_row_means(mean, std):
    """Residual streams of real checkpoints carry large per-row means and outlier channels.  E[x^2] - E[x]^2 in fp32 loses
    the variance there; the slot statistics are shifted sums merged with the parallel-variance formula, so the folded
    LayerNorm still matches nn.LayerNorm.  The rows are produced by a GEMM whose residual carries the
    offset, exactly like attn.to_out + x in BasicTransformerBlock (util.py:536-540)."""
    from videomv_b200 import ops, packing
    M, N, K = 4096, 640, 320
    a, w = _r(M, K, seed=5), _r(N, K, scale=K ** -0.5 * std, seed=6)
    g = torch.Generator(device="cuda").manual_seed(7)
    res = (mean + std * 0.3 * torch.randn(M, N, generator=g, device="cuda")).half()
    res[:, 17] += 30 * std                                                    # an outlier channel
    bn = ops.gemm_block_n(N, variant=2)
    rs = torch.empty(M, ops.rowstats_slots(N, bn), 2, device="cuda")
    h = ops.gemm(a, w, residual=res, rowstats_out=rs, variant=2)
    N2 = 256
    w2 = torch.randn(N2, N, device="cuda") * N ** -0.5
    b2 = torch.randn(N2, device="cuda")
    gamma, beta = 1 + 0.2 * torch.randn(N, device="cuda"), 0.2 * torch.randn(N, device="cuda")
    wg, bf = packing.fold_layernorm(w2, b2, gamma, beta)
    wp = wg.half().contiguous()
    colsum = wp.float().sum(1).contiguous()
    out = ops.gemm(h, wp, bias=bf, ln_stats=rs, ln_colsum=colsum, ln_src=(N, bn), ln_eps=1e-5, variant=2)
    # the statistics the epilogue produced (of the fp32 values before the fp16 store) vs an fp64 reference of the same rows
    exact = a.double() @ w.double().t() + res.double()
    m_, v_ = _merge_slots(rs, N, bn)
    assert torch.allclose(m_, exact.mean(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(v_, exact.var(1, unbiased=False), rtol=2e-4, atol=1e-6), ((v_ - exact.var(1, unbiased=False)).abs() / v_).max()
    # end to end: LayerNorm of the stored fp16 rows -> Linear, against fp64 maths on those rows.  The folded form computes
    # rstd * (acc - mean * colsum) where acc and mean*colsum are ~|mean|/std larger than the result, so its error grows
    # linearly with |mean|/std (fp32 accumulation of the big common-mode term; measured on B200: max|d| 6.6e-3 at ratio 50,
    # 1.5e-2 at ratio 240): bound it by that, not by the north-star per-element rtol.
    ref = F.linear(F.layer_norm(h.double(), (N,), gamma.double(), beta.double(), 1e-5), w2.double(), b2.double()).float()
    amp = max(abs(mean) / std, 1.0)
    assert_close(f"folded LN mean {mean} std {std}", out, ref, rtol=2e-3, atol=2e-3 + 6e-5 * amp)


def test_bad_args_raise():
    from videomv_b200 import ops
    a, w = _r(128, 100), _r(64, 100)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w)           # K not a multiple of 64
