"""Peer-memory layout exchange / statistics all-reduce kernels (csrc/peer.cu) on ONE GPU: the P ranks are simulated one
after the other in this process (their "peer" pointers are local buffers, nowait skips the flag wait), which checks the data
movement, the flag / epoch protocol state and the world=1 path with the real wait.  The true multi-process path (CUDA IPC,
NVLink, concurrent waits) is tests/test_sharded_gpu.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _exchange(src, dsts, flags, ctrl, world, rank, direction, B, Fl, HWl, C, nowait):
    from videomv_b200 import _lib, ops
    p = _lib.PeerExchangeParams()
    p.src = src.data_ptr()
    for q in range(world):
        p.dst[q] = dsts[q].data_ptr()
        p.flags[q] = flags[q].data_ptr()
    p.epoch = ctrl.data_ptr()
    p.done = ctrl.data_ptr() + 4
    p.world, p.rank, p.direction = world, rank, direction
    p.B, p.Fl, p.HWl, p.C, p.nowait = B, Fl, HWl, C, nowait
    _lib.check(_lib.lib().vmv_peer_exchange(ctypes.byref(p), ops._stream()), "vmv_peer_exchange")


@pytest.mark.parametrize("P,B,F,HW,C", [(2, 2, 24, 1024, 320), (4, 1, 24, 64, 1280), (8, 2, 24, 16, 1280), (2, 1, 4, 16, 8)])
def test_exchange_simulated_ranks(P, B, F, HW, C):
    Fl, HWl = F // P, HW // P
    g = torch.Generator(device="cuda").manual_seed(0)
    full = torch.randn(B, F, HW, C, generator=g, device="cuda").half()
    xa = [full[:, r * Fl:(r + 1) * Fl].reshape(B * Fl * HW, C).contiguous() for r in range(P)]          # layout A per rank
    want_b = [full[:, :, q * HWl:(q + 1) * HWl].reshape(B * F * HWl, C).contiguous() for q in range(P)]  # layout B per rank
    yb = [torch.zeros(B * F * HWl, C, dtype=torch.float16, device="cuda") for _ in range(P)]
    flags = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(P)]
    ctrl = [torch.zeros(2, dtype=torch.int32, device="cuda") for _ in range(P)]
    for r in range(P):
        _exchange(xa[r], yb, flags, ctrl[r], P, r, 0, B, Fl, HWl, C, nowait=1)
    torch.cuda.synchronize()
    for q in range(P):
        assert torch.equal(yb[q], want_b[q]), f"frames->pixels, rank {q}"
        assert flags[q][:P].tolist() == [1] * P and ctrl[q].tolist() == [1, 0]
    ya = [torch.zeros(B * Fl * HW, C, dtype=torch.float16, device="cuda") for _ in range(P)]
    for r in range(P):
        _exchange(yb[r], ya, flags, ctrl[r], P, r, 1, B, Fl, HWl, C, nowait=1)
    torch.cuda.synchronize()
    for q in range(P):
        assert torch.equal(ya[q], xa[q]), f"pixels->frames, rank {q}"
        assert flags[q][:P].tolist() == [2] * P and ctrl[q].tolist() == [2, 0]


def test_world1_with_real_wait_and_allreduce():
    from videomv_b200 import _lib, ops
    x = torch.randn(24 * 64, 320, device="cuda").half()
    y = torch.zeros_like(x)
    flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    ctrl = torch.zeros(2, dtype=torch.int32, device="cuda")
    for it in range(3):
        _exchange(x, [y], [flags], ctrl, 1, 0, it & 1, 1, 24, 64, 320, nowait=0)
    torch.cuda.synchronize()
    assert torch.equal(x, y) and flags[0].item() == 3

    # statistics all-reduce, 4 simulated ranks: pass 1 fills every rank's slots (sums incomplete: nowait), pass 2 repeats
    # the same partials and must produce the full sum on every rank, in rank order
    P, n = 4, 128
    parts = [torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(P)]
    slots = [torch.zeros(P * n, dtype=torch.float64, device="cuda") for _ in range(P)]
    fl = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(P)]
    ep = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(P)]
    res = None
    for _ in range(2):
        res = []
        for r in range(P):
            d = parts[r].clone()
            p = _lib.PeerAllreduceParams()
            p.data = d.data_ptr()
            for q in range(P):
                p.slots[q] = slots[q].data_ptr()
                p.flags[q] = fl[q].data_ptr()
            p.epoch = ep[r].data_ptr()
            p.world, p.rank, p.n, p.nowait = P, r, n, 1
            _lib.check(_lib.lib().vmv_peer_allreduce_f64(ctypes.byref(p), ops._stream()), "vmv_peer_allreduce_f64")
            res.append(d)
    torch.cuda.synchronize()
    want = parts[0].clone()
    for r in range(1, P):
        want = want + parts[r]
    for r in range(P):
        assert torch.equal(res[r], want), r


@pytest.mark.parametrize("P,B,F,HW,C,K,mode,split", [
    (2, 2, 24, 256, 320, 320, "linear", 0), (4, 1, 24, 64, 1280, 1280, "linear", 0), (2, 1, 8, 256, 128, 64, "conv", 0),
    (4, 2, 8, 64, 128, 128, "tconv", 0), (8, 1, 24, 16, 1280, 640, "linear", 0),
    (2, 1, 24, 16, 1280, 2560, "linear", 4), (2, 1, 8, 16, 256, 512, "tconv", 3), (4, 1, 4, 64, 128, 256, "conv", 4)])
def test_gemm_scatter_simulated_ranks(P, B, F, HW, C, K, mode, split):
    """vmv_gemm with `scatter`: the layout exchange inside the GEMM epilogue.  Every simulated rank runs its GEMM on its rows
    of the source layout and the epilogue stores them straight into all ranks' destination-layout tensors; the result must
    equal GEMM followed by the stand-alone exchange.  direction 0 (frames -> pixels) for linear / conv, 1 for the temporal conv."""
    from videomv_b200 import _lib, ops, packing
    Fl, HWl = F // P, HW // P
    g = torch.Generator(device="cuda").manual_seed(3)
    direction = 1 if mode == "tconv" else 0
    if mode == "conv":
        H = W = int(HW ** 0.5)
        w = packing.pack_conv3x3(torch.randn(C, K, 3, 3, generator=g, device="cuda") * (9 * K) ** -0.5)
    elif mode == "tconv":
        w = packing.pack_tconv3(torch.randn(C, K, 3, 1, 1, generator=g, device="cuda") * (3 * K) ** -0.5)
    else:
        w = (torch.randn(C, K, generator=g, device="cuda") * K ** -0.5).half()
    bias = torch.randn(C, generator=g, device="cuda")
    rows = B * Fl * HW                                                   # == B * F * HWl
    xs = [torch.randn(rows, K, generator=g, device="cuda").half() for _ in range(P)]
    res = [torch.randn(rows, C, generator=g, device="cuda").half() for _ in range(P)]
    dsts = [torch.zeros(rows, C, dtype=torch.float16, device="cuda") for _ in range(P)]
    flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(P)]
    plain = []
    for r in range(P):
        kw = dict(bias=bias, residual=res[r], split_k=split)         # split-K: the finish kernel scatters
        if mode == "conv":
            kw.update(mode=ops.CONV3X3, geom=(1, B * Fl, H, W))
        elif mode == "tconv":
            kw.update(mode=ops.TCONV3, geom=(B, F, HWl, 1))
        plain.append(ops.gemm(xs[r], w, **kw))
        sc = _lib.GemmScatter()
        sc.world, sc.rank, sc.direction, sc.nowait = P, r, direction, 1
        sc.B, sc.Fl, sc.HWl = B, Fl, HWl
        for q in range(P):
            sc.dst[q] = dsts[q].data_ptr()
            sc.flags[q] = flags[q].data_ptr()
        sc.epoch = flags[r].data_ptr() + 32
        sc.done = flags[r].data_ptr() + 36
        ops.gemm(xs[r], w, out=dsts[r], scatter=sc, **kw)
    torch.cuda.synchronize()
    if direction == 0:       # source [B, Fl, P(q), HWl, C] on rank r -> dst_q [B, P(r), Fl, HWl, C]
        full = torch.stack([pl.reshape(B, Fl, P, HWl, C) for pl in plain], 0)            # [r, B, Fl, q, HWl, C]
        want = [full[:, :, :, q].permute(1, 0, 2, 3, 4).reshape(rows, C) for q in range(P)]
    else:                    # source [B, P(q), Fl, HWl, C] on rank r -> dst_q [B, Fl, P(r), HWl, C]
        full = torch.stack([pl.reshape(B, P, Fl, HWl, C) for pl in plain], 0)            # [r, B, q, Fl, HWl, C]
        want = [full[:, :, q].permute(1, 2, 0, 3, 4).reshape(rows, C) for q in range(P)]
    for q in range(P):
        assert torch.equal(dsts[q], want[q].contiguous()), f"rank {q}"
        assert flags[q][:P].tolist() == [1] * P and flags[q][8].item() == 1 and flags[q][9].item() == 0


def test_allgather_simulated_ranks():
    """vmv_peer_allgather: the output all-gather of a sharded UNet call (frame shards x CFG halves), ranks simulated one after
    the other (nowait).  Every rank's buffer ends up holding [cfg half][B, C, F, h, w]; epochs advance by 2 per call."""
    from videomv_b200 import _lib, ops
    Wc, P, B, C, Fl, h, w = 2, 2, 1, 4, 6, 8, 8
    Wa, Fr = Wc * P, Fl * P
    g = torch.Generator(device="cuda").manual_seed(1)
    full = torch.randn(Wc, B, C, Fr, h, w, generator=g, device="cuda")
    dst = [torch.zeros_like(full) for _ in range(Wa)]
    flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(Wa)]
    for it in range(2):
        for r in range(Wa):
            ci, fr = r // P, r % P
            src = full[ci, :, :, fr * Fl:(fr + 1) * Fl].contiguous() + it
            p = _lib.PeerAllgatherParams()
            p.src = src.data_ptr()
            for q in range(Wa):
                p.dst[q] = dst[q].data_ptr()
                p.flags[q] = flags[q].data_ptr()
            p.epoch = flags[r].data_ptr() + 32
            p.done = flags[r].data_ptr() + 36
            p.world, p.rank, p.nowait = Wa, r, 1
            p.nouter, p.inner_bytes = B * C, Fl * h * w * 4
            p.dst_offset_bytes = (ci * B * C * Fr + fr * Fl) * h * w * 4
            p.dst_outer_stride_bytes = Fr * h * w * 4
            _lib.check(_lib.lib().vmv_peer_allgather(ctypes.byref(p), ops._stream()), "vmv_peer_allgather")
        torch.cuda.synchronize()
        for q in range(Wa):
            assert torch.equal(dst[q], full + it), (it, q)
            assert flags[q][:Wa].tolist() == [2 * (it + 1)] * Wa and flags[q][8].item() == 2 * (it + 1) and flags[q][9].item() == 0


# grids are sized so that ALL simulated ranks' kernels are co-resident on one GPU (P * B * min(148 // B, rows) <= 148 CTAs):
# on real multi-GPU runs every rank has its own device
FUSED_PEER_CASES = [(2, 2, 32, 320), (4, 1, 16, 1280), (2, 1, 64, 640)]


def _fused_peer_case(P, B, rows, C):
    import torch.nn.functional as F
    from videomv_b200 import _lib, ops
    from tests.util import assert_close
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [(torch.randn(B * rows, C, generator=g, device="cuda") * 1.5 + 0.3 * (r + 1)).half() for r in range(P)]
    gamma, beta = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    outs = [torch.empty_like(x) for x in xs]
    slots = [torch.zeros(P * B * 64, dtype=torch.float64, device="cuda") for _ in range(P)]
    ctrl = [torch.zeros(B * 16, dtype=torch.int32, device="cuda") for _ in range(P)]     # one 64-byte line per chunk
    arenas = [ops.GnArena("cuda", 1 << 20) for _ in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    torch.cuda.synchronize()
    for it in range(2):                                   # twice: the epochs advance
        for r in range(P):
            gp = _lib.GnPeer()
            gp.world, gp.rank, gp.stat_rows = P, r, P * rows
            for q in range(P):
                gp.slots[q] = slots[q].data_ptr()
                gp.flags[q] = ctrl[q].data_ptr()
            gp.epoch = ctrl[r].data_ptr() + 32                        # word 8 of line 0; line b is 64 bytes further
            with torch.cuda.stream(streams[r]):
                arenas[r].reset()
                bars, scr = arenas[r].take(C, rows, B)
                _lib.check(_lib.lib().vmv_groupnorm_fused_peer(xs[r].data_ptr(), C, C, None, 0, 0, rows, B, bars, scr,
                                                               gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, outs[r].data_ptr(), C,
                                                               ctypes.byref(gp), streams[r].cuda_stream), "vmv_groupnorm_fused_peer")
        torch.cuda.synchronize()
    full = torch.stack([x.float().reshape(B, rows, C) for x in xs], 1).reshape(B, P * rows, C)      # all ranks' rows of a sample
    ref = F.silu(F.group_norm(full.permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(B, P, rows, C)
    for r in range(P):
        assert_close(f"fused peer GN rank {r}", outs[r], ref[:, r].reshape(B * rows, C))
    assert all(c.view(B, 16)[:, 8].tolist() == [2] * B for c in ctrl)


def test_groupnorm_fused_peer_two_streams():
    """vmv_groupnorm_fused_peer: P simulated ranks run CONCURRENTLY on P streams of one GPU (small grids, so all kernels are
    co-resident) and exchange their partial statistics through local "peer" buffers.  Result == GroupNorm over all ranks' rows.
    Runs in a child process: the ranks wait for each other inside the kernels, so an environment that serialises kernel
    launches (CUDA_LAUNCH_BLOCKING, a profiler) makes the bounded waits trap -- that must not poison this process's context."""
    import os
    import subprocess
    import sys
    if os.environ.get("CUDA_LAUNCH_BLOCKING", "0") not in ("", "0"):
        pytest.skip("needs concurrent kernels on two streams")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests import test_peer_gpu as t\n"
            "for case in t.FUSED_PEER_CASES:\n"
            "    t._fused_peer_case(*case)\n"
            "print('fused-peer ok')\n") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "fused-peer ok" in r.stdout
