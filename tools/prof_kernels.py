"""One launch of each hot kernel on its level-0 (24 x 32 x 32, CFG batch 2) shape, bracketed by cudaProfilerStart/Stop:
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/kernels python tools/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops, packing  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).half()
    B, Fr, HW, C, heads = 2, 24, 1024, 320, 5
    M = B * Fr * HW
    x = r(M, C)
    res = r(M, C)
    w_cc = r(C, C, sc=C ** -0.5)
    bias = torch.randn(C, device=dev)
    w_geglu, b_geglu, bn = packing.pack_geglu(torch.randn(8 * C, C, device=dev) * C ** -0.5, torch.randn(8 * C, device=dev))
    colsum = w_geglu.float().sum(1).contiguous()
    w_qkv = r(3 * C, C, sc=C ** -0.5)
    w_conv = packing.pack_conv3x3(torch.randn(C, C, 3, 3, device=dev) * (9 * C) ** -0.5)
    w_conv2 = packing.pack_conv3x3(torch.randn(2 * C, 2 * C, 3, 3, device=dev) * (18 * C) ** -0.5)
    x2 = r(M // 4, 2 * C)
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    arena = ops.GnArena(dev, 32 << 20)
    qkv = r(M, 3 * C)
    o = torch.empty(M, C, device=dev, dtype=torch.float16)
    ld = 3 * C

    def run():
        arena.reset()
        src = (C, ops.gemm_block_n(C))
        rs = arena.take_rowstats(M, ops.rowstats_slots(*src))
        h = ops.gemm(x, w_cc, bias=bias, residual=res, rowstats_out=rs, w_static=True)                      # o-proj like
        ops.gemm(h, w_geglu, bias=b_geglu, act=ops.ACT_GEGLU, block_n=bn, ln_stats=rs, ln_src=src, ln_colsum=colsum,
                 w_static=True)                                                                             # ff1 GEGLU
        ops.gemm(x, w_qkv, ln_stats=rs, ln_src=src, ln_colsum=w_qkv.float().sum(1).contiguous(), w_static=True)   # QKV
        ops.gemm(x, w_conv, bias=bias, mode=ops.CONV3X3, geom=(1, B * Fr, 32, 32), residual=res, w_static=True)  # 3x3 N320
        ops.gemm(x2, w_conv2, mode=ops.CONV3X3, geom=(1, B * Fr, 16, 16), w_static=True)                    # 3x3 N640 K5760
        ops.groupnorm(x, gamma, beta, rows_per_batch=Fr * HW, eps=1e-5, silu=True, scratch=arena)           # 5-D GN
        ops.groupnorm(x, gamma, beta, rows_per_batch=HW, eps=1e-5, silu=True, scratch=arena)                # 4-D GN
        st = (Fr * HW * ld, ld, HW * ld)
        ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, outer=B, inner=HW, heads=heads, nq=Fr, nk=Fr,
                      q_strides=st, k_strides=st, v_strides=st, o_strides=(Fr * HW * C, C, HW * C))        # temporal
        st = (HW * ld, 0, ld)
        ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, outer=B * Fr, inner=1, heads=heads, nq=HW, nk=HW,
                      q_strides=st, k_strides=st, v_strides=st, o_strides=(HW * C, 0, C))                  # spatial (tcgen05)

    run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done")


if __name__ == "__main__":
    main()
