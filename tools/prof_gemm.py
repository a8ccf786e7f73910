"""Profiling driver: the dominant kernel on its biggest shapes (for `ncu --set full`).
  conv  : ResBlock 3x3 conv at 32x32, B*F=48, 320->320   (M=49152, N=320, K=2880)
  lin   : transformer GEGLU FF1 at 32x32, 320->2560       (M=49152, N=2560, K=320)
  conv8 : ResBlock 3x3 conv at 8x8, 1280->1280            (M=3072, N=1280, K=11520)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops, packing  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    which = sys.argv[2] if len(sys.argv) > 2 else "big"
    dev = "cuda"
    x = torch.randn(49152, 320, device=dev).half()
    if which == "big":
        w = packing.pack_conv3x3(torch.randn(320, 320, 3, 3, device=dev) * 0.02)
        b = torch.zeros(320, device=dev)
        wl = (torch.randn(2560, 320, device=dev) * 0.05).half()
        x8 = torch.randn(3072, 1280, device=dev).half()
        w8 = packing.pack_conv3x3(torch.randn(1280, 1280, 3, 3, device=dev) * 0.01)
        b8 = torch.zeros(1280, device=dev)
        for _ in range(reps):
            ops.gemm(x, w, bias=b, mode=ops.CONV3X3, geom=(1, 48, 32, 32))
            ops.gemm(x, wl)
            ops.gemm(x8, w8, bias=b8, mode=ops.CONV3X3, geom=(1, 48, 8, 8))
    else:
        # short-K, epilogue-heavy shapes of the transformer blocks
        w1 = (torch.randn(320, 320, device=dev) * 0.05).half()
        b1 = torch.randn(320, device=dev)
        res = torch.randn(49152, 320, device=dev).half()
        w2 = (torch.randn(960, 320, device=dev) * 0.05).half()
        wg, bg, bn = packing.pack_geglu(torch.randn(2560, 320, device=dev) * 0.05, torch.randn(2560, device=dev))
        for _ in range(reps):
            ops.gemm(x, w1, bias=b1, residual=res)
            ops.gemm(x, w2)
            ops.gemm(x, wg, bias=bg, act=ops.ACT_GEGLU, block_n=bn)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
