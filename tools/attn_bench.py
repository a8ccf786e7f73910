"""Graph-replay timing of the two attention kernels (impl 1 = strided mma.sync, impl 2 = tcgen05/TMEM) on the model's
long-sequence shapes (CFG batch 2: 48 frames)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops  # noqa: E402
from videomv_b200.profiling import replay_us  # noqa: E402


def main():
    dev = "cuda"
    for name, NF, HW, heads, L in [("spatial 32x32 C320", 48, 1024, 5, 0), ("spatial 16x16 C640", 48, 256, 10, 0),
                                   ("spatial 64x64 C320 (512^2)", 48, 4096, 5, 0), ("spatial 32x32 C640 (512^2)", 48, 1024, 10, 0),
                                   ("cross 32x32 C320 L77", 48, 1024, 5, 77), ("cross 64x64 C320 L77", 48, 4096, 5, 77)]:
        C = heads * 64
        if L == 0:
            qkv = torch.randn(NF * HW, 3 * C, device=dev).half()
            out = torch.empty(NF * HW, C, device=dev, dtype=torch.float16)
            ld = 3 * C
            st = (HW * ld, 0, ld)
            call = lambda impl: ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], out, outer=NF, inner=1, heads=heads, nq=HW,
                                              nk=HW, q_strides=st, k_strides=st, v_strides=st, o_strides=(HW * C, 0, C), impl=impl)
            flops = 4.0 * NF * heads * HW * HW * 64
        else:
            q = torch.randn(NF * HW, C, device=dev).half()
            kv = torch.randn(2 * L, 2 * C, device=dev).half()
            out = torch.empty_like(q)
            call = lambda impl: ops.attention(q, kv, kv[:, C:], out, outer=NF, inner=1, heads=heads, nq=HW, nk=L,
                                              q_strides=(HW * C, 0, C), k_strides=(L * 2 * C, 0, 2 * C),
                                              v_strides=(L * 2 * C, 0, 2 * C), o_strides=(HW * C, 0, C), kv_group=NF // 2, impl=impl)
            flops = 4.0 * NF * heads * HW * L * 64
        row = [f"{name:30s}"]
        for impl in (1, 2):
            us = replay_us(lambda: call(impl), reps=4)
            row.append(f"impl{impl} {us:8.1f}us {flops / us / 1e6:6.0f}TF")
        print(" ".join(row), flush=True)


if __name__ == "__main__":
    main()
