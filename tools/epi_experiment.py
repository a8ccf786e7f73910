"""Where does the time go in short-K GEMMs?  Times a few shapes under graph replay with VMV_GEMM_DEBUG knobs
(set the env var before running: 0 normal, 1 no TMA stores, 2 no epilogue body) and for several block_n."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops  # noqa: E402


def bench(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def main():
    dbg = os.environ.get("VMV_GEMM_DEBUG", "0")
    dev = "cuda"
    for (M, N, K, res, bias) in [(49152, 960, 320, False, False), (49152, 320, 320, True, True), (49152, 320, 320, False, False),
                                 (12288, 640, 640, True, True), (49152, 2560, 320, False, True), (3072, 1280, 1280, True, True)]:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
        b = torch.randn(N, device=dev) if bias else None
        r = torch.randn(M, N, device=dev).half() if res else None
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        row = [f"dbg{dbg} M{M} N{N} K{K} res{int(res)} bias{int(bias)}:"]
        for variant, bn in (((2, 160),) if int(dbg) else ((2, 160), (2, 128), (1, 160))):
            if N % bn and bn == 160:
                continue
            us = bench(lambda: ops.gemm(a, w, out=out, bias=b, residual=r, block_n=bn, variant=variant))
            row.append(f"v{variant}/bn{bn} {us:7.1f}us {2.0 * M * N * K / us / 1e6:6.0f}TF")
        print(" ".join(row), flush=True)


if __name__ == "__main__":
    main()
