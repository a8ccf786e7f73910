"""Where does the time go in the persistent CTA-pair GEMM?  Graph-replay timings of a few shapes with the
VMV_GEMM_DEBUG knobs (bit 1 = no global stores, 2 = no epilogue body, 4 = no TMA operand loads, 8 = no MMA issue).
The knob is read once per process, so the driver re-executes itself per setting."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(49152, 960, 320, False, False), (49152, 320, 320, True, True), (49152, 320, 2880, False, True),
          (12288, 1280, 1280, False, True), (3072, 1280, 1280, True, True)]


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


def worker():
    from videomv_b200 import ops
    dbg = os.environ.get("VMV_GEMM_DEBUG", "0")
    dev = "cuda"
    for (M, N, K, res, bias) in SHAPES:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
        b = torch.randn(N, device=dev) if bias else None
        r = torch.randn(M, N, device=dev).half() if res else None
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        row = [f"dbg{dbg:>2} M{M} N{N} K{K} res{int(res)} bias{int(bias)}:"]
        for bn in (160, 128, 256):
            if N % bn:
                continue
            us = bench(lambda: ops.gemm(a, w, out=out, bias=b, residual=r, block_n=bn, variant=2))
            row.append(f"bn{bn} {us:7.1f}us {2.0 * M * N * K / us / 1e6:6.0f}TF")
        print(" ".join(row), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
    else:
        for d in (sys.argv[1:] or ["0", "2", "6", "10", "14"]):
            env = dict(os.environ, VMV_GEMM_DEBUG=d)
            subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, timeout=300)
