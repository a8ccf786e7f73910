"""GroupNorm micro-benchmark on the shapes of one T2V 256^2 forward (24 frames; B = 2 CFG-batched, or `B=1` as argv[2]).
Each shape: `reps` back-to-back calls captured in a CUDA graph, over a rotating set of distinct buffers large enough
to defeat L2 ("cold") or over one buffer ("warm"); CUDA events around graph replays.  argv[1]: comma-separated values of
VMV_GN_MIN_KB (minimum KB of rows per CTA of the smem-resident kernel, 0 = one CTA per SM whatever the size).  Also
prints a plain copy (read + write) of the same tensor as the bandwidth yardstick."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops  # noqa: E402

# (rows, C1, C2, rows_per_batch, silu, count per forward, label)
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_R = 24576 * BATCH
SHAPES = [
    (_R, 320, 0, 24576, True, 26, "5-D 32x32 C320"),
    (_R // 4, 640, 0, 6144, True, 25, "5-D 16x16 C640"),
    (_R // 16, 1280, 0, 1536, True, 25, "5-D 8x8 C1280"),
    (_R // 64, 1280, 0, 384, True, 29, "5-D 4x4 C1280"),
    (_R, 320, 0, 1024, True, 12, "4-D 32x32 C320"),
    (_R, 640, 320, 1024, True, 2, "4-D 32x32 C640+320"),
    (_R // 4, 640, 0, 256, True, 11, "4-D 16x16 C640"),
    (_R // 4, 1280, 640, 256, True, 1, "4-D 16x16 C1280+640"),
    (_R // 16, 1280, 0, 64, True, 12, "4-D 8x8 C1280"),
    (_R // 16, 1280, 1280, 64, True, 2, "4-D 8x8 C1280+1280"),
    (_R // 64, 1280, 0, 16, True, 12, "4-D 4x4 C1280"),
]


def time_graph(fn, reps, rounds=3):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * rounds)


def main():
    opts = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
    dev = "cuda"
    arena = ops.GnArena(dev, 64 << 20)
    print(f"B={BATCH} | shape | n | MB in | copy cold us | " + " | ".join(f"min_kb {o} cold / warm us" for o in opts) + " |")
    tot = {o: [0.0, 0.0] for o in opts}
    for rows, C1, C2, rpb, silu, cnt, label in SHAPES:
        C = C1 + C2
        mb = rows * C * 2 / 1e6
        nbuf = max(2, int(400 / mb) + 1)                      # > 3x L2 of distinct inputs+outputs
        nbuf = min(nbuf, 64)
        xs = [torch.randn(rows, C1, device=dev).half() for _ in range(nbuf)]
        x2s = [torch.randn(rows, C2, device=dev).half() for _ in range(nbuf)] if C2 else None
        outs = [torch.empty(rows, C, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        reps = nbuf * 2

        def copy(i):
            outs[i % nbuf][:, :C1].copy_(xs[i % nbuf])
        t_copy = time_graph(copy, reps)
        cells = []
        for o in opts:
            os.environ["VMV_GN_MIN_KB"] = str(o)

            def cold(i):
                if i == 0:
                    arena.reset()
                ops.groupnorm(xs[i % nbuf], gamma, beta, rows_per_batch=rpb, eps=1e-5, silu=silu,
                              x2=None if x2s is None else x2s[i % nbuf], out=outs[i % nbuf], scratch=arena)

            def warm(i):
                if i == 0:
                    arena.reset()
                ops.groupnorm(xs[0], gamma, beta, rows_per_batch=rpb, eps=1e-5, silu=silu,
                              x2=None if x2s is None else x2s[0], out=outs[0], scratch=arena)
            tc, tw = time_graph(cold, reps), time_graph(warm, reps)
            tot[o][0] += cnt * tc
            tot[o][1] += cnt * tw
            cells.append(f"{tc:.1f} / {tw:.1f}")
        print(f"| {label} | {cnt} | {mb:.1f} | {t_copy:.1f} | " + " | ".join(cells) + " |", flush=True)
        del xs, x2s, outs
    print("per-forward totals (ms): " + "; ".join(f"min_kb {o}: cold {v[0] / 1e3:.3f} warm {v[1] / 1e3:.3f}" for o, v in tot.items()))


if __name__ == "__main__":
    main()
