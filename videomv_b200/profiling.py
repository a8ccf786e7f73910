"""Per-launch device timing for the roofline leg of bench.py and tools/gemm_breakdown.py.

`ops.PROFILE = []` makes every op record (family, algorithmic FLOPs, algorithmic bytes, start/end event, description,
replay closure).  Event pairs recorded around eager launches include the host-side launch gap (for the many small
kernels of this network that gap exceeds the kernel), so the *device* time of a launch is measured by replaying it:
8 back-to-back launches captured in a CUDA graph, replayed 3 times, CUDA events around the replays (L2-warm).
"""
from __future__ import annotations

import collections

import torch


def replay_us(replay, reps: int = 8, rounds: int = 3) -> float:
    pre = getattr(replay, "pre", None)          # e.g. GroupNorm: zero its scratch arena once per graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        if pre is not None:
            pre()
        replay()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        if pre is not None:
            pre()
        for _ in range(reps):
            replay()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * rounds)


def family_shape_times(prof, family):
    """prof: the list collected in ops.PROFILE during ONE forward.  Returns rows (desc, launches, flops each, device us
    each, bytes each) for every distinct shape of one kernel family, timed by graph replay."""
    shapes = collections.OrderedDict()
    for name, fl, by, a, b, desc, replay in prof:
        if name == family and replay is not None:
            g = shapes.setdefault(desc, [0, fl, replay, by])
            g[0] += 1
    return [(desc, n, fl, replay_us(replay), by) for desc, (n, fl, replay, by) in shapes.items()]


def gemm_shape_times(prof):
    return [r[:4] for r in family_shape_times(prof, "gemm_tc")]
