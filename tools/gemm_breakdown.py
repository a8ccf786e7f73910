"""Per-shape timing of every GEMM launch in one CFG-batched (B=2) T2V forward: CUDA events around each vmv_gemm call
(eager, no graph).  Prints shapes sorted by total time with achieved TFLOP/s."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from videomv_b200 import ops, synth, unet  # noqa: E402


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda")
    with torch.device(dev):
        model = unet.UNetSD_T2VBase(**bench.T2V_KWARGS)
    synth.fill_module_fast(model, seed=0)
    model.eval()
    host = bench.make_host_inputs("t2v", hw, seed=11)
    kw = bench.to_kwargs("t2v", host, dev)
    x = host["noise"].to(dev)
    t = torch.full((1,), 981, dtype=torch.long, device=dev)
    for _ in range(2):
        model.forward_cfg_pair(x, t, kw[0], kw[1])
    ops.PROFILE = []
    torch.cuda.synchronize()
    model.forward_cfg_pair(x, t, kw[0], kw[1])
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    fam = collections.OrderedDict()
    shapes = collections.OrderedDict()
    for name, fl, by, a, b, desc, replay in prof:
        ms = a.elapsed_time(b)
        f = fam.setdefault(name, [0, 0.0, 0.0]); f[0] += 1; f[1] += fl; f[2] += ms
    print("family totals, eager event-to-event (ms; small kernels are CPU-launch bound here):",
          {k: (v[0], round(v[2], 3)) for k, v in fam.items()})
    # true device time per shape: 8 back-to-back launches captured in a CUDA graph, replayed 3x, L2-warm
    from videomv_b200.profiling import gemm_shape_times
    rows = gemm_shape_times(prof)
    tot = sum(n * us for _, n, _, us in rows)
    print("| shape | n | us each | total ms | share | TFLOP/s |\n|---|---:|---:|---:|---:|---:|")
    for desc, n, fl, us in sorted(rows, key=lambda r: -r[1] * r[3]):
        print(f"| {desc} | {n} | {us:.1f} | {n * us / 1e3:.3f} | {100 * n * us / tot:.1f}% | {fl / us / 1e6:.0f} |")
    print(f"gemm total {tot / 1e3:.3f} ms per B=2 forward, {sum(n * fl for _, n, fl, _ in rows) / tot / 1e6:.0f} TFLOP/s "
          f"({len(rows)} distinct shapes, {sum(r[1] for r in rows)} launches)")


if __name__ == "__main__":
    main()
