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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    b1 = "--b1" in sys.argv                 # one B=1 forward (what a rank of the CFG-split multi-GPU mode runs) instead of the CFG pair
    hw = int(args[0]) if args else 32
    dev = torch.device("cuda")
    with torch.device(dev):
        model = unet.UNetSD_T2VBase(**bench.T2V_KWARGS)
    synth.fill_module_fast(model, seed=0)
    model.eval()
    host = bench.make_host_inputs("t2v", hw, seed=11)
    kw = bench.to_kwargs("t2v", host, dev)
    x = host["noise"].to(dev)
    t = torch.full((1,), 981, dtype=torch.long, device=dev)
    run = (lambda: model(x, t, **kw[0])) if b1 else (lambda: model.forward_cfg_pair(x, t, kw[0], kw[1]))
    for _ in range(2):
        run()
    ops.PROFILE = []
    torch.cuda.synchronize()
    run()
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
    print(f"gemm total {tot / 1e3:.3f} ms per B={1 if b1 else 2} forward, {sum(n * fl for _, n, fl, _ in rows) / tot / 1e6:.0f} TFLOP/s "
          f"({len(rows)} distinct shapes, {sum(r[1] for r in rows)} launches)")
    from videomv_b200.profiling import family_shape_times
    for fam_name in ("groupnorm", "attention"):
        frows = family_shape_times(prof, fam_name)
        ftot = sum(n * us for _, n, _, us, _ in frows)
        print(f"\n{fam_name}: {ftot / 1e3:.3f} ms, {sum(r[1] for r in frows)} launches")
        for desc, n, fl, us, by in sorted(frows, key=lambda r: -r[1] * r[3]):
            print(f"| {desc} | {n} | {us:.1f} us | {n * us / 1e3:.3f} ms | {by / us / 1e3:.0f} GB/s | {fl / us / 1e6:.0f} TFLOP/s |")
    # whole forward from a graph
    model.enable_cuda_graphs(True)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"\ngraph replay of the whole forward: {e0.elapsed_time(e1) / 10:.3f} ms, {model.graph_launches()} kernels")


if __name__ == "__main__":
    main()
