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
    agg = collections.OrderedDict()
    for name, fl, by, a, b, desc in prof:
        ms = a.elapsed_time(b)
        f = fam.setdefault(name, [0, 0.0, 0.0]); f[0] += 1; f[1] += fl; f[2] += ms
        if name == "gemm_tc":
            g = agg.setdefault(desc, [0, 0.0, 0.0]); g[0] += 1; g[1] += fl; g[2] += ms
    print("family totals (ms):", {k: (v[0], round(v[2], 3)) for k, v in fam.items()})
    tot = sum(v[2] for v in agg.values())
    print(f"| shape | n | total ms | share | us each | TFLOP/s |\n|---|---:|---:|---:|---:|---:|")
    for desc, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"| {desc} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {1e3 * ms / n:.1f} | {fl / ms / 1e9:.0f} |")
    print(f"gemm total {tot:.3f} ms, {sum(v[1] for v in agg.values()) / tot / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
