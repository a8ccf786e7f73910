"""Profiling driver: N eager (no CUDA graph) CFG-batched UNet forwards of the T2V 24x32x32 workload.
Run under ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from videomv_b200 import synth, unet  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    hw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda")
    with torch.device(dev):
        model = unet.UNetSD_T2VBase(**bench.T2V_KWARGS)
    synth.fill_module_fast(model, seed=0)
    model.eval()
    host = bench.make_host_inputs("t2v", hw, seed=11)
    kw = bench.to_kwargs("t2v", host, dev)
    x = host["noise"].to(dev)
    t = torch.full((1,), 981, dtype=torch.long, device=dev)
    for i in range(n):
        if i == n - 1:                       # profile only the last forward: run ncu with --profile-from-start off
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        model.forward_cfg_pair(x, t, kw[0], kw[1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done")


if __name__ == "__main__":
    main()
