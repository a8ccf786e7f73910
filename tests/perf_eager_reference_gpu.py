"""Manual GPU measurement (not collected by pytest): the reference's own eager-PyTorch CUDA path on this B200.

The reference cannot travel to the GPU box, so this times the oracle restatement (same torch ops, same order:
oracle/unet_oracle.py, pinned to the reference) on cuda, with `memory_efficient_attention` mapped to
F.scaled_dot_product_attention exactly like the survey's stub.  Rows B1 (fp32, TF32 convs as torch defaults) and B1h
(fp16 autocast) of BASELINE.md.  Output: one JSON line per row.
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle  # noqa: E402
from videomv_b200 import synth, unet  # noqa: E402
import bench  # noqa: E402


def sdpa_core(q, k, v, heads):
    b, nq, inner = q.shape
    d = inner // heads
    sp = lambda z: z.reshape(b, z.shape[1], heads, d).permute(0, 2, 1, 3)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.permute(0, 2, 1, 3).reshape(b, nq, inner)


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    unet_oracle.attention_core = sdpa_core
    torch.backends.cudnn.benchmark = True                      # inference_text2video_entrance.py:83
    dev = torch.device("cuda")
    kw = dict(bench.T2V_KWARGS)
    with torch.device(dev):
        model = unet.UNetSD_T2VBase(**kw)
    synth.fill_module_fast(model, seed=0)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    d = bench.make_host_inputs("t2v", hw, seed=11)
    x, y, cam = d["noise"].to(dev), d["y"].to(dev), d["cam"].to(dev)
    t = torch.tensor([981], device=dev)
    for name, autocast in (("B1 fp32 (TF32 conv default)", False), ("B1h fp16 autocast", True)):
        def fwd():
            with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
                return unet_oracle.unet_t2v_forward(sd, x, t, y, cam)
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            fwd()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"row": name, "latent": [4, 24, hw, hw], "ms_per_unet_forward_b1": ms,
                          "frames_per_s_50step_cfg": 24 / (ms * 100 / 1e3)}), flush=True)


if __name__ == "__main__":
    main()
