"""Debug: is the whole-loop CUDA graph idempotent?  Compares eager step loops and graph replays of the 10-step guided loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_oracle_cpu import load_case  # noqa: E402
from tests.test_unet_gpu import build  # noqa: E402
from videomv_b200.sampler import DiffusionDDIM  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, _ = build(meta, meta["seed_w"])
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(1, 4, 24, 8, 8, generator=g).cuda()
    kw_c = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(y=torch.randn(d["y"].shape, generator=g).cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    s = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120))
    run = lambda **kw: s.ddim_sample_loop(noise, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=steps, **kw)
    eager = [run() for _ in range(4)]
    print("eager runs equal:", [bool(torch.equal(eager[0], e)) for e in eager[1:]])
    model.enable_cuda_graphs(True)
    pg = [run() for _ in range(4)]
    print("per-step graphs equal eager:", [bool(torch.equal(eager[0], e)) for e in pg])
    model.enable_cuda_graphs(False)
    lg = [run(loop_graph=True) for _ in range(8)]
    print("loop graph replays equal eager:", [bool(torch.equal(eager[0], e)) for e in lg])
    print("loop graph replays equal replay 1:", [bool(torch.equal(lg[0], e)) for e in lg])
    print("loop graph replays equal replay 2:", [bool(torch.equal(lg[1], e)) for e in lg])
    print("rel diff replay1 vs replay2:", float((lg[0] - lg[1]).norm() / lg[0].norm()))


if __name__ == "__main__":
    main()
