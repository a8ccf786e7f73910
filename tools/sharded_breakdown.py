"""torchrun worker: where one sharded UNet call spends its device time.  Every rank runs one instrumented CFG-pair forward
in the headline multi-GPU mode (CFG split x frame sharding, `--frames` for pure frame sharding), then ALL ranks replay each
distinct launch shape of every kernel family in lockstep from CUDA graphs (peer kernels rendezvous with their peers, so the
figures include the exchange latency).  Rank 0 prints per-family totals and the per-shape tables of the peer families."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from videomv_b200 import ops, synth, unet  # noqa: E402
from videomv_b200.profiling import family_shape_times  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    with torch.device(dev):
        model = unet.UNetSD_T2VBase(**bench.T2V_KWARGS)
    synth.fill_module_fast(model, seed=0)
    model.eval()
    model.set_frame_sharding(cfg_split="--frames" not in sys.argv)
    host = bench.make_host_inputs("t2v", 32, seed=11)
    kw = bench.to_kwargs("t2v", host, dev)
    x = host["noise"].to(dev)
    t = torch.full((1,), 981, dtype=torch.long, device=dev)
    run = lambda: model.forward_cfg_pair(x, t, kw[0], kw[1])
    for _ in range(2):
        run()
    ops.PROFILE = []
    torch.cuda.synchronize()
    dist.barrier()
    run()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    sh = model._engine().shard
    fams = ["gemm_tc", "groupnorm", "groupnorm_peer", "attention", "peer_exchange", "peer_gather"]
    out = {}
    for f in fams:
        dist.barrier()
        out[f] = family_shape_times(prof, f)
    # whole call from a graph
    model.enable_cuda_graphs(True)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"# {sh.describe()} on {world} GPUs: device time of one UNet call (CFG pair) on rank 0, graph replay of each distinct shape")
        acc = 0.0
        for f in fams:
            rows = out[f]
            ms = sum(n * us for _, n, _, us, _ in rows) / 1e3
            acc += ms
            print(f"| {f} | {sum(r[1] for r in rows)} launches | {ms:.3f} ms |")
        print(f"| sum of families | | {acc:.3f} ms |\n| whole call, graph replay (all ranks in lockstep) | {model.graph_launches()} kernels | {total:.3f} ms |")
        print("\ngemm_tc launches that feed a layout exchange (res1 + mode 1 conv2 / mode 0 proj_out -> pixels; mode 2 / mode 0 -> frames):")
        for desc, n, fl, us, by in sorted(out["gemm_tc"], key=lambda r: r[0]):
            if "scatter" in desc or (" res1 " in desc and ("mode1" in desc or "mode2" in desc or "mode0" in desc) and "act0" in desc):
                print(f"| {desc} | {n} | {us:.1f} us | {n * us / 1e3:.3f} ms |")
        for f in ("peer_exchange", "groupnorm_peer", "peer_gather"):
            print(f"\n{f}:")
            for desc, n, fl, us, by in sorted(out[f], key=lambda r: -r[1] * r[3]):
                print(f"| {desc} | {n} | {us:.1f} us | {n * us / 1e3:.3f} ms | {by / us / 1e3:.0f} GB/s |")
    dist.barrier()
    model.enable_cuda_graphs(False)
    model._engine()._graphs.clear()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
