"""torchrun worker: frame-sharded forward (P ranks, NCCL) must equal the single-GPU forward.
Launched by tests/test_sharded_gpu.py (needs >= 2 GPUs)."""
import datetime
import os
import sys

os.environ.setdefault("TORCH_NCCL_BLOCKING_WAIT", "1")       # a stuck collective raises after the timeout instead of hanging
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_oracle_cpu import load_case  # noqa: E402
from tests.test_unet_gpu import build  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=90))

    def say(msg):
        print(f"[sharded] rank {rank}: {msg}", flush=True)

    # which NCCL primitives work on this box (diagnostic; the engine uses all_gather + all_reduce by default)
    z = torch.ones(1024, device="cuda") * (rank + 1)
    dist.all_reduce(z)
    torch.cuda.synchronize()
    say(f"all_reduce ok ({z[0].item()})")
    g = torch.empty(dist.get_world_size() * 1024, device="cuda")
    dist.all_gather_into_tensor(g, z)
    torch.cuda.synchronize()
    say("all_gather_into_tensor ok")
    meta, d, _ = load_case("t2v_small_t981_cam")                  # 24 frames, 8x8 latent -> 1x1 at the deepest level
    world = dist.get_world_size()
    # deepest level must have >= world pixels: use a 16x16 latent (-> 2x2) for 2..4 ranks
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 4, 24, 16, 16, generator=g).cuda()
    model, _ = build(meta, meta["seed_w"])
    kw = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    ref = model(x, d["t"].cuda(), **kw)
    torch.cuda.synchronize()
    say("single-GPU reference forward done")
    ok = torch.tensor([1], device="cuda")
    outs = {}
    for exch in os.environ.get("VMV_CHECK_EXCHANGES", "peer,gather").split(","):
        # "peer": ONE kernel per exchange over NVLink peer memory (csrc/peer.cu); "gather": the NCCL baseline
        model.enable_cuda_graphs(False)
        model.set_frame_sharding(exchange=exch)
        out = model(x, d["t"].cuda(), **kw)
        torch.cuda.synchronize()
        rel = ((out - ref).norm() / ref.norm()).item()
        sh = model._engine().shard
        print(f"[sharded] rank {rank}/{world} [{exch}]: rel_l2 vs single-GPU {rel:.3e}, {sh.peer_ops} peer-memory kernels + "
              f"{sh.collectives} NCCL collectives per forward", flush=True)
        ok *= 1 if rel < 6e-3 else 0
        outs[exch] = out
        # graphs with the captured exchanges
        try:
            model.enable_cuda_graphs(True)
            og = model(x, d["t"].cuda(), **kw)
            for _ in range(3):
                og2 = model(x, d["t"].cuda(), **kw)
            torch.cuda.synchronize()
            relg = ((og2 - ref).norm() / ref.norm()).item()
            print(f"[sharded] rank {rank} [{exch}]: graph replay rel_l2 {relg:.3e}", flush=True)
            ok *= 1 if relg < 6e-3 else 0
        except Exception as e:  # noqa: BLE001
            print(f"[sharded] rank {rank} [{exch}]: graph capture failed: {e!r}", flush=True)
            ok *= 0
        model.enable_cuda_graphs(False)
        model._engine()._graphs.clear()
        torch.cuda.synchronize()
    if len(outs) == 2:
        a_, b_ = list(outs.values())
        print(f"[sharded] rank {rank}: peer vs gather rel_l2 {((a_ - b_).norm() / ref.norm()).item():.3e}", flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.barrier()
    code = 0 if int(ok.item()) == 1 else 1
    # CUDA graphs that captured NCCL kernels must die before the communicator does
    model.enable_cuda_graphs(False)
    model._engine()._graphs.clear()
    torch.cuda.synchronize()
    if os.environ.get("VMV_TRY_A2A", "0") == "1":      # diagnostic, last: a timeout here may poison the communicator
        try:
            snd = torch.ones(2 * 1024, device="cuda")
            rcv = torch.empty_like(snd)
            dist.all_to_all_single(rcv, snd)
            torch.cuda.synchronize()
            say("all_to_all_single ok")
        except Exception as e:  # noqa: BLE001
            say(f"all_to_all_single FAILED: {e!r}")
    sys.stdout.flush()
    os._exit(code)          # skip NCCL teardown: destroy_process_group hung here on the sandboxed box
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
