"""torchrun worker: ONE sample sharded over P ranks (pure frame sharding, and CFG split x frame sharding; NVLink peer-memory
exchange, NCCL baseline at P = 2) must equal the single-GPU forward -- eager and CUDA-graph, forward() and forward_cfg_pair().
Launched by tests/test_sharded_gpu.py (needs >= 2 GPUs) and by hand at P = 4 / 8 (profiles/r2_sharded_check_n*.log)."""
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

    z = torch.ones(1024, device="cuda") * (rank + 1)
    dist.all_reduce(z)
    torch.cuda.synchronize()
    say(f"all_reduce ok ({z[0].item()})")
    meta, d, _ = load_case("t2v_small_t981_cam")                  # 24 frames, 8x8 latent -> 1x1 at the deepest level
    world = dist.get_world_size()
    # the deepest level (latent / 8) must have >= (ranks of a frame group) pixels: 16x16 latent (-> 2x2) up to 4 ranks, 32x32 for 8
    hw = 32 if world > 4 else 16
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 4, 24, hw, hw, generator=g).cuda()
    y_u = torch.randn(d["y"].shape, generator=g).cuda()
    model, _ = build(meta, meta["seed_w"])
    kw = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(kw, y=y_u)
    ref = model(x, d["t"].cuda(), **kw)
    ref_c, ref_u = model.forward_cfg_pair(x, d["t"].cuda(), kw, kw_u)
    torch.cuda.synchronize()
    say("single-GPU reference forward + CFG pair done")
    ok = torch.tensor([1], device="cuda")
    outs = {}
    default = "peer:frames,peer:cfgframes" + (",gather:frames" if world == 2 else "")
    for item in os.environ.get("VMV_CHECK_EXCHANGES", default).split(","):
        # "peer": ONE kernel per exchange over NVLink peer memory (csrc/peer.cu); "gather": the NCCL baseline
        # "frames": pure frame sharding; "cfgframes": the CFG pair split over two rank groups x frame sharding inside each
        exch, _, how = item.partition(":")
        cfg_split = how == "cfgframes"
        if cfg_split and world % 2:
            continue
        model.enable_cuda_graphs(False)
        model.set_frame_sharding(exchange=exch, cfg_split=cfg_split)
        out = model(x, d["t"].cuda(), **kw)
        oc, ou = model.forward_cfg_pair(x, d["t"].cuda(), kw, kw_u)
        torch.cuda.synchronize()
        rel = ((out - ref).norm() / ref.norm()).item()
        relp = max(((oc - ref_c).norm() / ref_c.norm()).item(), ((ou - ref_u).norm() / ref_u.norm()).item())
        sh = model._engine().shard
        print(f"[sharded] rank {rank}/{world} [{item}: {sh.describe()}]: rel_l2 vs single-GPU {rel:.3e} (forward) {relp:.3e} (CFG pair), "
              f"{sh.fused_ops} exchanges fused into GEMM epilogues + {sh.peer_ops} peer-memory kernels + {sh.collectives} NCCL collectives per forward", flush=True)
        ok *= 1 if (rel < 6e-3 and relp < 6e-3) else 0
        outs[item] = out
        # graphs with the captured exchanges
        try:
            model.enable_cuda_graphs(True)
            og = model(x, d["t"].cuda(), **kw)
            for _ in range(3):
                og2 = model(x, d["t"].cuda(), **kw)
            gc_, gu_ = model.forward_cfg_pair(x, d["t"].cuda(), kw, kw_u)
            gc_, gu_ = model.forward_cfg_pair(x, d["t"].cuda(), kw, kw_u)
            torch.cuda.synchronize()
            relg = ((og2 - ref).norm() / ref.norm()).item()
            relgp = max(((gc_ - ref_c).norm() / ref_c.norm()).item(), ((gu_ - ref_u).norm() / ref_u.norm()).item())
            same = bool(torch.equal(og, og2)) and bool(torch.equal(og2, out))
            print(f"[sharded] rank {rank} [{item}]: graph replay rel_l2 {relg:.3e} (forward) {relgp:.3e} (CFG pair); "
                  f"eager == capture == replay bit-identical: {same}", flush=True)
            ok *= 1 if (relg < 6e-3 and relgp < 6e-3 and same) else 0
        except Exception as e:  # noqa: BLE001
            print(f"[sharded] rank {rank} [{item}]: graph capture failed: {e!r}", flush=True)
            ok *= 0
        model.enable_cuda_graphs(False)
        model._engine()._graphs.clear()
        torch.cuda.synchronize()
        # every rank must hold the same output (it is gathered, not recomputed)
        chk = out.clone()
        dist.broadcast(chk, src=0)
        ok *= 1 if torch.equal(chk, out) else 0
    if len(outs) >= 2:
        vals = list(outs.items())
        for (na, a_), (nb_, b_) in zip(vals[:-1], vals[1:]):
            print(f"[sharded] rank {rank}: {na} vs {nb_} rel_l2 {((a_ - b_).norm() / ref.norm()).item():.3e}", flush=True)
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
