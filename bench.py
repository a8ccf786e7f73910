#!/usr/bin/env python
"""bench.py -- VideoMV headline benchmark on B200: multi-view frames/s of the UNet denoising hot path.

metric   : multi-view frames / second = 24 * samples / time of `ddim_sample_loop` (50 DDIM steps, eta 0,
           classifier-free guidance => 100 UNet evaluations + 50 scheduler updates per sample; autoencoder=None).
           SURVEY.md section 8(d), BASELINE.json `metric`.
step     : ONE full 50-step sample of one prompt (24 frames).
           N > 1 GPUs (default, BASELINE config 5): ONE sample spread over the N GPUs => "scaling": "strong".  The CFG pair is
           split over two rank groups and each group shards its 24 frames (videomv_b200/parallel.py: layout exchange at the
           temporal segments + GroupNorm-statistic all-reduce + one output all-gather per UNet call, all over NVLink peer
           memory).  The same line also carries `modes`: pure frame sharding (frames/N) and the reference's own multi-GPU
           mode, N independent replicas (tools/inferences/inference_text2video_entrance.py:79,152-170; no collective).
           `--parallel replicas|frames|cfgframes` makes one of them the headline instead.
value    : device-resident inputs, CUDA-event timed, max over ranks.
e2e      : the same loop through the reference-facing API (module registry class + DiffusionDDIM.ddim_sample_loop) with
           HOST (pinned) inputs: H2D of noise / text embeddings / cameras and D2H of the final latent inside the timed
           region, every step.
roofline : the tcgen05 GEMM / implicit-conv kernel family (97% of the FLOPs): algorithmic FLOPs / CUDA-event time of
           every launch of one CFG-batched forward, vs MEASURED_PEAKS.json bf16 sustained TFLOP/s.
cpu_baseline / --impl reference : the reference's own UNet classes on the host cores (unmodified files staged by
           oracle/stage_ref.py; the oracle port oracle/unet_oracle.py when they are absent), bounded sample, extrapolated
           to the same metric.
reference_cuda : (N=1) the reference's own eager-PyTorch CUDA path on this GPU -- stock classes, xformers mapped to SDPA, fp32
           as t2v_infer.yaml ships and fp16 autocast -- a few UNet forwards, extrapolated x100: what the north star's
           ">= 1.8x" is measured against.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload t2v256|t2v512|i2v256]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Resolved UNet kwargs: tools/modules/config.py:88-106 overlaid by configs/t2v_infer.yaml:20-41.
T2V_KWARGS = dict(in_dim=4, dim=320, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2, 4, 4], num_heads=8,
                  head_dim=64, num_res_blocks=2, attn_scales=[1.0, 0.5, 0.25], dropout=0.1, misc_dropout=0.4,
                  temporal_attention=True, temporal_attn_times=1, use_checkpoint=True, use_fps_condition=False,
                  use_camera_condition=True, use_lgm_refine=True, use_sim_mask=False, upper_len=128, default_fps=8)
WORKLOADS = {
    # name: (kind, latent hw, guide_scale, mean_type, algorithmic TFLOP per forward [SURVEY 8d])
    "t2v256": ("t2v", 32, 9.0, "eps", 7.467),
    "t2v512": ("t2v", 64, 9.0, "eps", 34.26),
    "i2v256": ("i2v", 32, 6.0, "v", 7.51),
}
FRAMES = 24
DDIM_STEPS = 50


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1406.8))),
                    tflops_burst=float(d.get("bf16_tflops", 1674.6)), hbm=float(d.get("hbm_gbs", 6579.0)), src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def pick_cpu_threads(run_once) -> int:
    """The oracle is many small torch ops: on a many-core host the full thread count can be SLOWER (oversubscription).
    Probe a few counts on the bounded sample and keep the fastest -- i.e. the best the host can do."""
    cores = os.cpu_count() or 1
    best_n, best_t = cores, None
    for n in sorted({min(cores, 16), min(cores, 32), min(cores, 64), cores}):
        torch.set_num_threads(n)
        t0 = time.time()
        run_once()
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best_n, best_t = n, dt
        elif dt > 1.5 * best_t:
            break                                   # getting worse with more threads: stop probing
    torch.set_num_threads(best_n)
    return best_n


def make_host_inputs(kind: str, hw: int, seed: int):
    """Synthetic inputs exactly as SURVEY.md 8(d): pinned HOST tensors."""
    from videomv_b200 import synth
    g = torch.Generator().manual_seed(seed)
    pin = lambda z: z.pin_memory() if torch.cuda.is_available() else z
    d = dict(noise=pin(torch.randn(1, 4, FRAMES, hw, hw, generator=g)),
             y=pin(torch.randn(1, 77, 1024, generator=g)), y_neg=pin(torch.randn(1, 77, 1024, generator=g)),
             cam=synth.orbit_cameras(FRAMES), fps=torch.tensor([8], dtype=torch.long))
    if kind == "i2v":
        d["image"] = pin(torch.randn(1, 1, 1024, generator=g))
        d["local_image"] = pin((torch.randn(1, 4, 1, hw, hw, generator=g) * 0.18215).repeat(1, 1, FRAMES, 1, 1).contiguous())
    return d


def to_kwargs(kind: str, d: dict, dev):
    cam = d["cam"]                                     # stays on CPU like the reference engine (moved inside forward)
    cond = dict(y=d["y"].to(dev, non_blocking=True), camera_data=cam, fps=d["fps"].to(dev))
    unc = dict(y=d["y_neg"].to(dev, non_blocking=True), camera_data=cam, fps=d["fps"].to(dev))
    if kind == "i2v":
        li = d["local_image"].to(dev, non_blocking=True)
        cond.update(image=d["image"].to(dev, non_blocking=True), local_image=li)
        unc.update(image=torch.zeros_like(d["image"]).to(dev), local_image=li)   # use_zero_infer (i2vgen entrance :128,268)
    return [cond, unc]


# --------------------------------------------------------------------------------------------------------------
def _ref_kwargs(kind):
    kw = dict(T2V_KWARGS, use_lgm_refine=False)
    return dict(kw, concat_dim=4) if kind == "i2v" else kw


def _shapes(kind):
    from videomv_b200 import unet
    cls = unet.UNetSD_T2VBase if kind == "t2v" else unet.UNetSD_I2VGen
    with torch.device("meta"):
        return {k: tuple(v.shape) for k, v in cls(**_ref_kwargs(kind)).state_dict().items()}


def make_cpu_reference(kind, hw):
    """Returns (fwd(frames) -> seconds, kind string, weight-generation seconds): one UNet forward of the reference on the
    host cores -- the reference's own class when its files are reachable, else the oracle port."""
    from oracle import ref_import, unet_oracle
    from videomv_b200 import synth
    t0 = time.time()
    sd = synth.synth_state_dict(_shapes(kind), seed=0)
    t_w = time.time() - t0
    d = make_host_inputs(kind, hw, seed=11)
    model = None
    if ref_import.available():
        T2V, I2V = ref_import.load_reference()
        if kind == "i2v":
            torch.Tensor.cuda = lambda self, *a, **k: self           # unet_i2vgen.py:334 hard-codes .cuda(); CPU arm only
        model = (T2V if kind == "t2v" else I2V)(**_ref_kwargs(kind)).eval()
        model.load_state_dict(sd, strict=True)

    @torch.no_grad()
    def fwd(frames):
        x = d["noise"][:, :, :frames].contiguous()
        t = torch.tensor([981])
        cam = d["cam"][:, :frames]
        t0 = time.time()
        if model is not None:
            kw = dict(y=d["y"], camera_data=cam, fps=d["fps"])
            if kind == "i2v":
                kw.update(image=d["image"], local_image=d["local_image"][:, :, :frames])
            model(x, t, **kw)
        elif kind == "t2v":
            unet_oracle.unet_t2v_forward(sd, x, t, d["y"], cam)
        else:
            unet_oracle.unet_i2v_forward(sd, x, t, d["y"], d["image"], d["local_image"][:, :, :frames], cam, fps=d["fps"])
        return time.time() - t0

    return fwd, ("reference" if model is not None else "port"), t_w


def run_reference(args, kind, hw, tflop_fwd):
    """CPU arm: the reference UNet on the host cores, bounded sample, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fwd, ref_kind, t_w = make_cpu_reference(kind, hw)
    # bounded sample: probe with 4 frames; use full 24-frame forwards only if the whole run stays within ~4 minutes
    threads = pick_cpu_threads(lambda: fwd(4))
    t4 = fwd(4)
    n = args.steps + args.warmup
    frames = FRAMES if t4 * 6 * n < 240 else 4
    scale = FRAMES / frames
    times = [fwd(frames) for _ in range(n)][args.warmup:]
    t_fwd = sum(times) / len(times) * scale                      # one 24-frame UNet forward
    t_sample = t_fwd * 2 * DDIM_STEPS
    value = FRAMES / t_sample
    what = "the reference's own UNetSD class (unmodified files)" if ref_kind == "reference" else "oracle port"
    sample = (f"{len(times)} forward(s) of {what} on 1x4x{frames}x{hw}x{hw} (fp32, best of 16/32/64/{cores} threads = {threads}), "
              f"scaled x{scale:g} to 24 frames, x100 to a 50-step CFG sample; weights generated in {t_w:.0f}s")
    line = {"impl": "reference", "metric": "multi-view frames/sec (24-view, 50-step DDIM, CFG)", "value": value,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_sample * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload, "frames": FRAMES, "latent": [4, FRAMES, hw, hw], "ddim_steps": DDIM_STEPS,
                       "guidance": "cfg"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "host_cores": cores, "kind": ref_kind, "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_cuda_leg(kind, hw, dev, state_dict):
    """The reference's eager-PyTorch CUDA path on this GPU: stock classes (oracle port when the files are absent),
    xformers.memory_efficient_attention mapped to F.scaled_dot_product_attention, cudnn.benchmark as the reference engine
    sets it (inference_text2video_entrance.py:83).  A few B=1 UNet forwards, extrapolated x100 to a 50-step CFG sample."""
    from oracle import ref_import, unet_oracle
    d = make_host_inputs(kind, hw, seed=11)
    x, y, cam, fps = d["noise"].to(dev), d["y"].to(dev), d["cam"].to(dev), d["fps"].to(dev)
    t = torch.tensor([981], device=dev)
    extra = {}
    if kind == "i2v":
        extra = dict(image=d["image"].to(dev), local_image=d["local_image"].to(dev))
    prev_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    model = None
    if ref_import.available():
        T2V, I2V = ref_import.load_reference()
        with torch.device(dev):
            model = (T2V if kind == "t2v" else I2V)(**_ref_kwargs(kind)).eval()
        model.load_state_dict(state_dict, strict=True)
    sd = None if model is not None else {k: v.detach() for k, v in state_dict.items()}
    out = {"impl": "the reference's own UNetSD class (unmodified files, stock eager PyTorch)" if model is not None
           else "oracle port of the reference on CUDA (reference files not reachable)"}

    @torch.no_grad()
    def fwd():
        if model is not None:
            return model(x, t, y=y, camera_data=cam, fps=fps, **extra)
        if kind == "t2v":
            return unet_oracle.unet_t2v_forward(sd, x, t, y, cam)
        return unet_oracle.unet_i2v_forward(sd, x, t, y, extra["image"], extra["local_image"], cam, fps=fps)

    for name, autocast in (("fp32", False), ("fp16_autocast", True)):
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            for _ in range(2):
                fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fwd()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name] = {"ms_per_unet_forward_b1": ms, "frames_per_s": FRAMES / (ms * 2 * DDIM_STEPS / 1e3)}
    torch.backends.cudnn.benchmark = prev_bench
    out["sample"] = "3 timed B=1 UNet forwards per precision after 2 warm-ups, x100 to a 50-step CFG sample (two calls per step, diffusion_ddim.py:149-155)"
    del model, sd
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------------------
def run_native(args, kind, hw, gs, mean_type, tflop_fwd):
    import torch.distributed as dist
    from videomv_b200 import ops, synth, unet
    from videomv_b200.sampler import DiffusionDDIM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    unet.register_with_reference()                       # no-op on the GPU box (reference tree absent)
    cls = unet.UNetSD_T2VBase if kind == "t2v" else unet.UNetSD_I2VGen
    kw = dict(T2V_KWARGS) if kind == "t2v" else dict(T2V_KWARGS, concat_dim=4)
    with torch.device(dev):
        model = cls(**kw)
    synth.fill_module_fast(model, seed=0)                # one checkpoint for every rank, like the reference's replicas
    model.eval()
    diffusion = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085,
                              last_beta=0.0120, zero_terminal_snr=False), mean_type=mean_type, var_type="fixed_small")
    # headline mode: one sample over all GPUs (BASELINE config 5); CFG split x frame sharding when the rank count is even
    head = args.parallel
    if head == "auto":
        head = "single" if world == 1 else ("cfgframes" if world % 2 == 0 else "frames")
    if world == 1:
        head = "single"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(mode, steps, warmup, with_e2e=True):
        """Time `steps` samples in `mode`; returns a dict (all ranks compute it, values are max over ranks)."""
        sharded = mode in ("frames", "cfgframes")
        model.enable_cuda_graphs(False)
        if sharded:
            model.set_frame_sharding(cfg_split=(mode == "cfgframes"))
        else:
            model.set_frame_sharding(enable=False)
        model.enable_cuda_graphs(not args.no_graphs)
        # yaml seed 11, + rank like the reference engine (:79); sharding works on ONE sample => same inputs everywhere
        host = make_host_inputs(kind, hw, seed=11 + (0 if sharded else rank))

        def sample_from_host():
            noise = host["noise"].to(dev, non_blocking=True)
            kwargs = to_kwargs(kind, host, dev)
            out = diffusion.ddim_sample_loop(noise, model, model_kwargs=kwargs, guide_scale=gs, ddim_timesteps=DDIM_STEPS,
                                             eta=0.0, batch_cfg=not args.two_call)
            return out.to("cpu", non_blocking=False)

        dev_noise = host["noise"].to(dev)
        dev_kwargs = to_kwargs(kind, host, dev)

        def sample_resident():
            return diffusion.ddim_sample_loop(dev_noise, model, model_kwargs=dev_kwargs, guide_scale=gs,
                                              ddim_timesteps=DDIM_STEPS, eta=0.0, batch_cfg=not args.two_call)

        for _ in range(warmup):
            sample_resident()
        if with_e2e:
            sample_from_host()
        eng = model._engine()
        n0 = ops.launch_count()
        r0 = sum(getattr(g, "replays", 0) for g in eng._graphs.values())
        clocks = ClockSampler(local)
        clocks.start()
        ms = timed(sample_resident, steps)
        clk = clocks.stop()
        eager_launches = ops.launch_count() - n0
        replays = sum(getattr(g, "replays", 0) for g in eng._graphs.values()) - r0
        per_replay = max([g.launches for g in eng._graphs.values()] or [0])
        ms_e2e = timed(sample_from_host, steps) if with_e2e else None
        nsamp = 1 if sharded else world                  # samples finished per step across the job
        sh = eng.shard
        if sharded:
            par = (f"{sh.describe()}: ONE sample on {world} GPUs" + (", cond / uncond halves of the CFG pair on two rank groups" if sh.cfg_ways > 1 else "")
                   + (f", 24 frames sharded {sh.world}-way inside a group (layout exchange at the temporal segments + GroupNorm-statistic all-reduce)" if sh.world > 1 else "")
                   + f"; {sh.fused_ops} layout exchanges inside GEMM epilogues + {sh.peer_ops} peer-memory kernels + {sh.collectives} NCCL collectives per UNet call")
        else:
            par = f"replicas x{world} (one sample per GPU, no collective)" if world > 1 else "single GPU"
        return dict(mode=mode, sharded=sharded, ms=ms, ms_e2e=ms_e2e, clk=clk, nsamp=nsamp, steps=steps, parallelism=par,
                    launches=int(eager_launches + replays * per_replay), host=host, dev_noise=dev_noise, dev_kwargs=dev_kwargs,
                    value=FRAMES * steps * nsamp / (ms / 1e3),
                    e2e_value=None if ms_e2e is None else FRAMES * steps * nsamp / (ms_e2e / 1e3))

    # the other multi-GPU modes first (short), the headline last so that the model is left configured for it
    extra_modes = {}
    if world > 1 and not args.quick and not args.no_modes:
        for m in ("replicas", "frames", "cfgframes"):
            if m == head or (m == "cfgframes" and world % 2):
                continue
            r = measure(m, 1, 1, with_e2e=False)
            extra_modes[m] = {"value": r["value"], "unit": "frames/s", "ms_per_sample": r["ms"], "scaling": "strong" if r["sharded"] else "weak",
                              "parallelism": r["parallelism"], "steps": 1, "warmup": 1}
    res = measure(head, args.steps, max(args.warmup, 3))
    sharded, ms, clk, host, dev_noise, dev_kwargs = res["sharded"], res["ms"], res["clk"], res["host"], res["dev_noise"], res["dev_kwargs"]
    value, e2e_value, launches = res["value"], res["e2e_value"], res["launches"]
    h2d = sum(v.numel() * v.element_size() for k, v in host.items() if k not in ("cam", "fps"))
    d2h = host["noise"].numel() * 4

    line = {"metric": "multi-view frames/sec (24-view, 50-step DDIM, CFG)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": args.workload, "model": cls.__name__ + " 1.41B (configs/t2v_infer.yaml)" if kind == "t2v"
                       else cls.__name__ + " (configs/i2vgen_xl_infer.yaml)",
                       "frames": FRAMES, "latent": [4, FRAMES, hw, hw], "ddim_steps": DDIM_STEPS,
                       "guidance": f"cfg {gs}, cond+uncond as one batch-2 UNet call" if not args.two_call else f"cfg {gs}, two calls",
                       "parallelism": res["parallelism"],
                       "cuda_graphs": not args.no_graphs,
                       "l2": "working set > L2: 2.83 GB of fp16 weights streamed per UNet call (no explicit flush)"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches)}
    if extra_modes:
        line["modes"] = extra_modes

    if rank == 0 and args.quick:
        print(json.dumps(line), flush=True)
    elif rank == 0:
        # ---- roofline leg: one instrumented (eager, per-launch CUDA events) CFG-batched forward
        pk = _peaks()
        model.enable_cuda_graphs(False)
        if model._engine().shard is not None:
            model.set_frame_sharding(enable=False)   # the instrumented forward below runs on rank 0 alone
        xt = dev_noise
        t = torch.full((1,), 981, dtype=torch.long, device=dev)
        for _ in range(2):
            model.forward_cfg_pair(xt, t, dev_kwargs[0], dev_kwargs[1])
        ops.PROFILE = []
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        model.forward_cfg_pair(xt, t, dev_kwargs[0], dev_kwargs[1])
        ev1.record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        fam = {}
        for name, fl, by, a, b, _desc, _replay in prof:
            f = fam.setdefault(name, [0, 0.0, 0.0, 0.0])
            f[0] += 1; f[1] += fl; f[2] += by; f[3] += a.elapsed_time(b)
        g = fam.get("gemm_tc", [0, 0.0, 0.0, 1e-9])
        # device time of every GEMM launch of that forward: each distinct shape replayed from a CUDA graph (the eager
        # event pairs above include host launch gaps, which exceed the kernel for the many small launches)
        from videomv_b200.profiling import gemm_shape_times
        rows = gemm_shape_times(prof)
        dev_ms = sum(n * us for _, n, _, us in rows) / 1e3
        if args.shapes_out:
            with open(args.shapes_out, "w") as f:
                f.write("| shape | n | us each | total ms | share | TFLOP/s |\n|---|---:|---:|---:|---:|---:|\n")
                for desc, n, fl, us in sorted(rows, key=lambda r: -r[1] * r[3]):
                    f.write(f"| {desc} | {n} | {us:.1f} | {n * us / 1e3:.3f} | {100 * n * us / (dev_ms * 1e3):.1f}% | {fl / us / 1e6:.0f} |\n")
                f.write(f"gemm total {dev_ms:.3f} ms per B=2 forward ({len(rows)} distinct shapes)\n")
        achieved = g[1] / (dev_ms * 1e-3) / 1e12
        # DRAM traffic of the same kernel family from the committed ncu capture of one forward (profiles/r2_traffic.json,
        # made by tools/summarize_traffic.py): bytes per launch, like `achieved` is FLOPs per launch / time per launch
        traffic, alg_bytes = None, sum(r_[2] for r_ in prof if r_[0] == "gemm_tc")
        tj = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.isfile(tj):
            try:
                tr = json.load(open(tj)).get("gemm_tc2_kernel")
                traffic = tr["dram_bytes_per_launch"] if tr else None
            except Exception:  # noqa: BLE001
                traffic = None
        # binding roofline per launch shape: max(FLOPs / tensor peak, algorithmic bytes / HBM peak), summed over the forward
        shape_bytes = {}
        for r_ in prof:
            if r_[0] == "gemm_tc":
                shape_bytes.setdefault(r_[5], r_[2])
        t_bind = sum(n * max(fl / (pk["tflops"] * 1e12), shape_bytes.get(desc, 0.0) / (pk["hbm"] * 1e9)) for desc, n, fl, _ in rows)
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tc2_kernel<BN,STAGES,.> (tcgen05 CTA-pair GEMM / implicit conv family)",
                            "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                            "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes / max(g[0], 1),
                            "frac_of_binding_roofline": t_bind / (dev_ms * 1e-3),
                            "peak_source": pk["src"] + " bf16_tflops_sustained",
                            "launches_per_forward": g[0], "distinct_shapes": len(rows),
                            "algorithmic_tflop_per_forward_b2": g[1] / 1e12, "kernel_ms_per_forward_b2": dev_ms,
                            "how": "algorithmic FLOPs of the 422 launches of one CFG-batched forward / their device time "
                                   "(CUDA events around graph replays of each distinct shape, L2-warm)",
                            "eager_event_ms_per_forward_b2": g[3]}
        # device time of the other kernel families, measured the same way (graph replay of every distinct shape)
        from videomv_b200.profiling import family_shape_times
        gn_rows = family_shape_times(prof, "groupnorm")
        at_rows = family_shape_times(prof, "attention")
        gn_ms = sum(n * us for _, n, _, us, _ in gn_rows) / 1e3
        at_ms = sum(n * us for _, n, _, us, _ in at_rows) / 1e3
        line["breakdown_ms_per_forward_b2"] = {"gemm_tc": round(dev_ms, 3), "groupnorm": round(gn_ms, 3), "attention": round(at_ms, 3),
                                               "graph_forward_total": round(ms / args.steps / DDIM_STEPS, 3),
                                               "how": "graph-replay device time per family; total = timed sample / 50 steps"}
        if gn_rows:
            by = sum(n * b_ for _, n, _, _, b_ in gn_rows)
            line["roofline_hbm_norms"] = {"bound": "hbm", "kernel": "gn_smem_kernel / gn_fused_kernel (GroupNorm+SiLU, 166 launches)",
                                          "achieved": by / (gn_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                          "frac": by / (gn_ms * 1e-3) / 1e9 / pk["hbm"],
                                          "algorithmic_bytes": "one fp16 read + one fp16 write of each normalised tensor"}
        if at_rows:
            fl = sum(n * f_ for _, n, f_, _, _ in at_rows)
            line["attention_tflops"] = {"achieved": fl / (at_ms * 1e-3) / 1e12, "launches": sum(r_[1] for r_ in at_rows)}
        t_fwd_alg = 2 * tflop_fwd                               # cond + uncond
        line["forward_tflops"] = {"algorithmic_tflop_per_step": t_fwd_alg * DDIM_STEPS,
                                  "achieved_tflops": t_fwd_alg * DDIM_STEPS * args.steps / (ms / 1e3) * 1.0,
                                  "frac_of_peak": t_fwd_alg * DDIM_STEPS * args.steps / (ms / 1e3) / pk["tflops"]}
        # ---- cpu baseline leg (N=1 only): bounded oracle sample on the host cores
        if world == 1 and not args.no_cpu_baseline:
            # ---- the same sample through the REFERENCE's own sampler class driving this module (its two-call CFG pattern,
            # diffusion_ddim.py:149-155, its elementwise update): the path a user of inference.py gets with only the registry
            # swap of INTEGRATION.md section 1 and no other change
            try:
                from oracle import ref_import
                if ref_import.available():
                    RefDDIM = ref_import.load_reference_ddim()
                    rd = RefDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120,
                                 zero_terminal_snr=False), mean_type=mean_type, var_type="fixed_small", loss_type="mse")
                    model.enable_cuda_graphs(not args.no_graphs)

                    def ref_loop():
                        noise = host["noise"].to(dev, non_blocking=True)
                        kwargs = to_kwargs(kind, host, dev)
                        with torch.no_grad():
                            return rd.ddim_sample_loop(noise, model, model_kwargs=kwargs, guide_scale=gs, ddim_timesteps=DDIM_STEPS, eta=0.0).to("cpu")
                    ref_loop()
                    torch.cuda.synchronize()
                    t0 = time.time()
                    ref_loop()
                    dt = time.time() - t0
                    line["e2e_reference_sampler"] = {"value": FRAMES / dt, "unit": "frames/s", "samples": 1,
                                                     "how": "the reference's unmodified DiffusionDDIM.ddim_sample_loop calling this module twice per step "
                                                            "(B=1 graphs), host inputs, wall clock"}
                    model.enable_cuda_graphs(False)
            except Exception as e:  # noqa: BLE001
                line["e2e_reference_sampler"] = {"error": repr(e)}
            # ---- the whole 50-step loop as ONE CUDA graph: same kernels, no per-step host work
            try:
                model.enable_cuda_graphs(False)
                loop = lambda: diffusion.ddim_sample_loop(dev_noise, model, model_kwargs=dev_kwargs, guide_scale=gs,
                                                          ddim_timesteps=DDIM_STEPS, eta=0.0, loop_graph=True)
                t0 = time.time()
                loop()
                torch.cuda.synchronize()
                t_build = time.time() - t0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    loop()
                e1.record()
                torch.cuda.synchronize()
                line["loop_graph"] = {"value": FRAMES * 2 / (e0.elapsed_time(e1) / 1e3), "unit": "frames/s", "samples": 2,
                                      "capture_s": round(t_build, 1), "kernels_in_graph": int(list(model.__dict__["_loop_graphs"].values())[0][4]),
                                      "how": "ddim_sample_loop(..., loop_graph=True): 50 steps x (CFG-batched UNet + fused CFG/DDIM update) in one graph"}
                model.__dict__["_loop_graphs"] = {}
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                line["loop_graph"] = {"error": repr(e)}
            # ---- the reference's own CUDA path on this GPU (bounded sample), then the CPU baseline
            try:
                rc = reference_cuda_leg(kind, hw, dev, model.state_dict())
                for k_ in ("fp32", "fp16_autocast"):
                    rc[k_]["speedup_of_this_repo_e2e"] = e2e_value / rc[k_]["frames_per_s"]
                line["reference_cuda"] = rc
            except Exception as e:  # noqa: BLE001
                line["reference_cuda"] = {"error": repr(e)}
            try:
                line["cpu_baseline"] = cpu_baseline(kind, hw)
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        model.enable_cuda_graphs(False)
        model._engine()._graphs.clear()              # graphs go before the communicator (peer mode captures no NCCL kernel)
        torch.cuda.synchronize()
        sys.stdout.flush()
        # the JSON line is out; a teardown that does not return (seen once on a sandboxed box) must not hang the job
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


def cpu_baseline(kind, hw):
    cores = os.cpu_count() or 1
    fwd, ref_kind, _ = make_cpu_reference(kind, hw)
    frames = 4                                                      # bounded sample: 4 of 24 frames, one forward
    threads = pick_cpu_threads(lambda: fwd(frames))
    best = min(fwd(frames), fwd(frames))
    t_sample = best * (FRAMES / frames) * 2 * DDIM_STEPS
    what = "the reference's own UNetSD class (unmodified files)" if ref_kind == "reference" else "oracle port"
    return {"value": FRAMES / t_sample, "unit": "frames/s", "cores": threads, "host_cores": cores, "kind": ref_kind,
            "sample": f"best of 2 UNet forwards of {what} on 1x4x{frames}x{hw}x{hw} fp32 ({best:.2f}s, best of 16/32/64/{cores} "
                      f"threads = {threads}), x{FRAMES // frames} frames x100 calls"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="t2v256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--two-call", action="store_true", help="cond and uncond as two B=1 UNet calls (reference call pattern)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shapes-out", default="", help="write the per-shape GEMM device-time table of the roofline leg here")
    ap.add_argument("--quick", action="store_true", help="A/B runs: value + e2e only (no roofline leg, no cpu baseline)")
    ap.add_argument("--parallel", default="auto", choices=["auto", "replicas", "frames", "cfgframes"],
                    help="N>1 headline mode.  auto = cfgframes (ONE sample on all GPUs: CFG pair split over two rank groups x "
                         "frame sharding inside each; strong scaling, BASELINE config 5); 'frames' = pure frame sharding; "
                         "'replicas' = one independent sample per GPU (weak scaling, the reference's own mode)")
    ap.add_argument("--no-modes", action="store_true", help="N>1: skip the short measurements of the non-headline modes")
    args = ap.parse_args()
    kind, hw, gs, mean_type, tflop = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, kind, hw, tflop)
    else:
        run_native(args, kind, hw, gs, mean_type, tflop)


if __name__ == "__main__":
    main()
