"""End-to-end parity on the GPU: drop-in module (CUDA path through the C ABI) vs the reference's golden outputs and the
CPU oracle, for T2V and I2V, eager and CUDA-graph, single and CFG-batched."""
import json
import os

import numpy as np
import pytest
import torch

from tests.test_oracle_cpu import load_case, run_oracle

pytestmark = pytest.mark.gpu

# End-to-end tolerance.  Per-kernel parity is held to rtol 1e-3 / atol 1e-4 (tests/test_*_gpu.py).  Through ~60
# residual blocks with fp16 activation storage the fp32-reference distance is dominated by accumulated fp16 rounding
# (2^-11 per store); DESIGN.md section "Numerics" reports the measured drift.  Bounds below have ~3x headroom.
E2E_REL_L2 = 4e-3
E2E_MAX_ABS = 3e-2


_MODELS = {}


def build(meta, seed_w, share=False):
    """share=True: full-size models (1.4 B parameters, ~40 s of weight synthesis) are built once per (kind, seed)."""
    from videomv_b200 import synth, unet
    key = (meta["kind"], json.dumps(meta["kwargs"], sort_keys=True), seed_w)
    if share and key in _MODELS:
        return _MODELS[key]
    cls = unet.UNetSD_T2VBase if meta["kind"] == "t2v" else unet.UNetSD_I2VGen
    model = cls(**meta["kwargs"])
    sd = synth.synth_state_dict(meta["shapes"], seed=seed_w)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    if share:
        _MODELS.clear()                      # one full-size model resident at a time
        _MODELS[key] = (model, sd)
    return model, sd


def call(model, meta, d):
    kw = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())     # camera arrives on CPU like the reference
    if meta["kind"] == "i2v":
        kw.update(image=d["image"].cuda(), local_image=d["local_image"].cuda())
    return model(d["x"].cuda(), d["t"].cuda(), **kw)


def metrics(name, out, ref):
    out, ref = out.float().cpu(), ref.float().cpu()
    rel = ((out - ref).norm() / ref.norm()).item()
    mx = (out - ref).abs().max().item()
    print(f"[e2e] {name}: rel_l2={rel:.3e} max_abs={mx:.3e} max|ref|={ref.abs().max().item():.3f}")
    return rel, mx


@pytest.mark.parametrize("case", ["t2v_small", "t2v_small_t981_cam", "i2v_small"])
def test_small_models_match_reference_golden(case):
    meta, d, _ = load_case(case)
    model, sd = build(meta, meta["seed_w"])
    out = call(model, meta, d)
    assert out.shape == d["ref"].shape and out.dtype == torch.float32 and out.is_cuda
    rel, mx = metrics(case + " vs reference golden", out, d["ref"])
    assert rel < E2E_REL_L2 and mx < E2E_MAX_ABS
    rel2, _ = metrics(case + " vs oracle (CPU fp32)", out, run_oracle(meta, sd, d))
    assert rel2 < E2E_REL_L2


def test_full_size_config1_matches_reference_golden():
    meta, d, _ = load_case("t2v_config1")
    model, _ = build(meta, meta["seed_w"], share=True)
    out = call(model, meta, d)
    rel, mx = metrics("t2v_config1 (1.41B params, 1x4x4x32x32) vs reference golden", out, d["ref"])
    assert rel < E2E_REL_L2 and mx < E2E_MAX_ABS


# The benchmark shapes themselves (BASELINE configs 2 and 3): full width, 24 x 32 x 32 and 4 x 64 x 64 latents, orbit
# cameras, against outputs of the UNMODIFIED reference (oracle/gen_golden.py bench_t2v / bench_512).
@pytest.mark.parametrize("case", ["t2v_24x32", "t2v_4x64"])
def test_full_size_benchmark_shapes_match_reference_golden(case):
    meta, d, _ = load_case(case)
    model, _ = build(meta, meta["seed_w"], share=True)
    out = call(model, meta, d)
    rel, mx = metrics(f"{case} (benchmark shape) vs reference golden", out, d["ref"])
    assert rel < E2E_REL_L2 and mx < E2E_MAX_ABS
    model.enable_cuda_graphs(True)
    g1 = call(model, meta, d)
    g2 = call(model, meta, d)
    model.enable_cuda_graphs(False)
    assert torch.equal(g1, g2) and torch.equal(g1, out), "eager, graph capture and graph replay must be bit-identical"


def test_fp16_reference_drifts_as_far_as_we_do():
    """The north-star tolerance (rtol 1e-3 / atol 1e-4) is stated for fp16.  The reference's own fp16 mode is autocast
    (i2vgen_xl_infer.yaml use_fp16: True, inference_*_entrance.py:243-249).  Evidence that our distance to the fp32 golden
    is the fp16 storage floor and not an implementation difference: the reference arithmetic (oracle restatement, pinned
    to the unmodified reference at 2.6e-6) run under torch.autocast(fp16) on the same inputs lands at least as far from
    the fp32 golden as this implementation does (SURVEY.md section 7 'hard parts')."""
    from oracle import unet_oracle
    meta, d, _ = load_case("t2v_24x32")
    model, sd = build(meta, meta["seed_w"], share=True)
    ours = call(model, meta, d).float().cpu()
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref16 = unet_oracle.unet_t2v_forward(sd_cuda, d["x"].cuda(), d["t"].cuda(), d["y"].cuda(), d["cam"].cuda(),
                                             fps=d["fps"].cuda()).float().cpu()
    del sd_cuda
    gold = d["ref"]
    rel_ours = ((ours - gold).norm() / gold.norm()).item()
    rel_ref16 = ((ref16 - gold).norm() / gold.norm()).item()
    mx_ours, mx_ref16 = (ours - gold).abs().max().item(), (ref16 - gold).abs().max().item()
    print(f"[e2e] distance to the fp32 reference golden (24x32x32): ours rel_l2 {rel_ours:.3e} max {mx_ours:.3e} | "
          f"reference arithmetic under fp16 autocast rel_l2 {rel_ref16:.3e} max {mx_ref16:.3e}")
    assert rel_ours <= rel_ref16 * 1.05 and mx_ours <= mx_ref16 * 1.5


def test_full_size_i2v_config1_matches_reference_golden():
    meta, d, _ = load_case("i2v_config1")
    model, _ = build(meta, meta["seed_w"], share=True)
    out = call(model, meta, d)
    rel, mx = metrics("i2v_config1 vs reference golden", out, d["ref"])
    assert rel < E2E_REL_L2 and mx < E2E_MAX_ABS


def test_full_size_i2v_benchmark_shape_matches_reference_golden():
    meta, d, _ = load_case("i2v_24x32")                       # BASELINE config 4
    model, _ = build(meta, meta["seed_w"], share=True)
    out = call(model, meta, d)
    rel, mx = metrics("i2v_24x32 (benchmark shape) vs reference golden", out, d["ref"])
    assert rel < E2E_REL_L2 and mx < E2E_MAX_ABS
    _MODELS.clear()


def test_new_prompt_is_not_served_from_a_stale_cache():
    """The reference engines build fresh y tensors per caption (inference_text2video_entrance.py:170,240); after sample A's
    tensors die, CPython and the caching allocator hand the SAME id() / address to sample B's tensors.  The
    step-invariant caches must not answer B with A's context."""
    import gc
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, _ = build(meta, meta["seed_w"])
    cold, _ = build(meta, meta["seed_w"])
    x, t = d["x"].cuda(), d["t"].cuda()
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randn(d["y"].shape, generator=g) for _ in range(6)]
    uncond = torch.randn(d["y"].shape, generator=g).cuda()
    seen = set()
    for i, p in enumerate(prompts):
        y = p.cuda()                                               # fresh tensor per prompt, freed at the end of the iteration
        seen.add((id(y), y.data_ptr()))
        kw_c = dict(y=y, camera_data=d["cam"], fps=d["fps"].cuda())
        kw_u = dict(y=uncond, camera_data=d["cam"], fps=d["fps"].cuda())
        out = model(x, t, **kw_c)
        yo, _ = model.forward_cfg_pair(x, t, kw_c, kw_u)
        cold.invalidate_engine()                                   # a model that has never seen another prompt
        want = cold(x, t, y=p.cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
        assert torch.equal(out, want), f"prompt {i}: forward() answered from a stale conditioning cache"
        assert metrics(f"prompt {i} cfg pair vs cold model", yo, want)[0] < SELF_REL_L2
        del y, kw_c, kw_u, out, yo
        gc.collect()


# Two fp16 evaluations of the same function (different batch => different tiling / split-K => different fp32 summation
# order => different fp16 roundings) differ by about sqrt(2) x their individual distance to the fp32 truth.
SELF_REL_L2 = 6e-3


def test_graph_replay_and_cfg_pair_match_oracle_and_eager():
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, sd = build(meta, meta["seed_w"])
    eager = call(model, meta, d)
    g = torch.Generator().manual_seed(9)
    y_u = torch.randn(d["y"].shape, generator=g)
    kw_c = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(y=y_u.cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    ref_c = run_oracle(meta, sd, d)
    ref_u = run_oracle(meta, sd, dict(d, y=y_u))
    # CFG pair: one batch-2 evaluation; each half must match the fp32 oracle as well as a batch-1 call does
    yo, uo = model.forward_cfg_pair(d["x"].cuda(), d["t"].cuda(), kw_c, kw_u)
    assert metrics("cfg pair cond vs oracle", yo, ref_c)[0] < E2E_REL_L2
    assert metrics("cfg pair uncond vs oracle", uo, ref_u)[0] < E2E_REL_L2
    assert metrics("cfg pair cond vs eager B=1", yo, eager)[0] < SELF_REL_L2
    # graphs: same kernels in the same order, and every reduction on the path has a fixed order (GroupNorm: per-CTA slots
    # summed in CTA order; LayerNorm: per-tile slots merged in slot order; no floating-point atomics), so eager runs,
    # graph capture and graph replays are bit-identical.
    assert torch.equal(call(model, meta, d), eager)
    model.enable_cuda_graphs(True)
    g1 = call(model, meta, d)
    g2 = call(model, meta, d)                     # replay
    assert torch.equal(g1, eager) and torch.equal(g2, eager)
    assert metrics("graph replay vs oracle", g2, ref_c)[0] < E2E_REL_L2
    x2 = d["x"].cuda() * 0.5
    e3 = model(x2, d["t"].cuda(), **kw_c)
    model.enable_cuda_graphs(False)
    e4 = model(x2, d["t"].cuda(), **kw_c)
    assert torch.equal(e3, e4)
    assert model.graph_launches() > 100


def test_sampler_matches_reference_golden_with_fake_model():
    from oracle import ddim_oracle
    from videomv_b200.sampler import DiffusionDDIM
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddim_fake.npz"))
    noise, yc, yu = (torch.from_numpy(z[k]).cuda() for k in ("noise", "yc", "yu"))
    fm = lambda x, t, **kw: ddim_oracle.fake_model(x, t, y=kw["y"])
    for mean_type, gs in (("eps", 9.0), ("v", 6.0)):
        s = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120),
                          mean_type=mean_type, var_type="fixed_small")
        out = s.ddim_sample_loop(noise.clone(), fm, model_kwargs=[dict(y=yc), dict(y=yu)], guide_scale=gs,
                                 ddim_timesteps=50, eta=0.0)
        ref = torch.from_numpy(z[mean_type])
        assert torch.allclose(out.cpu(), ref, rtol=2e-4, atol=2e-4), (out.cpu() - ref).abs().max()


def test_sampler_with_unet_pair_equals_two_call_loop():
    from videomv_b200.sampler import DiffusionDDIM
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, _ = build(meta, meta["seed_w"])
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(1, 4, 24, 8, 8, generator=g).cuda()
    kw_c = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(y=torch.randn(d["y"].shape, generator=g).cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    s = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120))
    a = s.ddim_sample_loop(noise, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, batch_cfg=False)
    b = s.ddim_sample_loop(noise, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, batch_cfg=True)
    assert torch.isfinite(a).all()
    # 10 chained steps with guidance 9 amplify the per-call fp16 differences
    assert metrics("sampler pair vs two-call", b, a)[0] < 3e-2


def test_reference_sampler_drives_our_module():
    """The drop-in contract end to end: the reference's OWN DiffusionDDIM (tools/modules/diffusions/diffusion_ddim.py,
    unmodified -- staged by oracle/stage_ref.py for the GPU box) calls our module exactly as inference.py does
    (:149-155: two calls per step with autoencoder=None, the four fp64 schedule tables, y / fps / camera_data kwargs), and the
    result equals our sampler's two-call loop (same model calls; fused CFG + DDIM kernel vs the reference's elementwise ops)."""
    from oracle import ref_import
    from videomv_b200.sampler import DiffusionDDIM
    if not ref_import.available():
        pytest.skip("reference sources not reachable (oracle/_ref not staged)")
    RefDDIM = ref_import.load_reference_ddim()
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, _ = build(meta, meta["seed_w"])
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(1, 4, 24, 8, 8, generator=g).cuda()
    kw_c = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(y=torch.randn(d["y"].shape, generator=g).cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    sp = dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120, zero_terminal_snr=False)
    ref = RefDDIM(schedule="linear_sd", schedule_param=sp, mean_type="eps", var_type="fixed_small", loss_type="mse")
    with torch.no_grad():
        a = ref.ddim_sample_loop(noise.clone(), model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, eta=0.0)
    ours = DiffusionDDIM(schedule="linear_sd", schedule_param=sp, mean_type="eps", var_type="fixed_small")
    b = ours.ddim_sample_loop(noise.clone(), model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, batch_cfg=False)
    assert a.shape == noise.shape and torch.isfinite(a).all()
    rel = metrics("reference DiffusionDDIM driving our UNet vs our sampler (two-call)", a, b)[0]
    # the two update arithmetics differ in the last fp32 bits; through 10 chained UNet calls with guidance 9 that is
    # amplified to the fp16 noise floor of the network (same bound as test_sampler_with_unet_pair_equals_two_call_loop)
    assert rel < 3e-2


def test_whole_sample_loop_as_one_cuda_graph():
    """the 10-step guided loop captured as ONE CUDA graph equals the step-by-step loop bit for
    bit (same kernels, same order) on every replay, also with new noise and after unrelated allocations in between."""
    from videomv_b200.sampler import DiffusionDDIM
    meta, d, _ = load_case("t2v_small_t981_cam")
    model, _ = build(meta, meta["seed_w"])
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(1, 4, 24, 8, 8, generator=g).cuda()
    kw_c = dict(y=d["y"].cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    kw_u = dict(y=torch.randn(d["y"].shape, generator=g).cuda(), camera_data=d["cam"], fps=d["fps"].cuda())
    s = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120))
    a = s.ddim_sample_loop(noise, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10)
    keep = []
    for _ in range(6):                                # replays, with allocator churn in between (the graph owns all it reads)
        keep.append(torch.randn(1 << 16, device="cuda"))
        r = s.ddim_sample_loop(noise, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, loop_graph=True)
        assert torch.equal(a, r)
    n2 = noise * 0.5
    a2 = s.ddim_sample_loop(n2, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10)
    c2 = s.ddim_sample_loop(n2, model, model_kwargs=[kw_c, kw_u], guide_scale=9.0, ddim_timesteps=10, loop_graph=True)
    assert torch.equal(a2, c2)
    assert len(model.__dict__["_loop_graphs"]) == 1
