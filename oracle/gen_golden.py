"""TEST INFRASTRUCTURE ONLY -- pin the oracle against the reference and write fixtures.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

For each case: build the unmodified reference UNet (oracle/ref_import.py), load the
seeded synthetic weights (videomv_b200/synth.py), run the reference forward on CPU fp32,
run the oracle restatement on the same state_dict/inputs, assert they agree, and store
  tests/golden/<case>.npz   : inputs, reference output, and a weight checksum
  tests/golden/<case>.json  : {param name: shape} of the reference state_dict + kwargs
Weights are NOT stored (5 GB at full size); tests regenerate them from the recipe.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, unet_oracle          # noqa: E402
from videomv_b200 import synth                      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _weight_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


def make_case(name, kind, kwargs, frames, hw, seed_w, seed_x, t_value, real_cam=False):
    T2V, I2V = ref_import.load_reference()
    torch.manual_seed(0)
    cls = T2V if kind == "t2v" else I2V
    if kind == "i2v":                                   # unet_i2vgen.py:334 hard-codes .cuda()
        torch.Tensor.cuda = lambda self, *a, **k: self
    model = cls(**kwargs).eval()
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed=seed_w)
    missing = model.load_state_dict(sd, strict=True)
    in_dim = kwargs["in_dim"]
    x, t, y, cam = synth.synth_inputs(1, frames, hw, hw, in_dim=in_dim, seed=seed_x, t_value=t_value)
    if real_cam:
        cam = synth.orbit_cameras(frames)
    extra = {}
    g = torch.Generator().manual_seed(seed_x + 100)
    fps = torch.tensor([8], dtype=torch.long)
    t0 = time.time()
    with torch.no_grad():
        if kind == "t2v":
            ref = model(x, t, y=y, camera_data=cam, fps=fps)
        else:
            image = torch.randn(1, 1, 1024, generator=g)
            local_image = torch.randn(1, 4, 1, hw, hw, generator=g).repeat(1, 1, frames, 1, 1) * 0.18215 * 5
            extra = dict(image=image, local_image=local_image)
            ref = model(x, t, y=y, camera_data=cam, fps=fps, image=image, local_image=local_image)
    t_ref = time.time() - t0
    t0 = time.time()
    if kind == "t2v":
        out = unet_oracle.unet_t2v_forward(sd, x, t, y, cam, fps=fps)
    else:
        out = unet_oracle.unet_i2v_forward(sd, x, t, y, extra["image"], extra["local_image"], cam, fps=fps)
    t_or = time.time() - t0
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"[{name}] ref {t_ref:.1f}s oracle {t_or:.1f}s  max|oracle-ref|={err:.3e}  max|ref|={scale:.3f} "
          f"std={ref.std().item():.3f}", flush=True)
    assert err <= 2e-5 * max(1.0, scale), f"oracle does not match the reference on {name}"
    # input sensitivity sanity (SURVEY section 4 trap 1)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        x=x.numpy(), t=t.numpy(), y=y.numpy().astype(np.float32), cam=cam.numpy(), fps=fps.numpy(),
                        ref=ref.numpy(), weight_checksum=np.float64(_weight_checksum(sd)),
                        **{k: v.numpy() for k, v in extra.items()})
    with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
        json.dump(dict(kind=kind, kwargs=kwargs, frames=frames, hw=hw, seed_w=seed_w, seed_x=seed_x,
                       oracle_vs_ref_maxabs=err, shapes=shapes), f)
    del model


def make_ddim_case():
    """Pin oracle/ddim_oracle.py against the reference DiffusionDDIM (imported unmodified) with a closed-form model."""
    import importlib.util
    from oracle import ddim_oracle
    ref_import.load_reference()          # sets sys.path / stubs
    pk = "tools.modules.diffusions"
    import types
    if pk not in sys.modules:
        m = types.ModuleType(pk); m.__path__ = [os.path.join(ref_import.REF_ROOT, "tools/modules/diffusions")]
        sys.modules[pk] = m
    for name in ("schedules", "losses", "diffusion_ddim"):
        spec = importlib.util.spec_from_file_location(f"{pk}.{name}", os.path.join(ref_import.REF_ROOT, "tools/modules/diffusions", name + ".py"))
        mod = importlib.util.module_from_spec(spec); sys.modules[f"{pk}.{name}"] = mod; spec.loader.exec_module(mod)
    RefDDIM = sys.modules[f"{pk}.diffusion_ddim"].DiffusionDDIM
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(1, 4, 6, 8, 8, generator=g)
    yc, yu = torch.randn(1, 77, 16, generator=g), torch.randn(1, 77, 16, generator=g)
    out = {}
    for mean_type, gs in (("eps", 9.0), ("v", 6.0)):
        ref = RefDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120,
                      zero_terminal_snr=False), mean_type=mean_type, var_type="fixed_small", loss_type="mse")
        fm = lambda x, t, **kw: ddim_oracle.fake_model(x, t, y=kw["y"])
        r = ref.ddim_sample_loop(noise.clone(), fm, model_kwargs=[dict(y=yc), dict(y=yu)], guide_scale=gs,
                                 ddim_timesteps=50, eta=0.0)
        o = ddim_oracle.DDIMOracle(mean_type=mean_type).ddim_sample_loop(noise.clone(), fm, [dict(y=yc), dict(y=yu)], gs, 50)
        err = (r - o).abs().max().item()
        print(f"[ddim_fake {mean_type}] max|oracle-ref|={err:.3e} max|ref|={r.abs().max().item():.3f}")
        assert err < 1e-5
        out[mean_type] = r.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "ddim_fake.npz"), noise=noise.numpy(), yc=yc.numpy(), yu=yu.numpy(),
                        eps=out["eps"], v=out["v"])


def make_vae_case(name, n, hw, seed_w, seed_z):
    """Pin oracle/vae_oracle.py against the reference AutoencoderKL.decode (imported unmodified) and store the fixture."""
    from oracle import vae_oracle
    VAE = ref_import.load_reference_vae()
    torch.manual_seed(0)
    model = VAE(**ref_import.VAE_KWARGS).eval()
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed=seed_w)
    model.load_state_dict(sd, strict=True)
    z = torch.randn(n, 4, hw, hw, generator=torch.Generator().manual_seed(seed_z)) * 5.0     # latents / scale_factor: std ~ 5
    t0 = time.time()
    with torch.no_grad():
        ref = model.decode(z)
    t_ref = time.time() - t0
    out = vae_oracle.vae_decode(sd, z)
    err = (out - ref).abs().max().item()
    print(f"[{name}] ref {t_ref:.1f}s  max|oracle-ref|={err:.3e}  max|ref|={ref.abs().max().item():.3f} std={ref.std().item():.3f}", flush=True)
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), f"oracle does not match the reference on {name}"
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), z=z.numpy(), ref=ref.numpy().astype(np.float32),
                        weight_checksum=np.float64(_weight_checksum(sd)))
    with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
        json.dump(dict(kind="vae", kwargs=ref_import.VAE_KWARGS, n=n, hw=hw, seed_w=seed_w, seed_z=seed_z,
                       oracle_vs_ref_maxabs=err, shapes=shapes), f)


def main():
    which = sys.argv[1:] or ["small", "full"]
    if "vae" in which:
        make_vae_case("vae_small", n=2, hw=8, seed_w=21, seed_z=22)          # 2 x 64 x 64 images
        make_vae_case("vae_256", n=1, hw=32, seed_w=21, seed_z=23)           # one 256 x 256 frame (BASELINE config 2's frames)
    if "ddim" in which:
        make_ddim_case()
    if "small" in which:
        make_case("t2v_small", "t2v", ref_import.SMALL_KWARGS, frames=4, hw=16, seed_w=3, seed_x=1, t_value=500)
        make_case("t2v_small_t981_cam", "t2v", ref_import.SMALL_KWARGS, frames=24, hw=8, seed_w=3, seed_x=2,
                  t_value=981, real_cam=True)
        make_case("i2v_small", "i2v", ref_import.SMALL_I2V_KWARGS, frames=4, hw=16, seed_w=4, seed_x=3, t_value=1)
    if "full" in which:
        make_case("t2v_config1", "t2v", ref_import.T2V_KWARGS, frames=4, hw=32, seed_w=7, seed_x=1, t_value=500)
    if "full_i2v" in which:
        make_case("i2v_config1", "i2v", ref_import.I2V_KWARGS, frames=4, hw=32, seed_w=8, seed_x=5, t_value=501)
    # the benchmark shapes themselves (BASELINE configs 2, 3, 4): full width, 24 x 32 x 32 / 4 x 64 x 64 latents
    if "bench_t2v" in which:
        make_case("t2v_24x32", "t2v", ref_import.T2V_KWARGS, frames=24, hw=32, seed_w=7, seed_x=11, t_value=981, real_cam=True)
    if "bench_i2v" in which:
        make_case("i2v_24x32", "i2v", ref_import.I2V_KWARGS, frames=24, hw=32, seed_w=8, seed_x=12, t_value=501, real_cam=True)
    if "bench_512" in which:
        make_case("t2v_4x64", "t2v", ref_import.T2V_KWARGS, frames=4, hw=64, seed_w=7, seed_x=13, t_value=241, real_cam=True)


if __name__ == "__main__":
    main()
