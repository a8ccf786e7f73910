"""CPU: the oracle restatement reproduces the committed golden vectors (which oracle/gen_golden.py produced by running
the UNMODIFIED reference in the build container), and the synthetic-weight recipe is reproducible."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ddim_oracle, unet_oracle
from videomv_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_case(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "weight_checksum"}, float(z["weight_checksum"])


def run_oracle(meta, sd, d):
    if meta["kind"] == "t2v":
        return unet_oracle.unet_t2v_forward(sd, d["x"], d["t"], d["y"], d["cam"], fps=d["fps"])
    return unet_oracle.unet_i2v_forward(sd, d["x"], d["t"], d["y"], d["image"], d["local_image"], d["cam"], fps=d["fps"])


@pytest.mark.parametrize("case", ["t2v_small", "t2v_small_t981_cam", "i2v_small"])
def test_oracle_matches_reference_golden(case):
    meta, d, checksum = load_case(case)
    sd = synth.synth_state_dict(meta["shapes"], seed=meta["seed_w"])
    got = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    assert abs(got - checksum) <= 1e-9 * checksum, "synthetic weight recipe is not reproducible on this machine"
    out = run_oracle(meta, sd, d)
    err = (out - d["ref"]).abs().max().item()
    assert err <= 2e-5 * max(1.0, d["ref"].abs().max().item()), err


@pytest.mark.parametrize("case", ["vae_small", "vae_256"])
def test_vae_oracle_matches_reference_golden(case):
    from oracle import vae_oracle
    meta, d, checksum = load_case(case)
    sd = synth.synth_state_dict(meta["shapes"], seed=meta["seed_w"])
    got = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    assert abs(got - checksum) <= 1e-9 * checksum
    out = vae_oracle.vae_decode(sd, d["z"])
    assert out.shape == d["ref"].shape
    err = (out - d["ref"]).abs().max().item()
    assert err <= 2e-5 * max(1.0, d["ref"].abs().max().item()), err


def test_oracle_is_input_dependent():
    meta, d, _ = load_case("t2v_small")
    sd = synth.synth_state_dict(meta["shapes"], seed=meta["seed_w"])
    d2 = dict(d, x=d["x"] + 0.1)
    assert (run_oracle(meta, sd, d2) - d["ref"]).abs().max().item() > 1e-3       # SURVEY section 4 trap 1


def test_ddim_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "ddim_fake.npz"))
    noise, yc, yu = (torch.from_numpy(z[k]) for k in ("noise", "yc", "yu"))
    fm = lambda x, t, **kw: ddim_oracle.fake_model(x, t, y=kw["y"])
    for mean_type, gs in (("eps", 9.0), ("v", 6.0)):
        out = ddim_oracle.DDIMOracle(mean_type=mean_type).ddim_sample_loop(noise.clone(), fm, [dict(y=yc), dict(y=yu)], gs, 50)
        assert torch.allclose(out, torch.from_numpy(z[mean_type]), rtol=1e-5, atol=1e-5)


def test_sampler_tables_match_oracle():
    from videomv_b200.sampler import DiffusionDDIM
    s = DiffusionDDIM(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.0120))
    o = ddim_oracle.DDIMOracle()
    assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
    assert s.ddim_steps(50).tolist() == list(range(981, 0, -20))
    c = s.step_coefficients(50, 9.0)
    assert c.shape == (50, 7) and torch.isfinite(c).all()


def test_orbit_cameras_shape_and_determinism():
    a, b = synth.orbit_cameras(24), synth.orbit_cameras(24)
    assert a.shape == (1, 24, 16) and torch.equal(a, b)
    m = a.reshape(24, 4, 4)
    assert torch.allclose(m[:, 3], torch.tensor([0., 0., 0., 1.]).expand(24, 4))
