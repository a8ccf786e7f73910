"""Host-side mirror of the reference sampler, the *caller* of the UNet hot path.

`DiffusionDDIM` keeps the constructor and `ddim_sample_loop` call contract of
tools/modules/diffusions/diffusion_ddim.py:19-69,246-260 for the inference configuration the shipped YAMLs use
(classifier-free guidance with two kwargs dicts, eta = 0, `fixed_small` variance, eps- or v-prediction).  It stays
Python/PyTorch glue as the north star asks; the only device work it adds is ONE fused kernel per step
(`vmv_cfg_ddim_step`: guidance combine + x0 + DDIM update) instead of the reference's ~12 elementwise launches, and -
when the model is one of ours - it evaluates the cond/uncond pair as a single batch-2 UNet call (`forward_cfg_pair`),
which is arithmetically identical per sample and halves weight traffic and launches.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops


def linear_sd_betas(num_timesteps: int = 1000, init_beta: float = 0.00085, last_beta: float = 0.0120) -> torch.Tensor:
    """tools/modules/diffusions/schedules.py:40-41 `linear_sd_schedule` (float64)."""
    return torch.linspace(init_beta ** 0.5, last_beta ** 0.5, num_timesteps, dtype=torch.float64) ** 2


class DiffusionDDIM:
    def __init__(self, schedule: str = "linear_sd", schedule_param: Optional[Dict] = None, mean_type: str = "eps",
                 var_type: str = "fixed_small", loss_type: str = "mse", epsilon: float = 1e-12,
                 rescale_timesteps: bool = False, noise_strength: float = 0.0, **kwargs):
        sp = dict(schedule_param or {})
        if schedule != "linear_sd":
            raise NotImplementedError("videomv_b200.sampler: only the 'linear_sd' schedule of the shipped configs")
        if sp.get("zero_terminal_snr", False):
            raise NotImplementedError("videomv_b200.sampler: zero_terminal_snr is not used by the shipped configs")
        if mean_type not in ("eps", "v"):
            raise NotImplementedError("videomv_b200.sampler: mean_type must be 'eps' (T2V) or 'v' (I2V)")
        if not var_type.startswith("fixed"):
            raise NotImplementedError("videomv_b200.sampler: learned variance is not used by the shipped configs")
        betas = linear_sd_betas(sp.get("num_timesteps", 1000), sp.get("init_beta", 0.00085), sp.get("last_beta", 0.0120))
        self.betas = betas
        self.num_timesteps = len(betas)
        self.mean_type, self.var_type = mean_type, var_type
        self.rescale_timesteps = rescale_timesteps
        alphas = 1 - betas
        self.alphas_cumprod = torch.cumprod(alphas, dim=0)                       # diffusion_ddim.py:52-53
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)

    def ddim_steps(self, ddim_timesteps: int) -> torch.Tensor:
        """diffusion_ddim.py:253: [981, 961, ..., 1] for 50 steps over 1000."""
        stride = self.num_timesteps // ddim_timesteps
        return (1 + torch.arange(0, self.num_timesteps, stride)).clamp(0, self.num_timesteps - 1).flip(0)

    def step_coefficients(self, ddim_timesteps: int, guide_scale: float) -> torch.Tensor:
        """[n_steps, 7] fp32 rows for vmv_cfg_ddim_step (see include/videomv_b200.h)."""
        stride = self.num_timesteps // ddim_timesteps
        rows = []
        for step in self.ddim_steps(ddim_timesteps).tolist():
            ac = self.alphas_cumprod[step]
            ac_prev = self.alphas_cumprod[max(step - stride, 0)]                 # :236 (t - stride).clamp(0)
            c_recip, c_recipm1 = torch.sqrt(1.0 / ac), torch.sqrt(1.0 / ac - 1)
            if self.mean_type == "eps":
                kx, ko = c_recip, c_recipm1                                      # :193-195
            else:
                kx, ko = torch.sqrt(ac), torch.sqrt(1.0 - ac)                    # :196-199
            rows.append([kx, ko, c_recip, c_recipm1, torch.sqrt(ac_prev), torch.sqrt(1 - ac_prev), guide_scale])
        return torch.tensor(rows, dtype=torch.float64).to(torch.float32)

    @torch.no_grad()
    def ddim_sample_loop(self, noise, model, autoencoder=None, model_kwargs=None, clamp=None, percentile=None,
                         condition_fn=None, guide_scale=None, ddim_timesteps: int = 20, eta: float = 0.0,
                         batch_cfg: bool = True, loop_graph: bool = False):
        """Same call as diffusion_ddim.py:247 (first / plain pass: autoencoder=None).  loop_graph=True (or VMV_LOOP_GRAPH=1)
        with one of this repo's UNets: the whole guided loop is captured once as ONE CUDA graph and replayed."""
        if autoencoder is not None or condition_fn is not None or clamp is not None or percentile is not None:
            raise NotImplementedError("videomv_b200.sampler: LGM refine pass / classifier guidance / clamping are "
                                      "outside the accelerated path (SURVEY.md section 8f)")
        if eta != 0.0:
            raise NotImplementedError("videomv_b200.sampler: eta must be 0 (deterministic DDIM), as in the shipped configs")
        xt = noise.contiguous().float()
        b = xt.size(0)
        dev = xt.device
        steps = self.ddim_steps(ddim_timesteps)
        if guide_scale is None:
            kw_c = kw_u = dict(model_kwargs or {})
            gs = 1.0
        else:
            assert isinstance(model_kwargs, list) and len(model_kwargs) == 2      # :147
            kw_c, kw_u = model_kwargs
            gs = float(guide_scale)
        coef = self.step_coefficients(ddim_timesteps, gs).to(dev)
        t_all = steps.to(device=dev, dtype=torch.long)
        pair = getattr(model, "forward_cfg_pair", None) if (batch_cfg and guide_scale is not None) else None
        if pair is not None and (loop_graph or os.environ.get("VMV_LOOP_GRAPH", "0") == "1") and hasattr(model, "cfg_sample_loop_graph"):
            return model.cfg_sample_loop_graph(xt, t_all, coef, kw_c, kw_u)
        tables = dict(sqrt_alphas_cumprod=self.sqrt_alphas_cumprod,
                      sqrt_one_minus_alphas_cumprod=self.sqrt_one_minus_alphas_cumprod,
                      sqrt_recip_alphas_cumprod=self.sqrt_recip_alphas_cumprod,
                      sqrt_recipm1_alphas_cumprod=self.sqrt_recipm1_alphas_cumprod)
        for i in range(len(steps)):
            t = t_all[i].expand(b).contiguous()
            if pair is not None:
                y_out, u_out = pair(xt, t, kw_c, kw_u)
            elif guide_scale is None:
                y_out = u_out = model(xt, t, autoencoder=None, **tables, **kw_c).float().contiguous()
            else:
                y_out = model(xt, t, autoencoder=None, **tables, **kw_c).float().contiguous()     # :149-151
                u_out = model(xt, t, autoencoder=None, **tables, **kw_u).float().contiguous()     # :153-155
            xt = ops.cfg_ddim_step(xt, y_out, u_out, coef[i])
        return xt
