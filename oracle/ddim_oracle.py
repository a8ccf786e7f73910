"""TEST INFRASTRUCTURE ONLY -- literal restatement of the reference DDIM sampler arithmetic.

Follows tools/modules/diffusions/diffusion_ddim.py:138-160 (`p_mean_variance` guidance combine), :193-199 (eps / v ->
x0), :211-244 (`ddim_sample`) and :247-260 (`ddim_sample_loop`) for the shipped inference settings (fixed_small
variance, no clamp, condition_fn=None, autoencoder=None).  Pinned against the reference class by oracle/gen_golden.py
(`ddim_fake` fixture).
"""
import torch


def _i(tensor, t, x):
    """diffusion_ddim.py:9-15"""
    shape = (x.size(0),) + (1,) * (x.ndim - 1)
    return tensor.to(x.device)[t].view(shape).to(x)


class DDIMOracle:
    def __init__(self, num_timesteps=1000, init_beta=0.00085, last_beta=0.0120, mean_type="eps"):
        betas = torch.linspace(init_beta ** 0.5, last_beta ** 0.5, num_timesteps, dtype=torch.float64) ** 2   # schedules.py:40
        self.num_timesteps = num_timesteps
        self.mean_type = mean_type
        self.alphas_cumprod = torch.cumprod(1 - betas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)

    @torch.no_grad()
    def ddim_sample_loop(self, noise, model, model_kwargs, guide_scale, ddim_timesteps=50, eta=0.0):
        b = noise.size(0)
        xt = noise
        stride = self.num_timesteps // ddim_timesteps
        steps = (1 + torch.arange(0, self.num_timesteps, stride)).clamp(0, self.num_timesteps - 1).flip(0)
        for step in steps:
            t = torch.full((b,), int(step), dtype=torch.long, device=xt.device)
            y_out = model(xt, t, **model_kwargs[0])
            u_out = model(xt, t, **model_kwargs[1])
            out = u_out + guide_scale * (y_out - u_out)
            if self.mean_type == "eps":
                x0 = _i(self.sqrt_recip_alphas_cumprod, t, xt) * xt - _i(self.sqrt_recipm1_alphas_cumprod, t, xt) * out
            else:
                x0 = _i(self.sqrt_alphas_cumprod, t, xt) * xt - _i(self.sqrt_one_minus_alphas_cumprod, t, xt) * out
            eps = (_i(self.sqrt_recip_alphas_cumprod, t, xt) * xt - x0) / _i(self.sqrt_recipm1_alphas_cumprod, t, xt)
            alphas = _i(self.alphas_cumprod, t, xt)
            alphas_prev = _i(self.alphas_cumprod, (t - stride).clamp(0), xt)
            sigmas = eta * torch.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
            direction = torch.sqrt(1 - alphas_prev - sigmas ** 2) * eps
            xt = torch.sqrt(alphas_prev) * x0 + direction
        return xt


def fake_model(x, t, y=None, **kw):
    """Closed-form stand-in for the UNet used by the sampler fixtures (bounded, depends on x, t and y)."""
    return 0.3 * torch.tanh(x) + 0.05 * y.mean() + 0.1 * torch.sin(t.float() / 100.0).view(-1, 1, 1, 1, 1)
