"""DDIM scheduler used when `diffusers` is not installed (it is absent from this image).

The reference loads `diffusers.DDIMScheduler.from_pretrained(model_key, subfolder="scheduler")`
(/root/reference/elastic_diffusion.py:153, diffusers pinned to 0.21.4 by environment.yaml:21).  The product only needs
the noise schedule tables and `set_timesteps` / `add_noise`; the DDIM *step* arithmetic itself (eta = 0, epsilon
prediction, no clipping) runs inside the fused CUDA epilogue (csrc/epilogue.cu) from the scalars produced by
`step_scalars`.  Attribute surface mirrors diffusers' so that a real `DDIMScheduler` can be passed instead.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class DDIMSchedule:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 set_alpha_to_one=False):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule="scaled_linear", clip_sample=False,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type="epsilon", timestep_spacing="leading", thresholding=False)
        root = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32)
        self.betas = root ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        stride = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * stride).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        ts = timesteps.to(original_samples.device)
        a = (ac[ts] ** 0.5).flatten()
        b = ((1 - ac[ts]) ** 0.5).flatten()
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise


def check_scheduler(s) -> None:
    """The fused epilogue hard-codes DDIM(eta=0, epsilon, no clipping / thresholding); refuse anything else loudly."""
    cfg = s.config
    get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    if get("prediction_type", "epsilon") != "epsilon":
        raise ValueError("libelastic_b200's DDIM epilogue implements epsilon prediction only")
    if get("clip_sample", False) or get("thresholding", False):
        raise ValueError("libelastic_b200's DDIM epilogue does not implement clip_sample / thresholding")
    for attr in ("betas", "alphas_cumprod", "final_alpha_cumprod"):
        if not hasattr(s, attr):
            raise ValueError(f"scheduler lacks `{attr}`; a DDIM scheduler is required (reference ed:153)")


def step_scalars(s, t):
    """fp32 scalars of one DDIM step at timestep `t`, computed with the same fp32 tensor expressions diffusers uses
    (so the floats are bit-identical to what the reference multiplies with)."""
    t = int(t)
    prev = t - s.config.num_train_timesteps // s.num_inference_steps
    a_t = s.alphas_cumprod[t]
    a_p = s.alphas_cumprod[prev] if prev >= 0 else s.final_alpha_cumprod
    b_t = 1 - a_t
    return dict(sqrt_beta_t=float(b_t ** 0.5), sqrt_alpha_t=float(a_t ** 0.5), sqrt_alpha_prev=float(a_p ** 0.5),
                sqrt_dir=float((1 - a_p - (0.0 * ((1 - a_p) / b_t * (1 - a_t / a_p)) ** 0.5) ** 2) ** 0.5))


def renoise_scalars(s, t_next):
    """Coefficients of undo_step (ed:692-704): x <- (1-beta)^0.5 x + beta^0.5 eps for t_next .. t_next+n-1."""
    n = s.config.num_train_timesteps // s.num_inference_steps
    a, b = [], []
    for i in range(n):
        beta = s.betas[int(t_next) + i]
        a.append(float((1 - beta) ** 0.5))
        b.append(float(beta ** 0.5))
    return a, b
