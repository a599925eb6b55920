"""TEST INFRASTRUCTURE - CPU oracle, not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this.

Restatement of `diffusers==0.21.4` `schedulers/scheduling_ddim.py::DDIMScheduler` (pinned by
/root/reference/environment.yaml:21; the package is NOT vendored in the reference and NOT installed in this
image, so this follows the published algorithm, "Denoising Diffusion Implicit Models", Song et al. 2021, eq. 12
with eta = 0, and the public Stable-Diffusion scheduler configs).  PARITY UNPINNED against diffusers itself: no
copy of diffusers exists offline and the reference has no tests / golden vectors; the
restatement is pinned only through the reference's own call sites:

  from_pretrained            elastic_diffusion.py:153
  scale_model_input          :402   (identity for DDIM)
  add_noise                  :358
  config.num_train_timesteps :693 ; num_inference_steps :693 ; betas :699
  set_timesteps              :1001 ; timesteps :1013,1038,1040
  step -> ['prev_sample'], ['pred_original_sample']   :776-780, 920-921, 1033-1035, 1054-1056

Config (all three BASELINE model families): beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012,
num_train_timesteps=1000, clip_sample=False, set_alpha_to_one=False, steps_offset=1,
prediction_type="epsilon", timestep_spacing="leading".
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class _Out(dict):
    __getattr__ = dict.__getitem__


class DDIMRestated:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 set_alpha_to_one=False):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule="scaled_linear", clip_sample=False,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type="epsilon", timestep_spacing="leading")
        # scaled_linear: linear in sqrt(beta)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps          # "leading" spacing
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def step(self, model_output, timestep, sample, eta=0.0):
        prev_t = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5                 # eq. 12, "predicted x_0"
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output               # "direction pointing to x_t"
        prev = a_prev ** 0.5 * x0 + direction
        return _Out(prev_sample=prev, pred_original_sample=x0)
