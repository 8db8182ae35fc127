"""Oracle restatement of the slice of ``diffusers.DDIMScheduler`` the reference calls.

TEST INFRASTRUCTURE.  PARITY UNPINNED (see oracle/__init__.py): diffusers is a
third-party dependency absent from /root/reference and from this image; the
reference pins no version.  API usage (``step(..., variance_noise=)``,
``.to_tuple()``, ``config.clip_sample = ...``) dates it to diffusers ~0.14-0.21.

Reference call sites this must satisfy:
  scheduler.step            src/diffusion_utils.py:100-107
  scheduler.set_timesteps   src/base_diffusion.py:60,113 ; src/ddpm_inversion.py:96
  scheduler.add_noise       src/ddpm_inversion.py:74
  alphas_cumprod / final_alpha_cumprod / timesteps / num_inference_steps /
  config.num_train_timesteps / config.clip_sample
                            src/diffusion_utils.py:18-22,79-80,117
                            src/ddim_inversion.py:23-38 ; src/attr_functions.py:148

All tables are float32 CPU tensors, and every per-step scalar derived from them
stays a float32 0-d tensor - this defines the rounding of the coefficients.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, fields
from types import SimpleNamespace

import numpy as np
import torch


class BaseOutput(OrderedDict):
    """Dataclass-friendly ordered dict (``diffusers.utils.BaseOutput`` behaviour the
    reference relies on: attribute access, ``to_tuple()``, integer indexing -
    src/SegDiffEditPipeline.py:33-37, src/diffusion_utils.py:107, src/metrics.py:99)."""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                OrderedDict.__setitem__(self, f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            OrderedDict.__setitem__(self, name, value)
        super().__setattr__(name, value)

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


@dataclass
class DDIMSchedulerOutput(BaseOutput):
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor | None = None


# Scheduler configurations of the three hub models the reference loads
# (src/models.py:21,38,48); values from the public model cards.
SCHEDULER_PRESETS = {
    "ddpm": dict(beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0),
    "ldm": dict(beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear",
                clip_sample=False, set_alpha_to_one=True, steps_offset=0),
    "sd": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
               clip_sample=False, set_alpha_to_one=False, steps_offset=1),
}


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02,
                 beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
                 steps_offset=0, clip_sample_range=1.0, prediction_type="epsilon"):
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is on the reference's path")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start,
            beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
            clip_sample_range=clip_sample_range, prediction_type=prediction_type)
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps,
                                        dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5,
                                        num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = (torch.tensor(1.0) if set_alpha_to_one
                                    else self.alphas_cumprod[0])
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(
            np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_preset(cls, name, **overrides):
        kw = dict(SCHEDULER_PRESETS[name])
        kw.update(overrides)
        return cls(**kw)

    @classmethod
    def from_config(cls, config, **kw):
        if isinstance(config, SimpleNamespace):
            return cls(**vars(config))
        raise NotImplementedError("hub ids cannot be resolved offline")

    def set_timesteps(self, num_inference_steps, device=None):
        """'leading' spacing: arange(T) * (N // T), reversed, + steps_offset."""
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict=True):
        """DDIM eq. 12 / 16 (arXiv:2010.02502)."""
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** (0.5) * model_output) / a_t ** (0.5)
        eps = model_output
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** (0.5)
        if use_clipped_model_output:
            eps = (sample - a_t ** (0.5) * x0) / b_t ** (0.5)
        direction = (1 - a_p - std_dev_t ** 2) ** (0.5) * eps
        prev_sample = a_p ** (0.5) * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator,
                                             dtype=model_output.dtype).to(model_output.device)
            prev_sample = prev_sample + std_dev_t * variance_noise
        if not return_dict:
            return (prev_sample, x0)
        return DDIMSchedulerOutput(prev_sample=prev_sample, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
