"""Native DDIM scheduler with the attribute / method surface the reference uses from
``diffusers.DDIMScheduler`` (src/diffusion_utils.py:18-22,79-80,100-107,117;
src/base_diffusion.py:60; src/ddpm_inversion.py:74,96).  Tables live on the host as fp32
tensors (as in diffusers); ``step`` enqueues one fused CUDA kernel."""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops
from .outputs import BaseOutput

PRESETS = {
    # google/ddpm-celebahq-256, CompVis/ldm-celebahq-256, CompVis/stable-diffusion-v1-4 scheduler configs
    "ddpm": dict(beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True,
                 set_alpha_to_one=True, steps_offset=0),
    "ldm": dict(beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear", clip_sample=False,
                set_alpha_to_one=True, steps_offset=0),
    "sd": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
               set_alpha_to_one=False, steps_offset=1),
}


@dataclass
class DDIMSchedulerOutput(BaseOutput):
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, clip_sample_range=1.0,
                 prediction_type="epsilon"):
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is supported")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
            steps_offset=steps_offset, clip_sample_range=clip_sample_range, prediction_type=prediction_type)
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                   dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"beta_schedule {beta_schedule!r}")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._coeff_cache = {}

    @classmethod
    def from_preset(cls, name: str, **overrides):
        kw = dict(PRESETS[name])
        kw.update(overrides)
        return cls(**kw)

    @classmethod
    def from_config(cls, config, **_):
        if isinstance(config, SimpleNamespace):
            return cls(**vars(config))
        if isinstance(config, dict):
            return cls(**config)
        raise NotImplementedError("hub ids cannot be resolved offline; pass a config namespace/dict")

    def set_timesteps(self, num_inference_steps: int, device=None):
        """'leading' spacing, kept on the HOST so the loop never synchronises on a timestep."""
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)
        self._coeff_cache.clear()

    # ------------------------------------------------------------------ host scalars
    def previous_timestep(self, t: int) -> int:
        return int(t) - self.config.num_train_timesteps // self.num_inference_steps

    def coeffs(self, t, eta: float = 0.0, mode: str = "ddim"):
        """Cached per-(t, eta, mode) fp32 coefficients (reference op order)."""
        key = (int(t), float(eta), mode, self.num_inference_steps)
        c = self._coeff_cache.get(key)
        if c is None:
            c = ops.step_coeffs(self.alphas_cumprod, self.final_alpha_cumprod, int(t),
                                self.previous_timestep(t), eta, mode)
            self._coeff_cache[key] = c
        return c

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[int(timestep)]
        a_p = self.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    # ------------------------------------------------------------------ device work
    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict=True):
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output is not on the reference's path")
        if eta > 0 and variance_noise is None:
            variance_noise = torch.randn(model_output.shape, generator=generator,
                                         dtype=model_output.dtype).to(model_output.device)
        prev, x0 = ops.guided_step(sample, model_output, self.coeffs(timestep, eta, "ddim"),
                                   clip=self.config.clip_sample, clip_range=self.config.clip_sample_range,
                                   noise=variance_noise if eta > 0 else None)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        ts = [int(t) for t in torch.as_tensor(timesteps).reshape(-1)]
        if len(ts) != 1:
            raise NotImplementedError("add_noise: one timestep per call (all the reference needs)")
        a = self.alphas_cumprod[ts[0]]
        # sa*x0 + sb*noise  ==  renoise with identity x0 extraction (c_a = 1, c_b = 0)
        return ops.axpby(original_samples, noise, float(a ** 0.5), float((1 - a) ** 0.5))
