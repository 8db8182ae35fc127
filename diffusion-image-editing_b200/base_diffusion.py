"""Unguided sampling wrapper (drop-in for src/base_diffusion.py)."""
from typing import Optional

import torch

from diffusion_utils import diffusion_loop, get_noise_pred, get_variance_noise, single_step
from transforms import tensor_to_pil
from utils import (create_progress_bar, generate_random_samples, get_device, initialize_random_samples,
                   process_lists_of_tensors, set_seed)


class Diffusion:
    """Holds a pipeline-like ``model`` (.unet, .scheduler, .device) and samples from it."""

    decode_is_identity = False

    def __init__(self, model) -> None:
        self.device = getattr(model, "device", None) or get_device()
        self.model = model
        self.unet = model.unet
        self.scheduler = model.scheduler
        self.data_dimensionality = self.unet.config.sample_size

    def encode(self, sample: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def additional_prep(self, model, prompt):
        return None

    def generate_image(self, xt, eta=0, zs=None, num_inference_steps=50, generator=None, prompt="",
                       cfg_scale=3.5, return_xts=False):
        """DDIM / DDPM sampling from x_T.  Returns (PIL image, [eps], [PIL x0-predictions], [PIL x_t] | None)."""
        self.model.scheduler.set_timesteps(num_inference_steps)
        text_emb = self.additional_prep(self.model, prompt)
        eps_hist, x0_hist, xt_hist = [], [], []
        if eta > 0 and zs is None:
            zs = generate_random_samples(num_inference_steps, self.model.unet, generator=generator)
        for step_idx, timestep in diffusion_loop(model=self.model, zs=zs):
            eps = get_noise_pred(self.model, xt, timestep, text_emb, cfg_scale)
            xt, x0_pred = single_step(self.model, eps, timestep, xt, eta, get_variance_noise(zs, step_idx, eta))
            eps_hist.append(eps)
            x0_hist.append(x0_pred)
            if return_xts:
                xt_hist.append(xt)
        img = tensor_to_pil(self.decode(xt))
        x0_imgs = process_lists_of_tensors(self, x0_hist)
        xt_imgs = process_lists_of_tensors(self, xt_hist) if return_xts else None
        return img, eps_hist, x0_imgs, xt_imgs

    def generate_images(self, num_images: int = 1, eta: float = 0, num_inference_steps: int = 50,
                        seed: Optional[int] = None, show_progbar: bool = True,
                        return_pred_original_samples: bool = True, prompt: str = "", cfg_scale: float = 3.5,
                        return_xts: bool = False):
        generator = set_seed(seed)
        self.scheduler.set_timesteps(num_inference_steps)
        all_xts, all_zs, all_imgs, all_eps, all_x0 = [], [], [], [], []
        for _ in create_progress_bar(range(num_images), show_progbar):
            xt, zs = initialize_random_samples(self.model, num_inference_steps=num_inference_steps, eta=eta,
                                               generator=generator)
            all_xts.append(xt)
            all_zs.append(zs)
            img, eps_hist, x0_imgs, _ = self.generate_image(
                xt=xt, eta=eta, zs=zs, num_inference_steps=num_inference_steps, generator=generator,
                prompt=prompt, cfg_scale=cfg_scale, return_xts=return_xts)
            all_imgs.append(img)
            all_eps.append(eps_hist)
            if return_pred_original_samples:
                all_x0.append(x0_imgs)
        return all_imgs, all_eps, all_x0, all_xts, all_zs
