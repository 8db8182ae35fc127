"""DDIM inversion x_0 -> x_T (drop-in for src/ddim_inversion.py)."""
from typing import Optional

import torch
from tqdm import tqdm

from b200edit import ops
from diffusion_utils import encode_text, get_noise_pred


def next_step(model, model_output, timestep: int, sample):
    """x_{t+1} from x_t (src/ddim_inversion.py:13-48): one fused kernel
    out = sqrt(a_next) * ((x - sqrt(1-a_cur) e)/sqrt(a_cur)) + sqrt(1-a_next) * e."""
    sch = model.scheduler
    t_next = int(timestep)
    t_cur = min(t_next - sch.config.num_train_timesteps // sch.num_inference_steps, 999)
    a_cur = sch.alphas_cumprod[t_cur] if t_cur >= 0 else sch.final_alpha_cumprod
    a_next = sch.alphas_cumprod[t_next]
    return ops.renoise(sample, model_output, float(a_cur ** 0.5), float((1 - a_cur) ** 0.5),
                       float(a_next ** 0.5), float((1 - a_next) ** 0.5))


@torch.no_grad()
def ddim_loop(model, latent, prompt: Optional[str] = None, cfg_scale=3.5):
    context = None
    if prompt is not None:
        context = torch.cat([encode_text(model, prompt), encode_text(model, "")])
    T = model.scheduler.num_inference_steps
    ts = [int(t) for t in model.scheduler.timesteps]
    for i in tqdm(range(T)):
        t = ts[len(ts) - i - 1]
        eps = get_noise_pred(model, latent, torch.tensor(t), text_emb=context, cfg_scale=cfg_scale)
        latent = next_step(model, eps, t, latent)
    return latent


@torch.no_grad()
def ddim_inversion(model, x0, prompt: Optional[str] = None, cfg_scale: float = 3.5):
    return ddim_loop(model, x0.clone().detach(), prompt=prompt, cfg_scale=cfg_scale)
