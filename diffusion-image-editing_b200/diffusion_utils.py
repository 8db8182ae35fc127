"""Step math / scheduler glue (drop-in for src/diffusion_utils.py).

Same names, arguments and return types as the reference; tensor arithmetic runs in hand-written
CUDA kernels (b200edit.ops), scalar coefficients stay fp32 0-d host tensors formed exactly like
the reference forms them, and the loop helpers never synchronise with the device."""
from typing import Optional, Tuple

import torch
from tqdm import tqdm

from b200edit import ops


def get_previous_timestep(model, timestep):
    return timestep - model.scheduler.config.num_train_timesteps // model.scheduler.num_inference_steps


def compute_alpha_products(model, timestep, prev_timestep):
    sch = model.scheduler
    a_t = sch.alphas_cumprod[int(timestep)]
    a_prev = sch.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else sch.final_alpha_cumprod
    return a_t, a_prev


def calculate_variance(model, timestep):
    a_t, a_prev = compute_alpha_products(model, timestep, get_previous_timestep(model, timestep))
    return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)


def compute_predicted_original_sample(sample, beta_prod_t, model_output, alpha_prod_t):
    """x0 = (x_t - sqrt(1-a_t) eps) / sqrt(a_t)  - DDIM eq. 12 (src/diffusion_utils.py:27-31)."""
    sa = torch.as_tensor(alpha_prod_t, dtype=torch.float32) ** 0.5
    sb = torch.as_tensor(beta_prod_t, dtype=torch.float32) ** 0.5
    return ops.pred_x0(sample, model_output, float(sa), float(sb))


def tokenize_text(model, prompt):
    return model.tokenizer([prompt], padding="max_length", max_length=model.tokenizer.model_max_length,
                           truncation=True, return_tensors="pt")


def encode_text(model, prompts):
    text_input = tokenize_text(model, prompts)
    with torch.no_grad():
        return model.text_encoder(text_input.input_ids.to(model.device))[0]


def prep_text(model, prompt: str) -> torch.Tensor:
    return torch.cat([encode_text(model, ""), encode_text(model, prompt)])


def get_noise_pred(model, latent, t, text_emb: Optional[torch.Tensor] = None, cfg_scale: float = 3.5):
    """Predicted noise; with a text embedding, classifier-free guidance
    e_first + s*(e_second - e_first) over the doubled batch (src/diffusion_utils.py:55-73)."""
    with torch.no_grad():
        if text_emb is None:
            return model.unet(latent, t)["sample"]
        both = model.unet(sample=torch.cat([latent] * 2), timestep=t, encoder_hidden_states=text_emb)["sample"]
        first, second = both.chunk(2)
        return ops.cfg_combine(first, second, cfg_scale)


def get_variance_noise(zs: Optional[torch.Tensor], step_idx: int, eta: float):
    return zs[step_idx] if zs is not None and eta != 0 else None


def single_step(model, model_output, timestep, sample, eta, variance_noise) -> Tuple[torch.Tensor, torch.Tensor]:
    """One DDIM step through the scheduler: (prev_sample, pred_original_sample)."""
    out = model.scheduler.step(model_output=model_output, timestep=timestep, sample=sample, eta=eta,
                               variance_noise=variance_noise)
    return out.to_tuple()


def diffusion_loop(model, zs=None, prog_bar=True):
    """Yields (step_idx, timestep) over the last len(zs) (or all) scheduler timesteps; step_idx
    restarts at 0 inside a Tskip-trimmed window (src/diffusion_utils.py:112-133).  Timesteps are
    host integers wrapped as 0-d CPU tensors: iterating never touches the device."""
    timesteps = model.scheduler.timesteps
    n = zs.shape[0] if zs is not None else len(timesteps)
    window = [int(t) for t in timesteps[-n:]]
    it = tqdm(window) if prog_bar else window
    for step_idx, t in enumerate(it):
        yield step_idx, torch.tensor(t)
