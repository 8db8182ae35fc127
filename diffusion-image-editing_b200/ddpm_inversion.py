"""Edit-friendly DDPM inversion and regeneration (drop-in for src/ddpm_inversion.py; algorithm of
Huberman-Spiegelglas et al., arXiv:2304.06140).  Index conventions: xts[idx] is the sample at
scheduler.timesteps[idx] (idx 0 = noisiest), xts[T] = x0; zs[idx] is the noise injected when
stepping from timesteps[idx] to the next cleaner level; zs[T-1] = 0."""
from typing import Optional

import torch
from tqdm import tqdm

from b200edit import ops
from diffusion_utils import (calculate_variance, compute_alpha_products, compute_predicted_original_sample,
                             encode_text, get_noise_pred, get_previous_timestep, get_variance_noise)

__all__ = ["mu_tilde", "sample_xts_from_x0", "forward_step", "inversion_forward_process", "invert",
           "reverse_step", "inversion_reverse_process", "sample", "calculate_variance", "get_variance_noise"]


def mu_tilde(model, xt, x0, timestep):
    """Posterior mean mu~(x_t, x_0), DDPM eq. 7 (src/ddpm_inversion.py:16-28)."""
    prev_t = get_previous_timestep(model, int(timestep))
    a_t, a_prev = compute_alpha_products(model, int(timestep), prev_t)
    a_bar = model.scheduler.alphas_cumprod[int(timestep)]
    c0 = (a_prev ** 0.5 * (1 - a_t)) / (1 - a_bar)
    ct = (a_t ** 0.5 * (1 - a_prev)) / (1 - a_bar)
    return ops.axpby(x0, xt, float(c0), float(ct))


def sample_xts_from_x0(model, x0, num_inference_steps=50, noise: Optional[torch.Tensor] = None):
    """Independent forward-noised copies x_t = sqrt(a_t) x0 + sqrt(1-a_t) n_t for every inference
    timestep, plus x0 itself: (T+1,C,H,W) (src/ddpm_inversion.py:31-55).  ``noise`` (T,C,H,W) may be
    injected (noise[idx] belongs to timesteps[idx]); otherwise it is drawn like the reference draws
    it: one randn_like(x0) per timestep, ascending t, on x0's device generator."""
    sch = model.scheduler
    ts = sch.timesteps
    T = len(ts)
    x0 = x0.reshape(1, *x0.shape[-3:])
    if noise is None:
        draws = [None] * T
        for idx in reversed(range(T)):
            draws[idx] = torch.randn_like(x0)[0]
        noise = torch.stack(draws)
    ac = sch.alphas_cumprod
    sa = (ac[ts] ** 0.5).to(x0.device)
    sb = ((1 - ac) ** 0.5)[ts].to(x0.device)
    return ops.sample_xts(x0, noise, sa, sb)


def forward_step(model, model_output, timestep, sample):
    """eta = 0 inversion step: x0-prediction re-noised to t + stride (src/ddpm_inversion.py:58-77)."""
    sch = model.scheduler
    t = int(timestep)
    n = sch.config.num_train_timesteps
    t_next = min(n - 2, t + n // sch.num_inference_steps)
    a_t, a_n = sch.alphas_cumprod[t], sch.alphas_cumprod[t_next]
    return ops.renoise(sample, model_output, float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_n ** 0.5),
                       float((1 - a_n) ** 0.5))


def inversion_forward_process(model, x0, etas=None, num_inference_steps=50, prompt: Optional[str] = None,
                              cfg_scale: float = 3.5, prog_bar=False, noise: Optional[torch.Tensor] = None):
    context = None
    if prompt is not None:
        context = torch.cat([encode_text(model, prompt), encode_text(model, "")])
    sch = model.scheduler
    sch.set_timesteps(num_inference_steps)
    ts = [int(t) for t in sch.timesteps]
    T = len(ts)
    eta_is_zero = etas is None or (type(etas) in [int, float] and etas == 0)
    if not eta_is_zero:
        if type(etas) in [int, float]:
            etas = [etas] * T
        xts = sample_xts_from_x0(model, x0, num_inference_steps=num_inference_steps, noise=noise)
        zs = torch.zeros((T,) + tuple(xts.shape[1:]), device=xts.device, dtype=torch.float32)
    else:
        xts, zs = None, None
    xt = x0
    order = list(reversed(range(T)))
    for idx in (tqdm(order) if prog_bar else order):
        t = ts[idx]
        if not eta_is_zero:
            xt = xts[idx][None]
        eps = get_noise_pred(model, xt, torch.tensor(t), context, cfg_scale)
        if eta_is_zero:
            xt = forward_step(model, eps, t, xt)
        else:
            # z_t = (x_{t-1} - mu_t)/(eta sqrt(var)); x_{t-1} <- mu_t + eta sqrt(var) z_t  (one kernel, in place)
            ops.extract_noise(xt, eps, xts[idx + 1], zs[idx], sch.coeffs(t, etas[idx], "ddpm"))
    if zs is not None:
        zs[-1].zero_()
    return xt, zs, xts


def invert(model, x0, num_inference_steps=50, eta=1, prompt: Optional[str] = None, cfg_scale: float = 3.5,
           prog_bar=True, noise: Optional[torch.Tensor] = None):
    """Returns (x_T, zs, xts) - Algorithm 1 of arXiv:2304.06140."""
    return inversion_forward_process(model, x0, num_inference_steps=num_inference_steps, etas=eta,
                                     prompt=prompt, cfg_scale=cfg_scale, prog_bar=prog_bar, noise=noise)


def reverse_step(model, model_output, timestep, sample, eta=0, variance_noise=None):
    """DDPM-style reverse step: no clipping, direction coefficient sqrt(1 - a_prev - eta*var)
    (src/ddpm_inversion.py:203-240)."""
    if eta > 0 and variance_noise is None:
        variance_noise = torch.randn(model_output.shape, device=model_output.device)
    prev, _ = ops.guided_step(sample, model_output, model.scheduler.coeffs(int(timestep), eta, "ddpm"),
                              noise=variance_noise if eta > 0 else None, want_x0=False)
    return prev


def inversion_reverse_process(model, xT, eta=0, zs=None, prompt: Optional[str] = None, cfg_scale: float = 3.5,
                              prog_bar=False):
    context = encode_text(model, prompt) if prompt is not None else None
    ts = [int(t) for t in model.scheduler.timesteps][-zs.shape[0]:]
    xt = xT.expand(1, -1, -1, -1) if xT.dim() == 3 else xT
    for idx, t in enumerate(tqdm(ts) if prog_bar else ts):
        eps = get_noise_pred(model, xt, torch.tensor(t), context, cfg_scale)
        xt = reverse_step(model, eps, t, xt, eta=eta, variance_noise=get_variance_noise(zs, idx, eta))
    return xt, zs


def sample(model, zs, xts, Tskip=36, eta=1, prompt: Optional[str] = None, cfg_scale: float = 3.5, prog_bar=True):
    """Regenerate from xts[Tskip] with the extracted noise maps zs[Tskip:]."""
    x0, _ = inversion_reverse_process(model, xT=xts[Tskip], eta=eta, zs=zs[Tskip:], prompt=prompt,
                                      cfg_scale=cfg_scale, prog_bar=prog_bar)
    return x0[None] if x0.dim() < 4 else x0
