"""Oracle restatement of the reference-owned element-wise step math.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED: every function here is
checked by tests/test_oracle_golden.py against vectors produced by the unmodified
reference (tests/golden/make_golden.py).

Design: per-step scalar coefficients are gathered once into ``StepCoeffs`` (fp32
0-d tensors, formed with the reference's op order), and each tensor formula is a
pure function of (tensors, coeffs).  The op ORDER of every expression matters for
bit-parity in fp32 (e.g. division by sqrt(alpha), not multiplication by 1/sqrt).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class StepCoeffs:
    """fp32 0-d tensors for one timestep t (and its predecessor t_prev)."""
    t: int
    t_prev: int
    a_t: torch.Tensor        # alphas_cumprod[t]
    a_prev: torch.Tensor     # alphas_cumprod[t_prev] or final_alpha_cumprod
    sqrt_a_t: torch.Tensor   # a_t ** 0.5
    sqrt_b_t: torch.Tensor   # (1 - a_t) ** 0.5
    sqrt_a_prev: torch.Tensor
    variance: torch.Tensor   # ((1-a_prev)/(1-a_t)) * (1 - a_t/a_prev)
    a_t_sq: torch.Tensor     # a_t ** 2  (guidance step size)


def step_coeffs(scheduler, t) -> StepCoeffs:
    """src/diffusion_utils.py:6-24,76-81 (calculate_variance, compute_alpha_products,
    get_previous_timestep)."""
    t = int(t)
    stride = scheduler.config.num_train_timesteps // scheduler.num_inference_steps
    t_prev = t - stride
    ac = scheduler.alphas_cumprod
    a_t = ac[t]
    a_prev = ac[t_prev] if t_prev >= 0 else scheduler.final_alpha_cumprod
    variance = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
    return StepCoeffs(t, t_prev, a_t, a_prev, a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5,
                      variance, a_t ** 2)


def pred_x0(sample, eps, c: StepCoeffs):
    """src/diffusion_utils.py:27-31  x0 = (x_t - sqrt(1-a_t) eps) / sqrt(a_t)."""
    return (sample - c.sqrt_b_t * eps) / c.sqrt_a_t


def ddim_step(sample, eps, c: StepCoeffs, eta=0.0, noise=None, clip=False, clip_range=1.0):
    """``single_step`` -> ``DDIMScheduler.step`` (src/diffusion_utils.py:90-109; DDIM eq. 12).
    Returns (x_prev, x0_pred)."""
    x0 = pred_x0(sample, eps, c)
    if clip:
        x0 = x0.clamp(-clip_range, clip_range)
    sigma = eta * c.variance ** 0.5
    direction = (1 - c.a_prev - sigma ** 2) ** 0.5 * eps
    x_prev = c.sqrt_a_prev * x0 + direction
    if eta > 0:
        x_prev = x_prev + sigma * noise
    return x_prev, x0


def ddpm_reverse_step(sample, eps, c: StepCoeffs, eta=0.0, noise=None):
    """``reverse_step`` (src/ddpm_inversion.py:203-240): no clipping; the direction
    coefficient uses eta*variance (not sigma**2).  Returns (x_prev, x0_pred)."""
    x0 = pred_x0(sample, eps, c)
    direction = (1 - c.a_prev - eta * c.variance) ** 0.5 * eps
    x_prev = c.sqrt_a_prev * x0 + direction
    if eta > 0:
        x_prev = x_prev + eta * c.variance ** 0.5 * noise
    return x_prev, x0


def extract_noise(x_t, x_tm1, eps, c: StepCoeffs, eta):
    """Edit-friendly inversion, one step (src/ddpm_inversion.py:135-169).
    Returns (z_t, corrected x_{t-1})."""
    x0 = pred_x0(x_t, eps, c)
    direction = (1 - c.a_prev - eta * c.variance) ** 0.5 * eps
    mu = c.sqrt_a_prev * x0 + direction
    z = (x_tm1 - mu) / (eta * c.variance ** 0.5)
    return z, mu + (eta * c.variance ** 0.5) * z


def ddim_next_step(sample, eps, scheduler, t):
    """``next_step`` (src/ddim_inversion.py:13-48)."""
    t = int(t)
    stride = scheduler.config.num_train_timesteps // scheduler.num_inference_steps
    t_cur, t_next = min(t - stride, 999), t
    ac = scheduler.alphas_cumprod
    a_cur = ac[t_cur] if t_cur >= 0 else scheduler.final_alpha_cumprod
    a_next = ac[t_next]
    x0 = (sample - (1 - a_cur) ** 0.5 * eps) / a_cur ** 0.5
    return a_next ** 0.5 * x0 + (1 - a_next) ** 0.5 * eps


def ddpm_forward_step(sample, eps, scheduler, t):
    """``forward_step`` (src/ddpm_inversion.py:58-77): eta=0 inversion step."""
    t = int(t)
    n = scheduler.config.num_train_timesteps
    t_next = min(n - 2, t + n // scheduler.num_inference_steps)
    ac = scheduler.alphas_cumprod
    a_t = ac[t]
    x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return ac[t_next] ** 0.5 * x0 + (1 - ac[t_next]) ** 0.5 * eps


def sample_xts(x0, scheduler, noises):
    """``sample_xts_from_x0`` (src/ddpm_inversion.py:31-55) with the noise injected:
    ``noises[idx]`` is the draw used for timesteps[idx].  x0: (1,C,H,W) -> (T+1,C,H,W)."""
    ac = scheduler.alphas_cumprod
    sqrt_1m = (1 - ac) ** 0.5
    ts = scheduler.timesteps
    out = torch.zeros((len(ts),) + tuple(x0.shape[1:]), dtype=x0.dtype)
    for idx, t in enumerate(ts):
        t = int(t)
        out[idx] = (x0 * (ac[t] ** 0.5) + noises[idx] * sqrt_1m[t])[0]
    return torch.cat([out, x0], dim=0)


def cfg_combine(e_first, e_second, scale):
    """src/diffusion_utils.py:67-70  (first half is 'uncond' by convention)."""
    return e_first + scale * (e_second - e_first)


def apply_mask(mask, zo, zv):
    """src/utils.py:23-28."""
    return mask * zv + ((1 - mask) * zo)


def to_uint8_image(x):
    """``tensor_to_pil`` numerics (src/transforms.py:8-35): trunc(clamp(x/2+0.5,0,1)*255)."""
    return ((x / 2 + 0.5).clamp(0, 1) * 255).to(torch.uint8)


# --------------------------------------------------------------------------- guidance


def color_guidance_update(x_post, eps, c: StepCoeffs, targets, weights, loss_scale,
                          mask=None, mask_grad=False, n_mean=None):
    """Closed form of ``AttrFunc.apply`` for Single/MultiColorAttrFunc with decode =
    identity (src/attr_functions.py:22-37,112-163).  ``targets[ch]`` is the colour target
    of channel ch or None; ``weights[ch]`` the loss weight (1 for single colour, the
    target itself for multi colour, :35).  Mirrors the autograd chain exactly:
      k = loss_scale*w/N ; g = -((k*sign(x0g - tau)) / sqrt_a_t) ; x += g * a_t**2
    """
    x0g = (x_post - c.sqrt_b_t * eps) / c.sqrt_a_t
    b, ch, h, w = x_post.shape
    n = float(b * h * w) if n_mean is None else float(n_mean)
    g = torch.zeros_like(x_post)
    for k in range(ch):
        if targets[k] is None:
            continue
        s = torch.sign(x0g[:, k] - targets[k])
        # autograd order: d(loss*scale)/d(mean) = scale (*w) ; /N ; *sign ; /sqrt_a
        coef = torch.tensor(float(loss_scale), dtype=torch.float32)
        if weights[k] != 1.0:
            coef = coef * torch.tensor(float(weights[k]), dtype=torch.float32)
        coef = coef / n
        g[:, k] = -((coef * s) / c.sqrt_a_t)
    if mask_grad and mask is not None:
        g = mask * g
    return x_post + g * c.a_t_sq, g


def autograd_guidance_update(x_post, eps, c: StepCoeffs, loss_fn, loss_scale,
                             mask=None, mask_grad=False):
    """Generic ``AttrFunc.apply`` via autograd (src/attr_functions.py:112-163) for any
    differentiable ``loss_fn(x0g)``; decode = identity."""
    x = x_post.detach().clone().requires_grad_(True)
    x0g = (x - c.sqrt_b_t * eps) / c.sqrt_a_t
    loss = loss_fn(x0g) * loss_scale
    g = -torch.autograd.grad(loss, x)[0]
    if mask_grad and mask is not None:
        g = mask * g
    return x.detach() + g * c.a_t_sq, g


def l2reg_loss(x0g, mask, x_ref, lam, base_loss):
    """``calculate_loss`` masked + L2-regularised variant (src/attr_functions.py:82-96):
    base(mask*img) + lambda * || (1 - mask*img) - x_0 ||_2."""
    img = mask * x0g
    return base_loss(img) + lam * torch.sqrt(torch.sum(((1 - img) - x_ref) ** 2))


def single_color_loss(img, idx, target):
    """src/attr_functions.py:22-25."""
    return torch.abs(img[:, idx] - target).mean()


def multi_color_loss(img, r, g, b):
    """src/attr_functions.py:28-37 (weights are the targets)."""
    return (single_color_loss(img, 0, r) * r + single_color_loss(img, 1, g) * g
            + single_color_loss(img, 2, b) * b)


def segmentation_area_loss(logits, class_ids):
    """``NetAttrFunc.loss`` head (src/attr_functions.py:213-219): softmax over classes,
    per-class area / (256*256) (literal), summed over the selected classes."""
    p = logits.squeeze(0).softmax(dim=0)
    area = p.sum(dim=(1, 2)) / (256 * 256)
    return area[class_ids].sum()


def classifier_logit_loss(logits80, idx_for_class, idx_of_interest=0, reg=(None, None, None)):
    """``ClassifierAttrFunc.loss`` head (src/attr_functions.py:237-257)."""
    attr = logits80.view(-1, 40, 2)
    val = attr[0][idx_for_class][idx_of_interest]
    if reg[0] is not None:
        other = attr[0][reg[0]][reg[1]]
        val = val + (other + reg[2][reg[1]]) ** 2
    return val
